"""Helpers shared by the GPU parity tests: move seeded numpy operands to the device and back."""
import numpy as np
import torch


def to_dev(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()


def to_host(t):
    return t.cpu().numpy().view(np.uint64)

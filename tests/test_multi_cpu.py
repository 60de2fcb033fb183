"""CPU tests of the multi-GPU host logic with world_size-2 gloo process groups (no GPU):
the limb-sharded key switch's ownership / all-gather layout as reported by the C library, and the data-parallel
sharding + max-over-ranks timing reduction used by bench.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import homulator_b200 as hml


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gather_worker(rank, world, port, L, alpha, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lay = hml.shard_layout(L, alpha, rank, world)
        N = 8
        # gather 1: this rank's slot holds its owned Q-limbs, each filled with its global limb id
        g1 = torch.full((world, lay["gather1_slots"], N), -1, dtype=torch.int64)
        for k, i in enumerate(lay["own_q"]):
            g1[rank, k] = i
        parts = [torch.empty_like(g1[0]) for _ in range(world)]
        dist.all_gather(parts, g1[rank].clone())
        g1 = torch.stack(parts)
        ok = all(int(g1[lay["owner"][i], lay["slot"][i], 0]) == i for i in range(L))
        # gather 2: [world][2][slots][N], accumulator c of owned P-limb j tagged 1000*c + j
        g2 = torch.full((world, 2, lay["gather2_slots"], N), -1, dtype=torch.int64)
        for k, j in enumerate(lay["own_p"]):
            for c in range(2):
                g2[rank, c, k] = 1000 * c + j
        parts = [torch.empty_like(g2[0]) for _ in range(world)]
        dist.all_gather(parts, g2[rank].clone())
        g2 = torch.stack(parts)
        for j in range(alpha):
            e = L + j
            for c in range(2):
                ok = ok and int(g2[lay["owner"][e], c, lay["slot"][e], 0]) == 1000 * c + j
        # every limb has exactly one owner, and the owned sets partition the extended basis
        owned = torch.zeros(L + alpha, dtype=torch.int64)
        for i in lay["own_q"]:
            owned[i] += 1
        for j in lay["own_p"]:
            owned[L + j] += 1
        dist.all_reduce(owned)
        ok = ok and bool((owned == 1).all())
        # max-over-ranks timing reduction, as bench.py does it
        t = torch.tensor([10.0 + rank])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = ok and float(t) == 10.0 + world - 1
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("L,alpha", [(35, 15), (7, 3), (2, 15), (5, 2)])
def test_shard_layout_all_gather_world2(L, alpha):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_gather_worker, args=(world, _free_port(), L, alpha, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))


def test_shard_layout_matches_reference_cluster_rule():
    """the reference maps limb l to cluster l % cluster (include/Driver.h:158,:178)"""
    for world in (1, 2, 3, 4, 8):
        seen = set()
        for r in range(world):
            lay = hml.shard_layout(35, 15, r, world)
            assert lay["owner"] == [e % world for e in range(50)]
            assert all(i % world == r for i in lay["own_q"]) and all((35 + j) % world == r for j in lay["own_p"])
            assert len(lay["own_q"]) <= lay["gather1_slots"] and len(lay["own_p"]) <= lay["gather2_slots"]
            seen |= set(lay["own_q"]) | {35 + j for j in lay["own_p"]}
        assert seen == set(range(50))

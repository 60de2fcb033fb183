"""First-light check of the tcgen05 base conversion: a few shapes against the oracle, then timing of both kernels
(HML_BCONV_UMMA=0/1 are read once per process, so the comparison runs this file twice)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from homulator_b200 import api as hml  # noqa: E402
from orc import Oracle, uniform_limbs  # noqa: E402
from gpu_common import to_dev, to_host  # noqa: E402


def main():
    N, ml, al = 65536, 45, 15
    ctx, o = hml.Context(N=N, max_level=ml, alpha=al), Oracle(N, 36, ml, al)
    ok = True
    for src, dst in ((list(range(15)), list(range(15, 35)) + list(range(45, 60))), (list(range(45, 60)), list(range(35))),
                     (list(range(30, 35)), list(range(30)) + list(range(45, 60)))):
        x = np.stack([uniform_limbs([o.moduli[i]], N, 400 + i)[0] for i in src])
        got = to_host(ctx.bconv(to_dev(x), src, dst))
        bad = 0
        for t in (0, 1, len(dst) // 2, len(dst) - 1):
            want = o.bconv(src, dst[t], x)
            if not np.array_equal(got[t], want):
                bad += 1
                d = np.nonzero(got[t] != want)[0]
                print("  mismatch target", t, "count", len(d), "first idx", d[:8], "got", got[t][d[:3]], "want", want[d[:3]])
        print("bconv %d -> %d: %s" % (len(src), len(dst), "OK" if not bad else "MISMATCH"), flush=True)
        ok &= not bad
    # timing: the ModDown shape of the bench (64 batches need the op path; here 1 batch x 35 targets, repeated)
    src, dst = list(range(45, 60)), list(range(35))
    xd = ctx.uniform(src, 7)
    out = ctx.bconv(xd, src, dst)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        ctx.bconv(xd, src, dst, out=out)
    e1.record()
    torch.cuda.synchronize()
    print("bconv 15->35 single batch: %.2f us per launch" % (e0.elapsed_time(e1) * 1000 / 50))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())

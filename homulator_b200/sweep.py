"""Parameter-set sweeps A-D over the current level, on real data.

The reference's benchmark matrix (reference script/README.md:17-22, script/para*/micro24_*_<op>.sh) starts one simulator
process per (operation, level) and writes `outLogs/para<S>/<cluster>/<op>/<maxLevel>_<alpha>/<op>_<maxLevel>_<alpha>_<level>.log`.
This module walks the same matrix through the C ABI on one GPU: every (set, op, level) is executed on seeded synthetic
operands, timed with CUDA events (L2 flushed between runs), and logged under the same path layout (one JSON line per
log), plus a markdown summary.

    python -m homulator_b200.sweep [--sets A,B,C,D] [--ops hmult,hrotate,hadd,pmult,padd] [--step 1] [--iters 5]
                                   [--out outLogs] [--summary profiles/sweep.md]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# reference script/README.md:17-22 (N, maxLevel, alpha, config file) and the lowest level each script reaches
SETS = {
    "A": dict(cfg="config_4_N15.cfg", max_level=28, alpha=28),
    "B": dict(cfg="config_4.cfg", max_level=45, alpha=15),
    "C": dict(cfg="config_4.cfg", max_level=24, alpha=6),
    "D": dict(cfg="config_4.cfg", max_level=26, alpha=9),
}
MIN_LEVEL = {"hmult": 2, "hrotate": 1, "hadd": 1, "pmult": 1, "padd": 1}  # micro24_*_hmult.sh stops at 2


def run(sets, ops, step, iters, out_dir, summary):
    import torch
    import homulator_b200 as hml

    flush = torch.empty(64 << 20, dtype=torch.int64, device="cuda")  # 512 MiB > L2
    rows = []
    for name in sets:
        ps = SETS[name]
        ctx = hml.Context(os.path.join(ROOT, "config", ps["cfg"]), ps["max_level"], ps["alpha"])
        ML, A = ps["max_level"], ps["alpha"]
        for op in ops:
            for L in range(ML, MIN_LEVEL[op] - 1, -step):
                q = list(range(L))
                a = ctx.uniform(q, 1, lead=(2,))
                b = ctx.uniform(q, 2, lead=(2,))
                pt = b[0]
                key = ctx.uniform(ctx.ext_mod_idx(L), 3, lead=(ctx.beta(L), 2)) if op in ("hmult", "hrotate") else None
                fn = {"hmult": lambda: ctx.hmult(L, a, b, key), "hrotate": lambda: ctx.hrotate(L, a, key, 5),
                      "hadd": lambda: ctx.hadd(L, a, b), "pmult": lambda: ctx.pmult(L, a, pt),
                      "padd": lambda: ctx.padd(L, a, pt)}[op]
                fn()
                ts = []
                for _ in range(iters):
                    flush.fill_(1)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    fn()
                    e1.record()
                    e1.synchronize()
                    ts.append(e0.elapsed_time(e1) * 1e3)
                ts.sort()
                c = ctx.counts(op, L)
                words = hml.algorithmic_words(op, L, A)
                rec = {"set": name, "op": op, "N": ctx.N, "maxLevel": ML, "L": L, "alpha": A, "us_median": ts[len(ts) // 2],
                       "us_min": ts[0], "iters": iters, "l2_flushed": True, "trace_total": c["total"], "driverTotal": c["driverTotal"],
                       "algorithmic_bytes": words * 8 * ctx.N}
                rows.append(rec)
                if out_dir:
                    d = os.path.join(out_dir, "para" + name, "gpu", op, "%d_%d" % (ML, A))
                    os.makedirs(d, exist_ok=True)
                    with open(os.path.join(d, "%s_%d_%d_%d.log" % (op, ML, A, L)), "w") as f:
                        f.write(json.dumps(rec) + "\n")
                del a, b, key
        ctx.close()
    if summary:
        with open(summary, "w") as f:
            f.write("# Parameter-set sweeps A-D on one B200 (single op, L2 flushed between runs; `python -m homulator_b200.sweep`)\n\n")
            f.write("Reference matrix: script/README.md:17-22, script/para*/micro24_*.sh (one simulator run per level).\n"
                    "`trace` = reference-shaped instruction count of the op at that level (equal to the reference's InsGen trace).\n\n")
            for name in sets:
                ps = SETS[name]
                f.write("## Set %s: %s, maxLevel %d, alpha %d\n\n" % (name, ps["cfg"], ps["max_level"], ps["alpha"]))
                sel = [r for r in rows if r["set"] == name]
                levels = sorted({r["L"] for r in sel}, reverse=True)
                f.write("| L | " + " | ".join("%s us" % o for o in ops) + " | hmult trace | hmult GB/s (unfused bytes) |\n")
                f.write("|---|" + "---|" * (len(ops) + 2) + "\n")
                for L in levels:
                    by = {r["op"]: r for r in sel if r["L"] == L}
                    cells = ["%.1f" % by[o]["us_median"] if o in by else "-" for o in ops]
                    hm = by.get("hmult")
                    f.write("| %d | %s | %s | %s |\n" % (L, " | ".join(cells), hm["trace_total"] if hm else "-",
                                                      "%.0f" % (hm["algorithmic_bytes"] / hm["us_median"] / 1e3) if hm else "-"))
                f.write("\n")
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sets", default="A,B,C,D")
    ap.add_argument("--ops", default="hmult,hrotate,hadd,pmult,padd")
    ap.add_argument("--step", type=int, default=1)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--out", default="outLogs")
    ap.add_argument("--summary", default="")
    args = ap.parse_args()
    sys.path.insert(0, ROOT)
    rows = run([s for s in args.sets.split(",") if s], [o for o in args.ops.split(",") if o], max(1, args.step), args.iters,
               args.out, args.summary)
    print(json.dumps({"runs": len(rows), "sets": args.sets, "ops": args.ops}))
    return 0


if __name__ == "__main__":
    sys.exit(main())

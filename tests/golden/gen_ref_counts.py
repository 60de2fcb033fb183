#!/usr/bin/env python3
"""Regenerate tests/golden/ref_counts.json from the UNMODIFIED reference.

Runs oracle/_ref/count.run (oracle/ref_count_harness.cpp linked with the reference's own
src/*.cpp, built by `make -C oracle ref`) for every (cfg, op, maxLevel, L, alpha) listed below and
stores the per-opcode / per-stage Instruction tallies.  Only runs in the build container
(/root/reference is not present on the GPU box); the JSON it writes is the committed fixture.
"""
import json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
RUN = os.path.join(ROOT, "oracle", "_ref", "count.run")

CASES = []
for L in (45, 35, 30, 20, 15, 2):
    CASES.append(("config_4.cfg", "hmult", 45, L, 15))
CASES += [("config_4.cfg", "hmult", 24, 24, 6), ("config_4.cfg", "hmult", 26, 26, 9),
          ("config_4.cfg", "hmult", 24, 7, 6), ("config_4.cfg", "hmult", 26, 10, 9)]
for L in (28, 14, 2):
    CASES.append(("config_4_N15.cfg", "hmult", 28, L, 28))
for L in (45, 35, 30, 20, 15, 2, 1):
    CASES.append(("config_4.cfg", "hrotate", 45, L, 15))
CASES += [("config_4.cfg", "hrotate", 24, 24, 6), ("config_4.cfg", "hrotate", 26, 26, 9),
          ("config_4_N15.cfg", "hrotate", 28, 28, 28), ("config_4_N15.cfg", "hrotate", 28, 5, 28)]
for op in ("hadd", "pmult", "padd"):
    CASES.append(("config_4.cfg", op, 45, 35, 15))
    CASES.append(("config_4_N15.cfg", op, 28, 3, 28))

def main():
    out = []
    for cfg, op, maxl, L, alpha in CASES:
        cmd = [RUN, os.path.join(REF, "config", cfg), op, str(maxl), str(L), str(alpha)]
        line = subprocess.run(cmd, check=True, capture_output=True, text=True).stdout.strip().splitlines()[-1]
        rec = json.loads(line)
        rec["cfg"] = cfg
        out.append(rec)
        print(cfg, op, maxl, L, alpha, rec["by_opcode"], rec["total"], rec["driverTotal"], file=sys.stderr)
    with open(os.path.join(ROOT, "tests", "golden", "ref_counts.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write("\n")

if __name__ == "__main__":
    main()

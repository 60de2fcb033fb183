"""GPU test of the drop-in CLI: `Homulator.run <cfg> <op> <maxLevel> <L> <alpha> [cluster]` — the
reference's only executable (bench_test/bench_micro24.cpp:5-52) — is executed on the device for all five operation names at
BASELINE.json configs[0]'s arguments, and the JSON line it prints is compared with the reference's own instruction trace
(tests/golden/ref_counts.json)."""
import json
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

import homulator_b200  # noqa: E402,F401  (builds / locates the library and the CLI)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "homulator_b200", "Homulator.run")
CFG = os.path.join(ROOT, "config", "config_4.cfg")
GOLD = {(r["cfg"], r["op"], r["maxLevel"], r["L"], r["alpha"]): r for r in json.load(open(os.path.join(ROOT, "tests", "golden", "ref_counts.json")))}


def run_cli(*args, timeout=600):
    r = subprocess.run([CLI, *[str(a) for a in args]], capture_output=True, text=True, timeout=timeout)
    return r


@pytest.mark.parametrize("op", ["hmult", "hrotate", "hadd", "pmult", "padd"])
def test_cli_runs_every_op_at_the_north_star_config(op):
    r = run_cli(CFG, op, 45, 35, 15, "--iters", 3, "--warmup", 1)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = r.stdout
    assert out.startswith("Configuration details are as follow:")  # the reference's config dump comes first (src/Config.cpp:40-51)
    assert "Welcome! Start executing %s" % op.upper() in out and "Completed!" in out
    line = json.loads(out.strip().splitlines()[-1])
    assert line["op"] == op and line["N"] == 65536 and line["L"] == 35 and line["us_median"] > 0
    gold = GOLD[("config_4.cfg", op, 45, 35, 15)]
    for opc in ("NTT", "INTT", "MULT", "BCONV_STEP2", "AUTO"):
        assert line["trace"][opc] == gold["by_opcode"].get(opc, 0), opc
    assert line["trace"]["total"] == gold["total"] and line["trace"]["driverTotal"] == gold["driverTotal"]
    ex = line["executed"]
    assert ex["kernel_launches"] > 0
    if op == "hmult":
        assert ex["ntt_limbs"] == 183 and ex["intt_limbs"] == 67
        assert "Malloc TensorD0Out from 0 to" in out
    if op == "hrotate":
        assert ex["auto_limbs"] == 70 and ex["ntt_limbs"] == 115 + 70 and ex["intt_limbs"] == 35 + 30
    if op in ("hadd", "pmult", "padd"):
        assert ex["ewe_limbs"] == 70
    # the per-kernel-class report (reference Statistic dump, include/Staistics.h:6-40)
    assert "NTT_(0) :" in out and "BCONV_(0) :" in out and "HBM_(0) :" in out


def test_cli_accepts_the_cluster_argument_and_rejects_bad_levels():
    r = run_cli(CFG, "hmult", 45, 2, 15, 1, "--iters", 2, "--warmup", 1)   # positional [cluster] like the reference
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["trace"]["total"] == GOLD[("config_4.cfg", "hmult", 45, 2, 15)]["total"]
    assert line["cluster"] == 1
    for bad in (0, 46, 99):
        r = run_cli(CFG, "hrotate", 45, bad, 15)
        assert r.returncode != 0 and "currentLevel" in r.stderr
    r = run_cli(CFG, "hmult", 45, 1, 15)
    assert r.returncode != 0 and "currentLevel" in r.stderr
    r = run_cli(CFG, "hrot", 45, 35, 15)   # reference: message on stdout, exit 0 (bench_micro24.cpp:49-51)
    assert r.returncode == 0 and "Error operation requirement" in r.stdout

"""The reference arm of ``bench.py`` runs on the CPU, so its side of the driver's contract is checked here: one JSON
line with the keys the driver reads, rank 0 only under a multi-rank launch, all host threads even when the launcher
exported OMP_NUM_THREADS=1 (torchrun does), and no GPU needed.  (The other arm needs a B200: tests -m gpu + bench runs.)"""
import json
import os
import subprocess
import sys

from helpers import ROOT


def _run(env_extra, *args):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1", *args],
                          capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)


def test_reference_arm_prints_one_contract_line_and_uses_the_host_threads():
    res = _run({"OMP_NUM_THREADS": "1", "RANK": "0", "WORLD_SIZE": "2"}, "--gpus", "2", "--steps", "1", "--warmup", "1")
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["steps"] == 1 and d["warmup"] == 1
    assert d["unit"] == "iter*problems/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and abs(d["value"] - 15 * d["config"]["batch_in_sample"] / (d["ms_per_step"] / 1e3)) < 1e-6 * d["value"]
    assert d["config"]["workload"].startswith("BASELINE configs[0]") and "model" not in d["config"]
    cb = d["cpu_baseline"]
    from oracle import ref_harness
    # the unmodified reference classes wherever they exist (checkout, or staged under oracle/_ref); the oracle port otherwise
    assert cb["kind"] == ("reference" if ref_harness.reference_root() else "port") and cb["value"] == d["value"] and cb["sample"]
    assert d["product_modules_loaded"] == []       # the CPU arm never imports the product package (nor its .so)
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    assert cb["cores"] == avail                    # not the launcher's OMP_NUM_THREADS=1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    res = _run({"RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2", "--steps", "1", "--warmup", "0")
    assert res.returncode == 0 and not [ln for ln in res.stdout.splitlines() if ln.startswith("{")]

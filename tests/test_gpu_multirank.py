"""Same-noise multi-GPU parity (needs >= 2 GPUs; skipped on a single-GPU box): N ranks under torchrun reproduce the
single-process loss, gradient and iterates on the same batch -- tests/probes/rank_parity.py."""
import os
import subprocess
import sys

import pytest
import torch

from helpers import ROOT

pytestmark = pytest.mark.gpu


def test_two_ranks_reproduce_the_single_process_run():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "probes", "rank_parity.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("RANK_PARITY")]
    assert res.returncode == 0 and line and line[0].endswith("OK"), (res.stdout[-2000:], res.stderr[-2000:])
    print(line[0])

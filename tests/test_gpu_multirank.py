"""GPU test of the N>1 path (needs >= 2 GPUs on the box; skipped otherwise): spawns torchrun on
tests/multigpu_check.py, which compares the NCCL-sharded library run with the oracle."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_two_gpus_match_oracle():
    import ldagroupedgibbssampler_b200 as L
    n = L.load().ldagpu_device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "multigpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "multigpu_check ok" in r.stdout

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


def test_one_process_two_gpus_match_oracle():
    """ldagpu_create_multi: one caller thread (standing in for the reference's one JVM) drives two GPUs and
    reproduces the oracle -- and so the one-process-per-GPU run -- bit for bit."""
    import ldagroupedgibbssampler_b200 as L
    if L.load().ldagpu_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, os.path.join(ROOT, "tests", "singleproc_multigpu_check.py"), "--gpus", "2"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "singleproc_multigpu_check ok" in r.stdout


def test_one_process_one_gpu_through_create_multi():
    """n_devices = 1 is the plain single-GPU handle behind the same entry point."""
    cmd = [sys.executable, os.path.join(ROOT, "tests", "singleproc_multigpu_check.py"), "--gpus", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_stalled_shard_times_out_instead_of_hanging():
    """Exchange-protocol fault injection: one shard never publishes its Phi rows; the others' bounded waits
    (LDAGPU_P2P_TIMEOUT_MS) must surface as the library's error within seconds."""
    import ldagroupedgibbssampler_b200 as L
    if L.load().ldagpu_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, LDAGPU_FAULT_STALL_SHARD="1", LDAGPU_P2P_TIMEOUT_MS="300")
    cmd = [sys.executable, os.path.join(ROOT, "tests", "singleproc_multigpu_check.py"), "--gpus", "2", "--stalled-shard"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and "stalled_shard ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]

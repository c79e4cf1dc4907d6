"""Multi-GPU checks that need >= 2 visible B200s (skipped on a 1-GPU box): one stereo pair split by rows over 2 ranks
(scenedepthestimation_b200/sharded.py, mccnn_sgm_sharded: NVLink peer-memory hand-over of the SGM path state inside the scan
kernels) must equal the single-GPU path bit for bit, and a rank that reports bad inputs must make every rank raise instead of
leaving kernels waiting."""
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(nproc, *args, timeout=600):
    port = 29600 + os.getpid() % 1500
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "run_sharded.py"), *args]
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)


def _need(n):
    if not torch.cuda.is_available() or torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")


@pytest.mark.parametrize("cfg", ["c1", "c5"])
def test_two_rank_sharded_pair_is_bit_identical(cfg):
    _need(2)
    r = _torchrun(2, cfg)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "bit-identical to single GPU: True" in r.stdout


@pytest.mark.parametrize("cfg", ["c1", "c5"])
def test_two_rank_sharded_pair_fused_mode_equals_single_gpu_fused(cfg):
    """mccnn_sgm_fused_sharded over real NVLink peer memory: the row sweeps' step-by-step FIFOs between the ranks, the column /
    diagonal hand-overs. Value for value the single-GPU fused mode."""
    _need(2)
    r = _torchrun(2, cfg, "fused")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "fused: bit-identical to single GPU: True" in r.stdout


def test_two_rank_fault_fused_mode():
    _need(2)
    r = _torchrun(2, "c1", "fused", "fault", timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("abandoned pair raised") == 2 and "bit-identical to single GPU: True" in r.stdout


def test_two_rank_fault_is_loud_not_a_hang():
    _need(2)
    r = _torchrun(2, "c1", "fault", timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("abandoned pair raised") == 2 and "bit-identical to single GPU: True" in r.stdout

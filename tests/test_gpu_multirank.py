"""-m gpu, needs >= 2 GPUs (skipped otherwise): block-sharded run over NCCL is bit-identical to
the single-process oracle (sharding must not change bits)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("overlap", [0, 1])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_run_matches_oracle(world, overlap):
    """overlap=1: the NCCL exchange runs behind the stage kernel (pyh_stage_overlapped, dispatch table + epoch wait)."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world + 16 * overlap), os.path.join(ROOT, "tests", "multirank_worker.py")]
    env = dict(os.environ, PYH_HALO_OVERLAP=str(overlap))
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert f"MULTIRANK OK world={world}" in out.stdout


@pytest.mark.parametrize("overlap", [0, 1])
def test_rank_boundary_on_east_west_edge(overlap):
    """2 blocks side by side on 2 ranks: the remote strips are columns (the 2x4 layout above only ever puts rank
    boundaries between block rows for 2 and 4 ranks)."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(29560 + overlap), os.path.join(ROOT, "tests", "multirank_worker.py")]
    env = dict(os.environ, PYH_HALO_OVERLAP=str(overlap), PYH_TEST_LAYOUT="2x1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MULTIRANK OK world=2" in out.stdout

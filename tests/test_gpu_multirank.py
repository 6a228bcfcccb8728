"""-m gpu, needs >= 2 GPUs (skipped otherwise): the block-sharded run -- Euler2D.solve() on N ranks, the strip exchange
and the dt all-reduce done by the library's own NCCL communicator inside the captured step graph (pyh_comm_init) -- is
bit-identical to the single-process oracle (sharding must not change bits)."""
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


def _run(world, port, **env):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "multirank_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, **env))
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert f"MULTIRANK OK world={world}" in out.stdout


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_run_matches_oracle(world):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    _run(world, 29500 + world)


def test_rank_boundary_on_east_west_edge():
    """2 blocks side by side on 2 ranks: the remote strips are columns (the 2x4 layout above only ever puts rank
    boundaries between block rows for 2 and 4 ranks)."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, 29560, PYH_TEST_LAYOUT="2x1")


def test_rank_boundaries_on_all_four_sides_small_blocks():
    """2 x 2 blocks on 4 ranks: every rank meets other ranks across an east / west AND a north / south edge.  26 x 22 blocks are
    too narrow to split off edge column strips, so the tile plan must fall back to one launch + exchange (the 8-rank run of the
    2 x 4 layout caught a plan that packed the east / west columns before the interior launch had written them)."""
    if _ngpu() < 4:
        pytest.skip("needs 4 GPUs")
    _run(4, 29570, PYH_TEST_LAYOUT="2x2", PYH_TEST_NORTH="10.0")


def test_rank_boundaries_on_all_four_sides_both_splits_active():
    """same layout with 90 x 40 blocks and 32-lane strips: 4 column strips, so the north / south edge rows AND the east / west
    edge column strips run ahead of the exchange and the interior overlaps it"""
    if _ngpu() < 4:
        pytest.skip("needs 4 GPUs")
    _run(4, 29572, PYH_TEST_LAYOUT="2x2", PYH_TEST_NORTH="10.0", PYH_TEST_NX="90", PYH_TEST_NY="40", PYH_MARCH_NT="32")


def test_single_stage_tableau_sharded():
    """ExplicitEuler1 alternates two state buffers: the exchange must follow the buffer the stage wrote."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, 29562, PYH_TEST_INTEGRATOR="ExplicitEuler1")


def test_step_by_step_driving_is_collective_and_bit_identical():
    """Euler2D.step() (get_dt -> integrate -> realizability check per call, two host syncs per step) instead of the device loop"""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, 29566, PYH_TEST_MODE="step")


def test_blocking_exchange_switch():
    """PYH_NO_HALO_OVERLAP=1: one launch per stage, exchange behind it (the diagnostic order) gives the same bits"""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, 29568, PYH_NO_HALO_OVERLAP="1")


def test_more_ranks_than_blocks():
    """2 blocks on 4 ranks: ranks 2 and 3 own nothing and only take part in the reductions (the reference tolerates idle
    ranks: `if len(self._blocks)` in Euler2D._solve, np.inf in get_dt)."""
    if _ngpu() < 4:
        pytest.skip("needs 4 GPUs")
    _run(4, 29564, PYH_TEST_LAYOUT="2x1")

"""-m gpu: BASELINE.json's full block size (2048 cells per side).

* the full width / full height of a block against the oracle on thin blocks (2048 x 6 and 6 x 2048, two blocks
  with an inter-block edge each): every column strip and every row strip of the stage kernel, bit for bit;
* the headline workload itself (8 blocks of 2048 x 2048, Roe + Venkatakrishnan, RK4) through size-independent
  properties: realizability, run-to-run bitwise determinism, conservation of mass and energy in the closed
  (all-reflecting) domain to rounding level, and the finite speed of propagation (2 cells per stage)."""
import numpy as np
import pytest

import cases
from test_gpu_parity import compare

pytestmark = pytest.mark.gpu


def test_full_width_row_of_strips_vs_oracle():
    blocks = cases.em_mesh(nbx=2, nby=1, east=10.0, north=0.03)
    # box in x (3 <= x <= 7) times a step in y: shocks cross the column strips, every row differs
    compare(blocks, 2048, 6, lambda x, y: cases.explosion_ic(x, np.full_like(x, 5.0)) * np.where(y[..., None] > 0.015, 1.0, 1.25), 2)


def test_full_height_column_of_strips_vs_oracle():
    blocks = cases.em_mesh(nbx=1, nby=2, east=0.03, north=10.0)
    compare(blocks, 6, 2048, lambda x, y: cases.explosion_ic(np.full_like(y, 5.0), y) * np.where(x[..., None] > 0.015, 1.0, 1.25), 2)


def test_headline_workload_size_independent_properties():
    import bench

    n, nb = 2048, 8
    blocks = bench.ws_mesh(1, nb)
    width, height = bench.BLOCK_LEN * nb, bench.BLOCK_LEN

    def ic(x, y):
        return bench.ws_ic(x, y, width, height)

    eng = cases.build_engine(blocks, n, n, ic)
    try:
        area = {g: eng.meshes[g].area for g in blocks}
        U0 = {g: eng.download(g) for g in blocks}

        def totals(U):
            return np.array([sum(float((U[g][..., k] * area[g]).sum()) for g in blocks) for k in (0, 3)])

        t0 = totals(U0)
        nsteps = 3
        t, done, bad, dts = eng.run(0.0, 1e9, max_steps=nsteps, record_dts=nsteps)
        assert done == nsteps and not bad and eng.realizable()
        U1 = {g: eng.download(g) for g in blocks}
        # conservation: interior face fluxes cancel pairwise, wall fluxes of mass / energy vanish by symmetry of the mirror state
        t1 = totals(U1)
        assert np.all(np.abs(t1 - t0) <= 1e-11 * np.abs(t0)), (t0, t1)
        # nothing travels more than 2 cells per stage: away from the box the state is still the initial one (to rounding:
        # a uniform state at rest has a residual of a few ulps on the non-uniform mesh), at the box edge it is not
        reach = 2 * 4 * nsteps + 2
        assert np.abs(U1[0] - U0[0]).max() <= 1e-12          # westernmost block; the box starts in block 2
        gm = 2   # holds the west edge of the box at x = 0.3 * width
        jedge = int(round((0.3 * width - gm * bench.BLOCK_LEN) / bench.BLOCK_LEN * n))
        assert np.abs(U1[gm][:, : jedge - reach] - U0[gm][:, : jedge - reach]).max() <= 1e-12
        mid = n // 2
        assert np.abs(U1[gm][mid, jedge - 2: jedge + 2] - U0[gm][mid, jedge - 2: jedge + 2]).max() > 1e-3
        # determinism: same inputs, same bits
        for gid in blocks:
            eng.upload(gid, U0[gid])
        eng.apply_bc()
        t2, done2, bad2, dts2 = eng.run(0.0, 1e9, max_steps=nsteps, record_dts=nsteps)
        assert list(dts2) == list(dts)
        for gid in blocks:
            assert np.array_equal(eng.download(gid), U1[gid]), gid
    finally:
        eng.close()

"""-m gpu: device-side two-state initial conditions (pyh_fill_box).  The fill kernel compares the
device's centroid planes with the box bounds; the result must equal, bit for bit, the numpy fill of the reference's examples
(np.where over block.mesh.x / .y, examples/explosion_multi/initial_condition.py:53-59, examples/dmr/initial_condition.py:55-58)
uploaded from the host -- on rectangular and on skewed (DMR ramp) blocks, with +-inf bounds and with `outside = None`."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


def _engine(blocks, nx, ny, **kw):
    return cases.build_engine(blocks, nx, ny, lambda x, y: np.zeros(x.shape + (4,)) + np.array([1.0, 0.0, 0.0, 2.5]), **kw)


def test_explosion_box_equals_numpy_fill():
    blocks = cases.em_mesh()
    eng = _engine(blocks, 37, 29)
    try:
        UL = cases.prim_to_cons_nd(np.array([4.6968, 0.0, 0.0, 404400.0]).reshape(1, 1, 4)).reshape(4)
        UR = cases.prim_to_cons_nd(np.array([1.1742, 0.0, 0.0, 101100.0]).reshape(1, 1, 4)).reshape(4)
        for g in blocks:
            eng.fill_box(g, 3.0, 7.0, 3.0, 7.0, UL, UR)
        for g in blocks:
            m = eng.meshes[g]
            assert np.array_equal(eng.download(g), cases.explosion_ic(m.x[:, :, 0], m.y[:, :, 0])), g
    finally:
        eng.close()


def test_half_plane_on_skewed_blocks_and_keep_outside():
    blocks = cases.dmr_mesh()
    eng = _engine(blocks, 33, 21, flux="HLLL", recon="primitive", integrator="RK2", CFL=0.4)
    try:
        UL = cases.prim_to_cons_nd(np.array([8.0, 8.25, 0.0, 116.5]).reshape(1, 1, 4)).reshape(4)
        UR = cases.prim_to_cons_nd(np.array([1.4, 0.0, 0.0, 1.0]).reshape(1, 1, 4)).reshape(4)
        for g in blocks:
            eng.fill_box(g, -np.inf, 0.95, -np.inf, np.inf, UL, UR)
        for g in blocks:
            m = eng.meshes[g]
            assert np.array_equal(eng.download(g), cases.dmr_ic(m.x[:, :, 0], m.y[:, :, 0])), g
        # a second box on top, leaving the rest untouched
        mark = np.array([2.0, 0.5, -0.25, 9.0])
        for g in blocks:
            eng.fill_box(g, 1.0, 2.5, 0.3, 0.9, mark, None)
        for g in blocks:
            m = eng.meshes[g]
            x, y = m.x[:, :, 0], m.y[:, :, 0]
            ref = cases.dmr_ic(x, y)
            inside = (x >= 1.0) & (x <= 2.5) & (y >= 0.3) & (y <= 0.9)
            ref = np.where(inside[..., None], mark, ref)
            assert np.array_equal(eng.download(g), ref), g
    finally:
        eng.close()


def test_box_initial_condition_through_the_facade_matches_the_numpy_one():
    """Euler2D with BoxInitialCondition (device fill, no state upload) == Euler2D with the reference-style numpy IC."""
    from test_gpu_facade import ExplosionInitialCondition, em_config, em_mesh

    from pyhype_b200.initial_conditions import BoxInitialCondition
    from pyhype_b200.solvers import Euler2D

    box = BoxInitialCondition(inside=(4.6968, 0.0, 0.0, 404400.0), outside=(1.1742, 0.0, 0.0, 101100.0), x0=3, x1=7, y0=3, y1=7)
    sims = []
    for ic in (ExplosionInitialCondition(), box):
        sim = Euler2D(config=em_config(nx=28, ny=24, t_final=0.002, initial_condition=ic), mesh_config=em_mesh())
        sim.solve()
        sims.append(sim)
    a, b = sims
    assert a.num_time_step == b.num_time_step and a.t == b.t
    for ba, bb in zip(a.blocks, b.blocks):
        assert np.array_equal(ba.state.data, bb.state.data)


def test_box_initial_condition_state_is_readable_before_solve():
    """host code may read block.state.data between apply_initial_condition and solve: the pending fill materialises on the host"""
    from test_gpu_facade import em_config, em_mesh

    from pyhype_b200.initial_conditions import BoxInitialCondition
    from pyhype_b200.solvers import Euler2D

    box = BoxInitialCondition(inside=(4.6968, 0.0, 0.0, 404400.0), outside=(1.1742, 0.0, 0.0, 101100.0), x0=3, x1=7, y0=3, y1=7)
    sim = Euler2D(config=em_config(nx=20, ny=16, t_final=0.001, initial_condition=box), mesh_config=em_mesh())
    sim.apply_initial_condition()
    for blk in sim.blocks:
        ref = cases.explosion_ic(blk.mesh.x[:, :, 0], blk.mesh.y[:, :, 0])
        assert np.array_equal(blk.state.data, ref)
    sim.apply_boundary_condition()
    for blk in sim.blocks:
        ref = cases.explosion_ic(blk.mesh.x[:, :, 0], blk.mesh.y[:, :, 0])
        assert np.array_equal(blk.state.data, ref)


def test_weak_scaling_box_equals_the_numpy_initial_condition():
    """bench.py fills its explosion box on the device (no 1 GiB upload per GPU): the same states and bounds must give what
    pyhype_b200.examples.ws_ic evaluates on the host, block by block, on a 3 x 2 grid of rectangular blocks."""
    blocks = cases.ws_mesh(2, 3)
    width, height = 3 * 1.25, 2 * 1.25
    eng = _engine(blocks, 40, 36)
    try:
        hi = cases.ws_ic(np.array([[0.5 * width]]), np.array([[0.5 * height]]), width, height)[0, 0]
        lo = cases.ws_ic(np.array([[0.0]]), np.array([[0.0]]), width, height)[0, 0]
        for g in blocks:
            eng.fill_box(g, 0.3 * width, 0.7 * width, 0.3 * height, 0.7 * height, hi, lo)
        n_in = 0
        for g in blocks:
            m = eng.meshes[g]
            ref = cases.ws_ic(m.x[:, :, 0], m.y[:, :, 0], width, height)
            got = eng.download(g)
            assert np.array_equal(got, ref), g
            n_in += int((got[..., 0] == hi[0]).sum())
        assert 0 < n_in < 6 * 40 * 36
    finally:
        eng.close()

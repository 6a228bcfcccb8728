"""Pins the oracle against the LIVE reference when /root/reference is present (build container
only; skipped on the GPU box, where the committed fixtures under tests/golden/ take over)."""
import numpy as np
import pytest

import cases
from oracle import muscl_oracle as mo
from oracle import refharness as rh

pytestmark = pytest.mark.skipif(not rh.available(), reason="reference tree not present")


def _compare(blocks, nx, ny, ic, nsteps, **cfg):
    from make_golden import ref_blocks  # noqa: F401  (same adaptor the fixtures were made with)

    class IC:
        def apply_to_block(self, block):
            block.state.data = np.ascontiguousarray(ic(block.mesh.x[:, :, 0], block.mesh.y[:, :, 0]))

    config = rh.make_config(nx=nx, ny=ny, initial_condition=IC(), **cfg)
    run = rh.RefRun(config, ref_blocks(blocks))
    recon = cfg.get("reconstruction_type", "conservative")
    prob = mo.Problem(blocks, nx, ny, flux=config.fvm_flux_function_type, limiter=config.fvm_slope_limiter_type,
                      recon=recon, integrator=config.time_integrator, CFL=config.CFL, nqp=config.fvm_num_quadrature_points)
    for b, rb in zip(prob.blocks.values(), run.blocks):
        b.U = rb.state.data.copy()
    prob.apply_bc()
    for b, r in zip(prob.blocks.values(), run.residuals()):
        prob.residual(b, keep=True)
        for k in ("gx", "gy", "phi", "FE", "FW", "FN", "FS", "R"):
            assert np.array_equal(b.dbg[k], r[k]), (b.gid, k)
    run.step(nsteps)
    _, dts = prob.run(0.0, config.t_final * 343.0, max_steps=nsteps)
    assert dts == run.dts
    for b, rb, gh in zip(prob.blocks.values(), run.blocks, run.ghosts()):
        assert np.array_equal(b.U, rb.state.data)
        for s in ("E", "W", "N", "S"):
            assert np.array_equal(b.ghost[s], gh[s])


@pytest.fixture(scope="module", autouse=True)
def _path():
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    rh.activate()
    yield


def test_live_explosion_multi_roe_rk4():
    _compare(cases.em_mesh(), 20, 20, cases.explosion_ic, 6)


def test_live_dmr_hlll_primitive_rk2():
    _compare(cases.dmr_mesh(), 20, 20, cases.dmr_ic, 8, fvm_flux_function_type="HLLL", time_integrator="RK2",
             CFL=0.4, reconstruction_type="primitive")


def test_live_wedge_dirichlet():
    _compare(cases.wedge_mesh(12), 14, 12, cases.wedge_ic, 8, fvm_flux_function_type="HLLL", time_integrator="RK2",
             CFL=0.3, reconstruction_type="primitive")


def test_live_nrm2_matches_numba_linalg_norm():
    import numba as nb

    @nb.njit
    def norms(v):
        out = np.zeros(v.shape[0])
        for i in range(v.shape[0]):
            out[i] = np.linalg.norm(v[i])
        return out

    rng = np.random.default_rng(3)
    v = rng.standard_normal((50000, 4)) * 10.0 ** rng.integers(-8, 8, size=(50000, 1))
    assert np.array_equal(norms(v), mo.nrm2_x87(v))


def test_live_shockbox_nan_of_the_limiter_and_the_abort():
    """examples/shockbox as shipped: in step 24 the Venkatakrishnan quotient of an overflowing slope is inf / inf and
    np.minimum.reduce carries the NaN into phi (limiters/base.py:179-186).  The oracle must put its NaNs where the reference
    does -- same cells, same step -- and both must refuse the state (Euler2D.py:144-152); the kernels are held to the oracle's
    NaN cells by tests/test_named_configs.py and tests/test_host_twin.py."""
    from make_golden import ref_blocks

    blocks, ic = cases.em_mesh(1, 1, 10.0, 10.0), cases.shockbox_ic

    class IC:
        def apply_to_block(self, block):
            block.state.data = np.ascontiguousarray(ic(block.mesh.x[:, :, 0], block.mesh.y[:, :, 0]))

    config = rh.make_config(nx=50, ny=50, initial_condition=IC(), time_integrator="RK2", CFL=0.4, t_final=2.0)
    run = rh.RefRun(config, ref_blocks(blocks))
    prob = cases.build_oracle(blocks, 50, 50, ic, flux="Roe", limiter="Venkatakrishnan", recon="conservative", integrator="RK2", CFL=0.4)
    run.step(23)
    t, dts = prob.run(0.0, 1e9, max_steps=23)
    assert dts == run.dts
    g = sorted(prob.blocks)[0]
    assert np.array_equal(prob.blocks[g].U, run.blocks[0].state.data) and not np.isnan(prob.blocks[g].U).any()
    s = run.solver
    dt = s.get_dt()
    with np.errstate(all="ignore"):
        s._update_solution_blocks(dt=dt)                  # step 24 without the reference's own check ...
        prob.step(prob.get_dt(t, 1e9))
    Uref = run.blocks[0].state.data
    assert np.isnan(Uref).sum() == 16
    assert np.array_equal(prob.blocks[g].U, Uref, equal_nan=True)
    assert not prob.realizable()
    with pytest.raises(SystemExit):                        # ... which aborts (the stub MPI's Abort)
        s._realizability_check()


@pytest.mark.parametrize("flux", ["Roe", "HLLL", "HLLE"])
def test_live_flux_functions_on_unrealizable_face_states_including_nan(flux):
    """The reference's own flux objects on face states with negative pressure / density on one side: the oracle's fluxes -- the
    yardstick the kernels' NaN behaviour is tested against (tests/test_host_twin.py) -- are the reference's, NaN for NaN."""
    import test_host_twin as T
    from pyhype.flux.HLLE import FluxHLLE
    from pyhype.flux.HLLL import FluxHLLL
    from pyhype.flux.Roe import FluxRoe
    from pyhype.states.primitive import PrimitiveState

    if flux == "HLLE":
        rh.patch_hlle()       # never runs unpatched
    cfg = rh.make_config(nx=40, ny=30)
    WL, WR = T.face_states(1200, seed=61)
    rng = np.random.default_rng(62)
    for Wp, lo in ((WL, 0), (WR, 1)):
        sel = rng.integers(0, 6, len(Wp))
        Wp[sel == lo, 3] *= -1.0
        Wp[sel == 2 + lo, 0] *= -1.0
    WL, WR = WL.reshape(30, 40, 4), WR.reshape(30, 40, 4)
    f = FluxRoe(cfg, 39, 30) if flux == "Roe" else {"HLLL": FluxHLLL, "HLLE": FluxHLLE}[flux](cfg, nx=40, ny=30)
    with np.errstate(all="ignore"):
        ref = f(PrimitiveState(fluid=cfg.fluid, array=WL.copy()), PrimitiveState(fluid=cfg.fluid, array=WR.copy()))
        mine = mo.FLUXES[flux](WL.copy(), WR.copy(), cases.GAMMA)
    assert np.isnan(ref).any(axis=-1).mean() > 0.3
    assert np.array_equal(mine, ref, equal_nan=True)


def test_live_limiter_functions_including_overflowing_slopes():
    """limiters/limiters.py:25-62 on slopes from 0 to 1e300 (Venkatakrishnan through its numba loop, the variant every example
    runs): the oracle's four limiter functions are the reference's, overflow to inf / inf = NaN included."""
    from pyhype.limiters.limiters import BarthJespersen, VanAlbada, VanLeer, Venkatakrishnan

    rng = np.random.default_rng(7)
    s = np.concatenate((np.zeros(8), rng.uniform(0, 6, 4000), 10.0 ** rng.uniform(-300, 300, 4000), [1.0, 2.0, 1e154, 1.4180058467152723e211, np.inf]))
    s = s.reshape(-1, 1, 1)
    with np.errstate(all="ignore"):
        ref = {"Venkatakrishnan": Venkatakrishnan._venkata(s.copy()), "VanAlbada": VanAlbada._limiter_func(s.copy()),
               "VanLeer": VanLeer._limiter_func(s.copy()), "BarthJespersen": BarthJespersen._limiter_func(s.copy())}
        for name, r in ref.items():
            assert np.array_equal(mo.LIMITERS[name](s.copy()), r, equal_nan=True), name
    assert np.isnan(ref["Venkatakrishnan"]).sum() > 500 and np.isnan(ref["VanAlbada"]).sum() > 500

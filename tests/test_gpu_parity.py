"""-m gpu parity tests: CUDA engine (through the C ABI) vs the oracle on the same inputs.
Bar: equal by value (max-abs-diff == 0) -- the reference's limiter makes the scheme sensitive to
single-ulp changes (SURVEY.md finding 3), so bitwise agreement is the meaningful target; the
north-star tolerance (1e-10 relative L-inf per conserved variable) is asserted as well."""
import numpy as np
import pytest

import cases

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("stage_path")]   # every test once per stage implementation (conftest.py)


def rel_linf(a, b):
    return max(np.abs(a[..., k] - b[..., k]).max() / max(np.abs(b[..., k]).max(), 1e-300) for k in range(4))


def check_equal(name, got, ref):
    if not np.array_equal(got, ref):
        d = np.abs(got - ref)
        idx = np.unravel_index(np.argmax(d), d.shape)
        raise AssertionError(f"{name}: max abs diff {d.max():.3e} at {idx}; rel Linf {rel_linf(got, ref):.3e}; "
                             f"{np.count_nonzero(d)} of {d.size} values differ")


def compare(blocks, nx, ny, ic, nsteps, **kw):
    prob = cases.build_oracle(blocks, nx, ny, ic, **kw)
    eng = cases.build_engine(blocks, nx, ny, ic, **kw)
    try:
        for gid, b in prob.blocks.items():
            for s in cases.SIDES:
                check_equal(f"ghost0 blk{gid} {s}", eng.download_ghost(gid, s), b.ghost[s])
        for gid, b in prob.blocks.items():
            prob.residual(b, keep=True)
            check_equal(f"gx blk{gid}", eng.debug_fetch(gid, "gx"), b.dbg["gx"])
            check_equal(f"gy blk{gid}", eng.debug_fetch(gid, "gy"), b.dbg["gy"])
            check_equal(f"phi blk{gid}", eng.debug_fetch(gid, "phi"), b.dbg["phi"])
            check_equal(f"R blk{gid}", eng.residual(gid), b.dbg["R"])
        t_final = 1e9
        t, dts = prob.run(0.0, t_final, max_steps=nsteps)
        tg, n, bad, dtg = eng.run(0.0, t_final, max_steps=nsteps, poll_every=7, record_dts=nsteps)
        assert n == nsteps and not bad
        assert list(dtg) == dts, (list(dtg), dts)
        assert tg == t
        for gid, b in prob.blocks.items():
            U = eng.download(gid)
            check_equal(f"state blk{gid} after {nsteps} steps", U, b.U)
            assert rel_linf(U, b.U) <= 1e-10
            for s in cases.SIDES:
                check_equal(f"ghost blk{gid} {s}", eng.download_ghost(gid, s), b.ghost[s])
    finally:
        eng.close()


def test_explosion_multi_roe_rk4():
    compare(cases.em_mesh(), 40, 40, cases.explosion_ic, 12)


def test_explosion_multi_ragged_tiles():
    compare(cases.em_mesh(), 37, 21, cases.explosion_ic, 6)


def test_explosion_single_cartesian_block():
    blocks = cases.em_mesh(nbx=1, nby=1)
    compare(blocks, 40, 80, cases.explosion_ic, 6)


def test_dmr_roe_conservative_rk2():
    compare(cases.dmr_mesh(), 30, 30, cases.dmr_ic, 10, integrator="RK2", CFL=0.4)


def test_dmr_roe_primitive_rk2():
    compare(cases.dmr_mesh(), 30, 30, cases.dmr_ic, 10, integrator="RK2", CFL=0.4, recon="primitive")


def test_dmr_hlll_primitive_rk2():
    compare(cases.dmr_mesh(), 50, 50, cases.dmr_ic, 20, flux="HLLL", integrator="RK2", CFL=0.4, recon="primitive")


def test_dmr_hlle_primitive_rk2():
    compare(cases.dmr_mesh(), 24, 24, cases.dmr_ic, 10, flux="HLLE", integrator="RK2", CFL=0.4, recon="primitive")


def test_em_hlll_conservative():
    compare(cases.em_mesh(), 24, 24, cases.explosion_ic, 6, flux="HLLL", integrator="RK2")


@pytest.mark.parametrize("lim", ["VanLeer", "VanAlbada", "BarthJespersen"])
def test_limiters(lim):
    compare(cases.em_mesh(), 20, 20, cases.explosion_ic, 5, limiter=lim)


@pytest.mark.parametrize("integ", ["ExplicitEuler1", "RK2", "Ralston2", "RK3", "RK3SSP", "Ralston3", "Ralston4", "DormandPrince5"])
def test_integrators(integ):
    compare(cases.em_mesh(), 20, 20, cases.explosion_ic, 5, integrator=integ, CFL=0.3)


def test_wedge_dirichlet_hlll_primitive():
    compare(cases.wedge_mesh(30), 30, 30, cases.wedge_ic, 15, flux="HLLL", integrator="RK2", CFL=0.3, recon="primitive")


def test_wedge_dirichlet_roe_conservative():
    compare(cases.wedge_mesh(24), 30, 24, cases.wedge_ic, 15, flux="Roe", integrator="RK2", CFL=0.3)


def test_smooth_ic_weak_scaling_layout():
    blocks = cases.em_mesh(nbx=4, nby=2, east=5.0, north=2.5)
    compare(blocks, 48, 48, cases.smooth_ic, 8)


@pytest.mark.parametrize("nqp", [2, 3])
def test_quadrature_points_explosion(nqp):
    compare(cases.em_mesh(), 37, 21, cases.explosion_ic, 5, nqp=nqp)


@pytest.mark.parametrize("nqp", [2, 3])
def test_quadrature_points_dmr_hlll_primitive(nqp):
    # non-Cartesian blocks, Slipwall / OutletDirichlet edge states evaluated per quadrature point
    compare(cases.dmr_mesh(), 30, 30, cases.dmr_ic, 8, flux="HLLL", integrator="RK2", CFL=0.4, recon="primitive", nqp=nqp)


def test_quadrature_points_wedge_dirichlet_roe():
    compare(cases.wedge_mesh(24), 30, 24, cases.wedge_ic, 8, flux="Roe", integrator="RK2", CFL=0.3, nqp=2)


def test_ssp_rk2_heun_tableau_dmr():
    """BASELINE.json's DMR config names SSP-RK2; the reference factory only has the midpoint "RK2".  The
    tableau plumbing is generic: the Heun tableau (pyhype_b200 key "SSPRK2") against the oracle driven with
    the same coefficients."""
    from pyhype_b200.time_marching import get_tableau

    compare(cases.dmr_mesh(), 30, 30, cases.dmr_ic, 10, flux="HLLL", integrator=get_tableau("SSPRK2"), CFL=0.4, recon="primitive")


@pytest.mark.parametrize("kw", [dict(), dict(flux="HLLL", recon="primitive", integrator="RK2"), dict(flux="HLLE", integrator="RK2")],
                         ids=["roe_cons", "hlll_prim", "hlle_cons"])
def test_plain_operator_fallbacks_on_huge_magnitudes(kw):
    """Every fast division / square-root / norm sequence has a validity range ([2^-255, 2^257), [2^-127, 2^127) for
    the HLLL norm); outside it a phase is re-evaluated with the plain IEEE operators / the integer norm.  Scaling
    the explosion state by 2^300 (density, momentum and energy alike: same velocities, same dt) drives EVERY
    phase of EVERY cell through those fallbacks; the result must still equal the oracle's bit for bit.  (Scaling
    DOWN is not an option: the reference's Roe shear wave strength lacks its factor rho -- SURVEY.md 8A.6 -- so the
    scheme is not homogeneous in the density and blows up for small densities, in the reference as well.)"""
    s = 2.0 ** 300
    compare(cases.em_mesh(), 20, 20, lambda x, y: cases.explosion_ic(x, y) * s, 3, **kw)


@pytest.mark.parametrize("nx,ny", [(1, 1), (2, 1), (1, 3), (3, 2), (5, 4)])
def test_tiny_blocks(nx, ny):
    """blocks of a handful of cells: every cell touches two or more block edges, strips are almost all halo"""
    compare(cases.em_mesh(), nx, ny, cases.explosion_ic, 3)

"""-m gpu: the CUDA engine, called through the C ABI, against the committed golden fixtures
(outputs of the unmodified reference) -- no oracle in the loop.  Bit-exact bar; the north-star
tolerance of 1e-10 relative L-inf per conserved variable is asserted explicitly as well."""
import numpy as np
import pytest

import cases
import golden_io

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("stage_path")]   # every test once per stage implementation (conftest.py)

DEVICE_FIXTURES = golden_io.names()


def rel_linf(a, b):
    return max(np.abs(a[..., k] - b[..., k]).max() / max(np.abs(b[..., k]).max(), 1e-300) for k in range(4))


@pytest.mark.parametrize("name", DEVICE_FIXTURES)
def test_engine_matches_reference_fixture(name, stage_path):
    fx = golden_io.Fixture(name)
    states = {g: fx[f"U0_{g}"] for g in fx.gids}
    eng = cases.build_engine(fx.blocks, fx.nx, fx.ny, None, states=states, **fx.scheme())
    try:
        assert eng.stage_path() == (stage_path if fx.scheme()["nqp"] == 1 else "fused")
        for g in fx.gids:
            m = eng.meshes[g]
            assert np.array_equal(m.area, fx[f"A_{g}"]) and np.array_equal(m.x[:, :, 0], fx[f"xc_{g}"])
            assert np.array_equal(m.theta_v[:, 1:], fx[f"thetaE_{g}"]) and np.array_equal(m.theta_h[1:], fx[f"thetaN_{g}"])
            for s in golden_io.SIDES:
                assert np.array_equal(eng.download_ghost(g, s), fx[f"ghost0_{g}_{s}"]), (g, s)
            for k in ("gx", "gy", "phi"):
                assert np.array_equal(eng.debug_fetch(g, k), fx[f"{k}_{g}"]), (g, k)
            R = eng.residual(g)
            assert np.array_equal(R, fx[f"R_{g}"]), (g, np.abs(R - fx[f"R_{g}"]).max())
        n = fx.meta["steps"]
        t, done, bad, dts = eng.run(0.0, 1e9, max_steps=n, poll_every=3, record_dts=n)
        assert done == n and not bad
        assert list(dts) == list(fx["dts"])
        for g in fx.gids:
            U = eng.download(g)
            assert rel_linf(U, fx[f"U_{g}"]) <= 1e-10
            assert np.array_equal(U, fx[f"U_{g}"]), (g, np.abs(U - fx[f"U_{g}"]).max())
    finally:
        eng.close()


def test_stepwise_api_equals_device_resident_loop():
    fx = golden_io.Fixture("em_roe_venkat_cons_rk4")
    states = {g: fx[f"U0_{g}"] for g in fx.gids}
    eng = cases.build_engine(fx.blocks, fx.nx, fx.ny, None, states=states, **fx.scheme())
    try:
        t = 0.0
        for dt_ref in fx["dts"]:
            dt = eng.get_dt(t, 1e9)
            assert dt == dt_ref
            eng.step(dt)
            assert eng.realizable()
            t += dt
        for g in fx.gids:
            assert np.array_equal(eng.download(g), fx[f"U_{g}"])
    finally:
        eng.close()


def test_dt_clamp_and_stop_at_t_final():
    fx = golden_io.Fixture("em_roe_venkat_cons_rk4")
    states = {g: fx[f"U0_{g}"] for g in fx.gids}
    eng = cases.build_engine(fx.blocks, fx.nx, fx.ny, None, states=states, **fx.scheme())
    try:
        dts = list(fx["dts"])
        t_final = dts[0] + dts[1] + 0.25 * dts[2]
        t, done, bad, got = eng.run(0.0, t_final, max_steps=-1, poll_every=2, record_dts=16)
        assert done == 3 and not bad
        assert list(got[:2]) == dts[:2]
        tt = 0.0 + dts[0]
        tt += dts[1]
        assert got[2] == t_final - tt          # solvers/base.py:132-136
        assert t == tt + got[2] and not (t < t_final)
    finally:
        eng.close()


def test_unrealizable_state_is_reported():
    fx = golden_io.Fixture("em_roe_venkat_cons_rk4")
    states = {g: fx[f"U0_{g}"].copy() for g in fx.gids}
    states[fx.gids[0]][3, 4, 0] = -1.0
    eng = cases.build_engine(fx.blocks, fx.nx, fx.ny, None, states=states, **fx.scheme())
    try:
        assert not eng.realizable()
        t, done, bad, _ = eng.run(0.0, 1e9, max_steps=4)
        assert bad
    finally:
        eng.close()


def test_error_paths_raise_like_the_reference():
    from pyhype_b200.engine import Engine
    from pyhype_b200.time_marching import TABLEAUX

    with pytest.raises(ValueError):
        Engine(8, 8, "AUSM", "Venkatakrishnan", "conservative", TABLEAUX["RK2"], 1.4, 0.5)
    with pytest.raises(ValueError):
        Engine(8, 8, "Roe", "Superbee", "conservative", TABLEAUX["RK2"], 1.4, 0.5)
    with pytest.raises(ValueError):
        Engine(8, 8, "Roe", "Venkatakrishnan", "conservative", TABLEAUX["RK2"], 1.4, 0.5, num_quadrature_points=4)
    blocks = cases.em_mesh()
    blocks[0]["BCTypeW"] = "Periodic"
    with pytest.raises(ValueError, match="has not been specialized"):
        cases.build_engine(blocks, 8, 8, cases.explosion_ic)

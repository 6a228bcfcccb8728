"""-m gpu: which kernels run a stage is decided by measurement (pyh_api.cu: tune_stage_path) -- the decision must not show in the
results.  The forced paths are covered by the `stage_path` fixture of the other GPU modules; here the DEFAULT: no PYH_SPLIT in the
environment, the context times both implementations at its first run() and keeps the faster."""
import numpy as np
import pytest

import cases
import golden_io

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["em_roe_venkat_cons_rk4", "dmr_hlll_venkat_prim_rk2", "em_int_ExplicitEuler1"])
def test_measured_stage_path_reproduces_the_reference_fixture(name, monkeypatch):
    monkeypatch.delenv("PYH_SPLIT", raising=False)
    fx = golden_io.Fixture(name)
    states = {g: fx[f"U0_{g}"] for g in fx.gids}
    eng = cases.build_engine(fx.blocks, fx.nx, fx.ny, None, states=states, **fx.scheme())
    try:
        assert eng.stage_path() == "fused" and eng.stage_path_tuning() == (0.0, 0.0)      # nothing measured before the first run
        n = fx.meta["steps"]
        t, done, bad, dts = eng.run(0.0, 1e9, max_steps=n, poll_every=3, record_dts=n)
        fused_ms, split_ms = eng.stage_path_tuning()
        assert fused_ms > 0.0 and split_ms > 0.0
        assert eng.stage_path() == ("split" if split_ms < 0.97 * fused_ms else "fused")
        assert done == n and not bad and list(dts) == list(fx["dts"])
        for g in fx.gids:
            assert np.array_equal(eng.download(g), fx[f"U_{g}"]), g
        # the measurement ran stage 0 a dozen times into scratch buffers: a second run() continues from the same state
        before = {g: eng.download(g) for g in fx.gids}
        assert eng.run(t, t, max_steps=4)[1] == 0
        for g in fx.gids:
            assert np.array_equal(eng.download(g), before[g])
    finally:
        eng.close()


def test_contexts_outside_the_split_stage_are_not_measured(monkeypatch):
    monkeypatch.delenv("PYH_SPLIT", raising=False)
    fx = golden_io.Fixture("em_nqp2")           # two quadrature points: fused kernel only
    states = {g: fx[f"U0_{g}"] for g in fx.gids}
    eng = cases.build_engine(fx.blocks, fx.nx, fx.ny, None, states=states, **fx.scheme())
    try:
        n = fx.meta["steps"]
        t, done, bad, dts = eng.run(0.0, 1e9, max_steps=n, record_dts=n)
        assert eng.stage_path() == "fused" and eng.stage_path_tuning() == (0.0, 0.0)
        assert done == n and list(dts) == list(fx["dts"])
        for g in fx.gids:
            assert np.array_equal(eng.download(g), fx[f"U_{g}"]), g
    finally:
        eng.close()

"""-m gpu: single-stage tableaux (ExplicitEuler1) alternate two state buffers inside ``pyh_run``.
The host flips the buffer roles once per ENQUEUED step while the device executes only the steps
up to ``t_final`` and replays a CUDA graph captured with fixed roles, so the cases below pin the
two ways the roles could drift apart (round-1 ADVICE, pyh_api.cu do_stage / pyh_run):

* ``t_final`` reached after an ODD number of executed steps with ``max_steps=-1`` (the host always
  enqueues an even number of steps);
* repeated ``run()`` calls with odd ``max_steps`` (1, then 5, then 5 ...), i.e. the cached graph
  replayed from the other buffer parity (``write_every_n_timesteps`` does exactly this).

Checked against the oracle by value (explicit_runge_kutta.py:47-89 with the one-row tableau)."""
import numpy as np
import pytest

import cases

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("stage_path")]   # every test once per stage implementation (conftest.py)

KW = dict(integrator="ExplicitEuler1", CFL=0.3)


def _pair(nx=24, ny=20):
    blocks = cases.em_mesh()
    prob = cases.build_oracle(blocks, nx, ny, cases.explosion_ic, **KW)
    eng = cases.build_engine(blocks, nx, ny, cases.explosion_ic, **KW)
    return prob, eng


def _check_states(prob, eng, what):
    for gid, b in prob.blocks.items():
        U = eng.download(gid)
        assert np.array_equal(U, b.U), f"{what}: block {gid} differs (max abs {np.abs(U - b.U).max():.3e})"


@pytest.mark.parametrize("nsteps", [5, 6, 7, 51])
def test_euler1_t_final_after_n_steps(nsteps):
    """t_final chosen so that the run ends after exactly nsteps steps (odd and even), max_steps = -1."""
    prob, eng = _pair()
    try:
        # dt sequence of a second oracle instance, then t_final in the middle of step nsteps
        ref = cases.build_oracle(cases.em_mesh(), 24, 20, cases.explosion_ic, **KW)
        _, dts = ref.run(0.0, 1e9, max_steps=nsteps)
        t_final = sum(dts[:-1]) + 0.5 * dts[-1]
        t_ref, dts_ref = prob.run(0.0, t_final)
        assert len(dts_ref) == nsteps
        tg, n, bad, dtg = eng.run(0.0, t_final, max_steps=-1, poll_every=50, record_dts=nsteps + 4)
        assert n == nsteps and not bad
        assert list(dtg) == dts_ref
        assert tg == t_ref
        _check_states(prob, eng, f"ExplicitEuler1 to t_final in {nsteps} steps")
        assert eng.realizable()
    finally:
        eng.close()


def test_euler1_repeated_runs_with_odd_step_counts():
    """run(1), run(5), run(5), run(2), run(3): the cached graph must follow the buffer parity."""
    prob, eng = _pair()
    try:
        t = tg = 0.0
        total = 0
        for k in (1, 5, 5, 2, 3, 4, 1):
            t, dts = prob.run(t, 1e9, max_steps=k)
            tg, n, bad, dtg = eng.run(tg, 1e9, max_steps=k, poll_every=3, record_dts=k)
            total += k
            assert n == k and not bad, (k, n, bad)
            assert list(dtg) == dts, (k, list(dtg), dts)
            assert tg == t
            _check_states(prob, eng, f"after {total} ExplicitEuler1 steps (call of {k})")
    finally:
        eng.close()


def test_euler1_step_then_run_then_step():
    """eager single steps (pyh_step) interleaved with pyh_run calls"""
    prob, eng = _pair()
    try:
        t = 0.0
        for k in (3, 4):
            dt = eng.get_dt(t, 1e9)
            t1, dts = prob.run(t, 1e9, max_steps=1)
            assert dts == [dt]
            eng.step(dt)
            t2, dts = prob.run(t1, 1e9, max_steps=k)
            tg, n, bad, dtg = eng.run(t1, 1e9, max_steps=k, record_dts=k)
            assert n == k and not bad and list(dtg) == dts and tg == t2
            t = t2
            _check_states(prob, eng, f"step + run({k})")
    finally:
        eng.close()

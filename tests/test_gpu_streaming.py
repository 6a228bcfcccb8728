"""-m gpu: the asynchronous state-streaming entry points (pyh_upload_state_async, pyh_commit_uploads,
pyh_download_state_async, pyh_transfers_sync) give exactly what the blocking ones give, also when
several independent jobs are pipelined back to back through one context (bench.py's e2e leg)."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


def _run_sync(eng, states, nsteps):
    for gid, U in states.items():
        eng.upload(gid, U)
    eng.apply_bc()
    eng.run(0.0, 1e9, max_steps=nsteps)
    return {gid: eng.download(gid) for gid in states}


def test_async_matches_blocking_and_pipelines():
    import torch

    nx = ny = 40
    blocks = cases.em_mesh()
    eng = cases.build_engine(blocks, nx, ny, cases.explosion_ic)
    try:
        rng = np.random.default_rng(7)
        base = {gid: eng.download(gid) for gid in sorted(blocks)}
        jobs = []
        for j in range(3):   # three independent inputs: the explosion state with a different smooth perturbation each
            st = {}
            for gid, U in base.items():
                V = U.copy()
                V[..., 0] *= 1.0 + 0.01 * (j + 1) * rng.random(V.shape[:2])
                V[..., 3] *= 1.0 + 0.01 * (j + 1) * rng.random(V.shape[:2])
                st[gid] = V
            jobs.append(st)
        want = [_run_sync(eng, st, 2) for st in jobs]

        pin_in = [{g: torch.from_numpy(U).pin_memory().numpy() for g, U in st.items()} for st in jobs]
        pin_out = [{g: torch.empty((ny, nx, 4), dtype=torch.float64).pin_memory().numpy() for g in st} for st in jobs]
        for g, U in pin_in[0].items():
            eng.upload_async(g, U)
        for j in range(len(jobs)):
            eng.commit_uploads()
            if j + 1 < len(jobs):
                for g, U in pin_in[j + 1].items():   # staged while job j computes
                    eng.upload_async(g, U)
            eng.apply_bc()
            dt = eng.get_dt(0.0, 1e9)
            eng.step(dt)
            dt = eng.get_dt(0.0, 1e9)
            eng.step(dt)
            for g in jobs[j]:
                eng.download_async(g, pin_out[j][g])
        eng.transfers_sync()
        for j in range(len(jobs)):
            for g in jobs[j]:
                assert np.array_equal(pin_out[j][g], want[j][g]), f"job {j} block {g}"
    finally:
        eng.close()

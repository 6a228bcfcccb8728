"""BASELINE.json's named configurations at their NAMED sizes against fingerprints of the unmodified reference
(tests/golden/named/*.npz, written by oracle/make_named_fingerprints.py in the build container):

* the weak-scaling block of bench.py: one block of 1024 x 1024 against the unmodified reference (the largest its object model
  fits in the build container), first 2 RK4 steps; and the headline size itself, one block of 2048 x 2048, first step, against the
  numpy restatement (`ws2048`, labelled so in its metadata);
* the remaining shipped examples at their shipped sizes: explosion (one 40 x 40 block) and implosion (40 x 40, RK2) from the initial
  condition to t_final, shockbox (50 x 50, RK2) up to the abort the reference itself raises in step 24, supersonic_step (ten 192 x 64
  blocks of an irregular topology, Mach 5 Dirichlet inlet, HLLL) for 100 steps: with these EVERY example the reference ships is
  replayed at the size it ships with;
* examples/supersonic_wedge at its shipped size (2 blocks of 60 x 60, Dirichlet inlet, reflection wall, 15 degree ramp), with
  the shipped HLLL flux and with the Roe flux BASELINE.json names, first 50 steps;
* examples/jet at its shipped size (9 stacked blocks of 1080 x 60, slip walls + Dirichlet inlet), shipped HLLL flux, first 50
  steps -- and with the HLLE flux BASELINE.json names (two-edit patched reference), whose run the reference itself aborts
  in step 17 with an unrealizable state: the engine must reproduce steps 5, 10, 16 AND report the abort in step 17;

* explosion_multi exactly as shipped -- 2 x 4 blocks of 150 x 150, Roe + Venkatakrishnan + Green-Gauss, RK4, CFL 0.7,
  reflection walls -- from the initial condition to t_final = 0.07 (1604 steps);
* the DMR scheme -- HLLL + Venkatakrishnan, primitive reconstruction, RK2, CFL 0.4 -- on 4 blocks of 500 x 500, the
  first 40 steps.

The reference's full states are too large to commit, so a fingerprint holds per checkpoint and block the sha256 of the
state BY VALUE (-0.0 folded onto +0.0: equal digests <=> np.array_equal), a strided subsample and the complete dt
sequence.  -m gpu: the CUDA engine replays the whole run through the C ABI and must reproduce every dt and every
digest (the 1e-10 tolerance of the north star is asserted on the subsample as well, so that a digest mismatch reports how
far off it is).  -m "not gpu": the numpy restatement (oracle/) and the CPU twin of the stage kernel's source on the
first checkpoint."""
import hashlib
import json
import os

import numpy as np
import pytest

import cases

NAMED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "named")


def value_digest(U):
    return hashlib.sha256((np.ascontiguousarray(U) + 0.0).tobytes()).hexdigest()


class Named:
    """A fingerprint file; quacks like golden_io.Fixture where tests/test_kernel_twin.py needs it to."""

    def __init__(self, name):
        path = os.path.join(NAMED, name + ".npz")
        if not os.path.exists(path):
            pytest.skip(f"{path} not generated (oracle/make_named_fingerprints.py {name})")
        self.z = np.load(path)
        self.meta = dict(json.loads(str(self.z["meta"])), gamma=cases.GAMMA)
        self.name = name
        self.gids = self.meta["gids"]
        self.nx, self.ny, self.stride = self.meta["nx"], self.meta["ny"], self.meta["stride"]
        self.blocks = getattr(cases, self.meta["mesh"])(*self.meta.get("mesh_args", []))
        self.ic = getattr(cases, self.meta["ic"])
        self.dts = self.z["dts"]
        self._u0 = None

    def scheme(self):
        m = self.meta
        return dict(flux=m["flux"], limiter=m["limiter"], recon=m["recon"], integrator=m["integrator"], CFL=m["CFL"], nqp=1)

    def __getitem__(self, key):           # U0_<gid> for the kernel twin's marshalling
        if key.startswith("U0_"):
            from pyhype_b200.mesh.quad_mesh import QuadMesh

            b = self.blocks[int(key[3:])]
            m = QuadMesh(self.nx, self.ny, NE=b["NE"], NW=b["NW"], SE=b["SE"], SW=b["SW"])
            return np.ascontiguousarray(self.ic(m.x[:, :, 0], m.y[:, :, 0]))
        return self.z[key]

    def check(self, n, states):
        """states: {gid: (ny, nx, 4)} after n steps"""
        for g in self.gids:
            U = states[g]
            sub, ref = U[:: self.stride, :: self.stride], self.z[f"sub_{n}_{g}"]
            for k in range(4):
                scale = max(np.abs(ref[..., k]).max(), 1e-300)
                assert np.abs(sub[..., k] - ref[..., k]).max() / scale <= 1e-10, (self.name, n, g, k)
            assert value_digest(U) == self.meta["digests"][f"{n}_{g}"], (self.name, n, g, np.abs(sub - ref).max())


ALL_NAMED = ["em", "dmr", "wedge", "wedge_roe", "jet", "jet_hlle", "ws1024", "ws2048", "explosion", "implosion", "shockbox", "step"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ALL_NAMED)
def test_named_config_at_named_size_reproduces_the_reference(name, stage_path):
    fp = Named(name)
    eng = cases.build_engine(fp.blocks, fp.nx, fp.ny, fp.ic, **fp.scheme())
    try:
        assert eng.stage_path() == stage_path
        done_total, t = 0, 0.0
        for n in fp.meta["checkpoints"]:
            k = n - done_total
            t, done, bad, dts = eng.run(t, fp.meta["t_final_nd"], max_steps=k, poll_every=16, record_dts=k)
            assert done == k and not bad, (name, n, done)
            assert np.array_equal(np.asarray(dts), fp.dts[done_total:n]), (name, n)
            done_total = n
            fp.check(n, {g: eng.download(g) for g in fp.gids})
        assert t == fp.meta["t_end"]
        if fp.meta["reached_t_final"]:
            assert not (t < fp.meta["t_final_nd"])
            assert eng.run(t, fp.meta["t_final_nd"], max_steps=4, poll_every=1)[1] == 0      # the run is over: no further step
        if "aborts_in_step" in fp.meta:
            # the reference's realizability check stops the run in the next step (Euler2D.py:144-152); the device loop takes
            # that step, flags the state and stops (it reports up to `poll_every` steps late, never silently)
            t2, done, bad, _ = eng.run(t, fp.meta["t_final_nd"], max_steps=3, poll_every=1)
            assert bad and done >= 1, (name, done, bad)
            assert not eng.realizable()
    finally:
        eng.close()


@pytest.mark.parametrize("name", [n for n in ALL_NAMED if n != "ws2048"])   # ws2048 IS the oracle's output (3 minutes of numpy)
def test_oracle_restatement_at_named_size_first_checkpoint(name):
    if name == "ws1024" and not os.environ.get("PYH_NAMED_ORACLE_ALL"):
        pytest.skip("half a minute of numpy on one 1024 x 1024 block: set PYH_NAMED_ORACLE_ALL=1 (the GPU suite replays it against the same fingerprint)")
    fp = Named(name)
    n = fp.meta["checkpoints"][0]
    if name.startswith("wedge"):
        n = fp.meta["checkpoints"][-1]          # 2 x 60^2: all 50 steps cost a second
    prob = cases.build_oracle(fp.blocks, fp.nx, fp.ny, fp.ic, **fp.scheme())
    short = {"dmr": 2, "jet": 2}.get(name)       # a million cells per step in numpy: half a minute each to the first checkpoint
    if short and not os.environ.get("PYH_NAMED_ORACLE_ALL"):
        # default suite: the first steps only, checked through the dt sequence -- every dt is the CFL minimum over the WHOLE state the
        # step before produced (the digests sit at the checkpoint; the full replay runs with PYH_NAMED_ORACLE_ALL=1, last run:
        # profiles/r02r_named_twin.txt, and the GPU suite replays these cases to every checkpoint against the same fingerprints)
        t, dts = prob.run(0.0, fp.meta["t_final_nd"], max_steps=short)
        assert np.array_equal(np.asarray(dts), fp.dts[:short])
        return
    t, dts = prob.run(0.0, fp.meta["t_final_nd"], max_steps=n)
    assert np.array_equal(np.asarray(dts), fp.dts[:n])
    fp.check(n, {g: prob.blocks[g].U for g in fp.gids})


@pytest.mark.parametrize("name", ["em", "dmr"])
def test_stage_kernel_source_twin_at_named_size_first_checkpoint(name):
    """The device-resident time loop of the product (k_dt, k_dt_finalize, k_stage_march in the shipped 128-thread x 64-row
    strip shape, k_ghost, k_step_end), compiled by g++ and run by the thread-block emulator of tests/host_twin/, at the named
    size: two column strips and three row strips per 150 x 150 block, four by eight per 500 x 500 block.  One and two
    minutes of emulation: they run only with PYH_NAMED_TWIN_ALL=1 (last run: profiles/r02r_named_twin.txt)."""
    if not os.environ.get("PYH_NAMED_TWIN_ALL"):
        pytest.skip("one (explosion_multi) and two (DMR) minutes of emulation: set PYH_NAMED_TWIN_ALL=1; the shockbox replay below is the "
                    "named-size twin run of the default suite")
    import test_kernel_twin as T

    fp = Named(name)
    n = fp.meta["checkpoints"][0]
    idx, Uout, dts, t, nsteps, bad = T.run_loop(T.build("default"), fp, 0.0, fp.meta["t_final_nd"], n, nt=128, tys=64)
    assert nsteps == n and not bad
    assert np.array_equal(dts, fp.dts[:n])
    fp.check(n, {g: Uout[idx[g]] for g in fp.gids})


def test_stage_kernel_source_twin_reports_the_abort_of_shockbox():
    """examples/shockbox: in step 24 a slope of 1.4e211 overflows the Venkatakrishnan quotient to inf / inf; np.minimum.reduce
    (limiters/base.py:179-186) carries the NaN into phi, four cells of the state turn NaN and the reference aborts
    (Euler2D.py:144-152).  The kernel source must produce the same NaN cells in the same step (its minimum over the faces
    is NaN-propagating for exactly this reason) and raise the flag; found by the first GPU run of the shockbox fingerprint."""
    import test_kernel_twin as T

    fp = Named("shockbox")
    n = fp.meta["aborts_in_step"]
    idx, Uout, dts, t, nsteps, bad = T.run_loop(T.build("default"), fp, 0.0, fp.meta["t_final_nd"], n + 2, nt=64, tys=16)
    assert bad and nsteps == n
    assert np.array_equal(dts[: n - 1], fp.dts[: n - 1])
    prob = cases.build_oracle(fp.blocks, fp.nx, fp.ny, fp.ic, **fp.scheme())
    with pytest.raises(RuntimeError, match="unrealizable"):
        prob.run(0.0, fp.meta["t_final_nd"], max_steps=n)
    g = fp.gids[0]
    assert np.isnan(prob.blocks[g].U).sum() == 16 and not prob.realizable()
    assert np.array_equal(Uout[idx[g]], prob.blocks[g].U, equal_nan=True)

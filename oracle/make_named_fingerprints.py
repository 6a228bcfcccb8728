"""Fingerprints of BASELINE.json's NAMED configurations at their NAMED sizes, produced by running the
UNMODIFIED reference (/root/reference, through oracle/refharness.py) in the build container.

TEST INFRASTRUCTURE ONLY (see oracle/refharness.py).

    python oracle/make_named_fingerprints.py em     # explosion_multi: 2x4 blocks of 150x150, Roe + Venkatakrishnan,
                                                    # RK4, CFL 0.7 as shipped, from the initial condition to t_final = 0.07
                                                    # (1604 steps; ~40 min of one core)
    python oracle/make_named_fingerprints.py dmr    # DMR: 4 blocks of 500x500, HLLL + Venkatakrishnan, primitive
                                                    # reconstruction, RK2, CFL 0.4, the first 40 steps (~10 min)

The full states are too large to commit (5.8 MB / 32 MB), so each checkpoint stores, per block: the sha256 of the
state BY VALUE (``(U + 0.0).tobytes()``: adding +0.0 maps -0.0 to +0.0, every other double to itself, so the digest
is equal exactly when ``np.array_equal`` would be), a strided subsample of the state (for a readable diff when a
digest differs) and the four sums; plus the complete dt sequence.  tests/test_gpu_named_configs.py replays the run
on the GPU from the same initial condition and compares digest by digest; tests/test_oracle_golden.py does the
same for the numpy restatement on the first checkpoint.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import refharness as rh  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "named")

CASES = {
    "em": dict(mesh="em_mesh", ic="explosion_ic", nx=150, ny=150, stride=15, checkpoints=[10, 50, 200, 600, 1200, -1],
               cfg=dict(t_final=0.07),
               what="examples/explosion_multi as shipped (CFL 0.7), initial condition -> t_final"),
    "dmr": dict(mesh="dmr_mesh", ic="dmr_ic", nx=500, ny=500, stride=50, checkpoints=[5, 20, 40],
                cfg=dict(fvm_flux_function_type="HLLL", time_integrator="RK2", CFL=0.4, reconstruction_type="primitive",
                         t_final=0.25),
                what="examples/dmr scheme at BASELINE.json's 500x500 blocks (shipped: 50x50), first 40 steps"),
    # BASELINE.json configs[4] towards the HEADLINE BLOCK SIZE: one block of the weak-scaling mesh (bench.py's workload is 8 blocks
    # of 2048 x 2048 per GPU; the reference's object model is killed by the container's memory limit at 2048 x 2048, so the
    # reference pins 1024 x 1024 and oracle/make_ws2048_fingerprint.py pins 2048 x 2048 with the numpy restatement), explosion box
    # inside, Roe + Venkatakrishnan, RK4, CFL 0.7; the first 2 steps
    "ws1024": dict(mesh="ws_mesh", mesh_args=(1, 1), ic="ws_ic_1x1", nx=1024, ny=1024, stride=32, checkpoints=[1, 2],
                   cfg=dict(t_final=0.07),
                   what="synthetic weak-scaling explosion, one block of 1024 x 1024 (the largest the unmodified reference fits in this container), first 2 RK4 steps"),
    # the headline block size itself, pinned by the NUMPY RESTATEMENT (oracle/muscl_oracle.py, which the cases above and the 28
    # fixtures pin to the reference): the reference's object model does not fit 2048 x 2048 in this container
    "ws2048": dict(mesh="ws_mesh", mesh_args=(1, 1), ic="ws_ic_1x1", nx=2048, ny=2048, stride=64, checkpoints=[1], by_oracle=True,
                   cfg=dict(t_final=0.07),
                   what="synthetic weak-scaling explosion at the headline block size: one block of 2048 x 2048, first RK4 step, by the numpy oracle"),
    # the remaining shipped examples at their shipped sizes (single-block problems are renumbered to block 0)
    "explosion": dict(mesh="em_mesh", mesh_args=(1, 1), ic="explosion_ic", nx=40, ny=40, stride=4, checkpoints=[50, -1],
                      cfg=dict(t_final=0.07),
                      what="examples/explosion as shipped (one 40 x 40 block, Roe, RK4, CFL 0.7), initial condition -> t_final"),
    "implosion": dict(mesh="em_mesh", mesh_args=(1, 1, 10.0, 10.0), ic="implosion_ic", nx=40, ny=40, stride=4, checkpoints=[100, -1],
                      cfg=dict(time_integrator="RK2", CFL=0.4, t_final=0.1),
                      what="examples/implosion as shipped (one 40 x 40 block, Roe, RK2, CFL 0.4), initial condition -> t_final"),
    "shockbox": dict(mesh="em_mesh", mesh_args=(1, 1, 10.0, 10.0), ic="shockbox_ic", nx=50, ny=50, stride=5, checkpoints=[10, 23], aborts_next=True,
                     cfg=dict(time_integrator="RK2", CFL=0.4, t_final=2.0),
                     what="examples/shockbox as shipped (one 50 x 50 block, Roe, RK2, CFL 0.4): the reference's own run stops with an unrealizable "
                          "state in step 24, so the fingerprint holds steps 10, 23 and the abort"),
    "step": dict(mesh="step_mesh", mesh_args=(64,), ic="step_ic", nx=192, ny=64, stride=16, checkpoints=[10, 100],
                 cfg=dict(fvm_flux_function_type="HLLL", time_integrator="RK2", CFL=0.3, reconstruction_type="primitive", t_final=20.0),
                 what="examples/supersonic_step as shipped (ten 192 x 64 blocks, Mach 5 Dirichlet inlet, HLLL, primitive reconstruction, RK2, CFL 0.3), first 100 steps"),
    # BASELINE.json configs[2]: examples/supersonic_wedge at its shipped size (2 blocks of 60 x 60, Dirichlet inlet, reflection wall,
    # 15 degree ramp), with the shipped flux (HLLL) and with the one BASELINE.json names (Roe); first 50 steps
    "wedge": dict(mesh="wedge_mesh", mesh_args=(60,), ic="wedge_ic", nx=60, ny=60, stride=6, checkpoints=[10, 50],
                  cfg=dict(fvm_flux_function_type="HLLL", time_integrator="RK2", CFL=0.3, reconstruction_type="primitive", t_final=20.0),
                  what="examples/supersonic_wedge as shipped (HLLL, primitive reconstruction, RK2, CFL 0.3), first 50 steps"),
    "wedge_roe": dict(mesh="wedge_mesh", mesh_args=(60,), ic="wedge_ic", nx=60, ny=60, stride=6, checkpoints=[10, 50],
                      cfg=dict(fvm_flux_function_type="Roe", time_integrator="RK2", CFL=0.3, reconstruction_type="primitive", t_final=20.0),
                      what="examples/supersonic_wedge with the Roe flux BASELINE.json names, first 50 steps"),
    # BASELINE.json configs[3]: examples/jet at its shipped size (9 stacked blocks of 1080 x 60), shipped flux (HLLL) and the
    # HLLE that BASELINE.json names (runs only in the two-edit patched reference); first 50 steps
    "jet": dict(mesh="jet_mesh", mesh_args=(60,), ic="jet_ic", nx=1080, ny=60, stride=12, checkpoints=[10, 50],
                cfg=dict(fvm_flux_function_type="HLLL", time_integrator="RK2", CFL=0.4, reconstruction_type="primitive", t_final=25.0),
                what="examples/jet as shipped (HLLL, primitive reconstruction, RK2, CFL 0.4), first 50 steps"),
    "jet_hlle": dict(mesh="jet_mesh", mesh_args=(60,), ic="jet_ic", nx=1080, ny=60, stride=12, checkpoints=[5, 10, 16], patch_hlle=True, aborts_next=True,
                     cfg=dict(fvm_flux_function_type="HLLE", time_integrator="RK2", CFL=0.4, reconstruction_type="primitive", t_final=25.0),
                     what="examples/jet with the HLLE flux BASELINE.json names: two-edit PATCHED reference (oracle/refharness.patch_hlle); the reference's own "
                          "run stops with an unrealizable state in step 17 (inlet start-up), so the fingerprint holds steps 5, 10, 16 and the abort"),
}


def value_digest(U):
    return hashlib.sha256((np.ascontiguousarray(U) + 0.0).tobytes()).hexdigest()


class OracleRun:
    """RefRun's interface (oracle/refharness.py) on top of the numpy restatement, for sizes the reference cannot hold."""

    class _Blk:
        def __init__(self, g, b):
            self.global_block_num, self._b = g, b

        @property
        def state(self):
            return type("S", (), {"data": self._b.U})()

    class _Solver:
        pass

    def __init__(self, cases, blocks, c, ic, config):
        recon = c["cfg"].get("reconstruction_type", "conservative")
        self.prob = cases.build_oracle(blocks, c["nx"], c["ny"], ic, flux=config.fvm_flux_function_type, limiter=config.fvm_slope_limiter_type,
                                       recon=recon, integrator=config.time_integrator, CFL=config.CFL)
        self.solver = self._Solver()
        self.solver.t = 0.0
        self.solver.t_final = config.t_final * 343.0
        self.dts = []

    @property
    def blocks(self):
        return [self._Blk(g, b) for g, b in sorted(self.prob.blocks.items())]

    def step(self, n=1):
        t, dts = self.prob.run(self.solver.t, self.solver.t_final, max_steps=n)
        self.solver.t = t
        self.dts += dts
        return self


def main(name):
    import cases
    from make_golden import ref_blocks

    c = CASES[name]
    if c.get("patch_hlle"):
        rh.activate()
        rh.patch_hlle()
    blocks = getattr(cases, c["mesh"])(*c.get("mesh_args", ()))
    ic = getattr(cases, c["ic"])

    class IC:
        def apply_to_block(self, block):
            block.state.data = np.ascontiguousarray(ic(block.mesh.x[:, :, 0], block.mesh.y[:, :, 0]))

    config = rh.make_config(nx=c["nx"], ny=c["ny"], initial_condition=IC(), **c["cfg"])
    if c.get("by_oracle"):
        run = OracleRun(cases, blocks, c, ic, config)
    else:
        run = rh.RefRun(config, ref_blocks(blocks))
    t_final = float(run.solver.t_final)
    out, meta = {}, dict(name=name, what=c["what"], nx=c["nx"], ny=c["ny"], stride=c["stride"], mesh=c["mesh"], mesh_args=list(c.get("mesh_args", ())), ic=c["ic"],
                         flux=config.fvm_flux_function_type, limiter=config.fvm_slope_limiter_type,
                         recon=c["cfg"].get("reconstruction_type", "conservative"), integrator=config.time_integrator,
                         CFL=config.CFL, t_final_nd=t_final, gids=sorted(blocks), checkpoints=[], digests={}, raw_sha16={},
                         generator=("oracle/make_named_fingerprints.py on the NUMPY ORACLE (oracle/muscl_oracle.py), not the reference" if c.get("by_oracle") else
                                    "oracle/make_named_fingerprints.py on the unmodified reference" + (" + HLLE 2-edit patch" if c.get("patch_hlle") else "")))
    t0 = time.time()
    for cp in c["checkpoints"]:
        target = 10**9 if cp < 0 else cp
        run.step(target - len(run.dts))
        n = len(run.dts)
        meta["checkpoints"].append(n)
        sums = []
        for blk in run.blocks:
            g = blk.global_block_num
            U = blk.state.data
            meta["digests"][f"{n}_{g}"] = value_digest(U)
            meta["raw_sha16"][f"{n}_{g}"] = hashlib.sha256(U.tobytes()).hexdigest()[:16]   # SURVEY.md section 8c's form
            out[f"sub_{n}_{g}"] = U[:: c["stride"], :: c["stride"]].copy()
            sums.append(U.sum(axis=(0, 1)))
        out[f"sums_{n}"] = np.array(sums)
        print(f"[{name}] step {n}  t = {run.solver.t:.6f} / {t_final:.6f}  ({time.time() - t0:.0f} s)", flush=True)
    if c.get("aborts_next"):
        # the reference's realizability check (Euler2D.py:144-152) must stop the run in the next step
        t_before = float(run.solver.t)
        try:
            run.step(1)
            raise RuntimeError("expected the reference to abort in the next step")
        except SystemExit:
            meta["aborts_in_step"] = n + 1
        run.dts = run.dts[:n]
        run.solver.t = t_before
    meta["t_end"] = float(run.solver.t)
    meta["reached_t_final"] = not (run.solver.t < t_final)
    out["dts"] = np.array(run.dts)
    out["meta"] = np.array(json.dumps(meta))
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main(sys.argv[1])

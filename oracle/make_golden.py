"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(/root/reference) in the build container (see oracle/refharness.py for the stubs).

    python oracle/make_golden.py            # rewrites tests/golden/*.npz
    python oracle/make_golden.py NAME ...   # only the named fixtures

Each fixture stores the block dictionary (vertices, neighbours, BC types, Dirichlet inlet
strips), a starting state U0 per block (the reference's own state after `pre` steps, i.e. a
post-transient field), the reference's dUdt / gradients / limiter at U0, its ghost strips, and
its state and dt sequence after `steps` further steps.  HLLE fixtures come from the two-edit
patched reference (refharness.patch_hlle) and say so in their metadata.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import refharness as rh  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
SIDES = ("E", "W", "N", "S")
DIRS = {"E": 1, "W": -1, "N": 2, "S": -2}


def ref_blocks(blocks):
    """Turn ndarray inlet states of tests/cases.py block dicts into reference PrimitiveDirichletBC objects."""
    rh.activate()
    from pyhype.boundary_conditions.base import PrimitiveDirichletBC
    from pyhype.fluids import Air
    from pyhype.states import PrimitiveState

    air = Air(a_inf=1.0, rho_inf=1.0)  # strips in cases.py are already non-dimensional
    out = {}
    for gid, b in blocks.items():
        nb = dict(b)
        for s in SIDES:
            v = b["BCType" + s]
            if isinstance(v, np.ndarray):
                nb["BCType" + s] = PrimitiveDirichletBC(primitive_state=PrimitiveState(fluid=air, array=v.copy()))
        out[gid] = nb
    return out


ONLY = set()   # fixture names given on the command line (empty = all)


def make(name, blocks, nx, ny, ic, pre, steps, note="", **cfg):
    import cases

    if ONLY and name not in ONLY:
        return

    rh.activate()
    kw = dict(nx=nx, ny=ny, initial_condition=rh.CallableIC(lambda x, y: None))
    kw.update(cfg)
    recon = kw.get("reconstruction_type", "conservative")

    class IC:
        def apply_to_block(self, block):
            block.state.data = np.ascontiguousarray(ic(block.mesh.x[:, :, 0], block.mesh.y[:, :, 0]))

    kw["initial_condition"] = IC()
    config = rh.make_config(**kw)
    run = rh.RefRun(config, ref_blocks(blocks))
    run.step(pre)
    out = {}
    meta = dict(
        name=name, nx=nx, ny=ny, pre=pre, steps=steps, note=note,
        flux=config.fvm_flux_function_type, limiter=config.fvm_slope_limiter_type, recon=recon,
        integrator=config.time_integrator, CFL=config.CFL, nqp=config.fvm_num_quadrature_points,
        gamma=1.4, gids=sorted(blocks),
        generator="oracle/make_golden.py on the unmodified reference" + (" + HLLE 2-edit patch" if config.fvm_flux_function_type == "HLLE" else ""),
        blocks={},
    )
    for gid, b in blocks.items():
        mb = {k: [float(v) for v in b[k]] for k in ("SW", "SE", "NW", "NE")}
        for s in SIDES:
            mb["Neighbor" + s] = b["Neighbor" + s]
            v = b["BCType" + s]
            if isinstance(v, np.ndarray):
                mb["BCType" + s] = "@dirichlet"
                out[f"dirichlet_{gid}_{s}"] = v
            else:
                mb["BCType" + s] = v
        meta["blocks"][str(gid)] = mb
    res = run.residuals()
    gh = run.ghosts()
    for blk, r, g in zip(run.blocks, res, gh):
        gid = blk.global_block_num
        out[f"U0_{gid}"] = blk.state.data.copy()
        for k in ("R", "gx", "gy", "phi"):
            out[f"{k}_{gid}"] = r[k]
        for s in SIDES:
            out[f"ghost0_{gid}_{s}"] = g[s]
        out[f"xc_{gid}"] = blk.mesh.x[:, :, 0].copy()
        out[f"yc_{gid}"] = blk.mesh.y[:, :, 0].copy()
        out[f"A_{gid}"] = blk.mesh.A[:, :, 0].copy()
        out[f"thetaE_{gid}"] = blk.mesh.face.E.theta[:, :, 0].copy()
        out[f"thetaN_{gid}"] = blk.mesh.face.N.theta[:, :, 0].copy()
    n0 = len(run.dts)
    run.solver.t = 0.0
    run.solver.t_final = 1e9
    run.step(steps)
    out["dts"] = np.array(run.dts[n0:])
    for blk in run.blocks:
        out[f"U_{blk.global_block_num}"] = blk.state.data.copy()
    out["meta"] = np.array(json.dumps(meta))
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


def main():
    import cases

    ONLY.update(sys.argv[1:])
    os.makedirs(GOLDEN, exist_ok=True)
    rh.activate()
    rh.patch_hlle()
    em = cases.em_mesh()
    make("em_roe_venkat_cons_rk4", em, 24, 24, cases.explosion_ic, pre=6, steps=8)
    make("em_ragged_roe_rk4", em, 19, 11, cases.explosion_ic, pre=3, steps=5)
    make("cart_roe_cons_rk4", cases.em_mesh(nbx=1, nby=1), 20, 40, cases.explosion_ic, pre=4, steps=5)
    make("dmr_hlll_venkat_prim_rk2", cases.dmr_mesh(), 24, 24, cases.dmr_ic, pre=8, steps=12,
         fvm_flux_function_type="HLLL", time_integrator="RK2", CFL=0.4, reconstruction_type="primitive")
    make("dmr_roe_cons_rk2", cases.dmr_mesh(), 20, 20, cases.dmr_ic, pre=5, steps=8,
         time_integrator="RK2", CFL=0.4)
    make("dmr_hlle_prim_rk2", cases.dmr_mesh(), 16, 16, cases.dmr_ic, pre=4, steps=6, note="patched oracle (2 edits)",
         fvm_flux_function_type="HLLE", time_integrator="RK2", CFL=0.4, reconstruction_type="primitive")
    make("wedge_hlll_prim_rk2", cases.wedge_mesh(16), 18, 16, cases.wedge_ic, pre=10, steps=10,
         fvm_flux_function_type="HLLL", time_integrator="RK2", CFL=0.3, reconstruction_type="primitive")
    make("wedge_roe_cons_rk2", cases.wedge_mesh(12), 16, 12, cases.wedge_ic, pre=10, steps=10,
         time_integrator="RK2", CFL=0.3)
    make("smooth_roe_rk4", cases.em_mesh(nbx=2, nby=2, east=5.0, north=5.0), 16, 16, cases.smooth_ic, pre=2, steps=6)
    for lim in ("VanLeer", "VanAlbada", "BarthJespersen"):
        make(f"em_lim_{lim}", em, 12, 12, cases.explosion_ic, pre=3, steps=4, fvm_slope_limiter_type=lim)
    for integ in ("ExplicitEuler1", "RK2", "Ralston2", "RK3", "RK3SSP", "Ralston3", "Ralston4", "DormandPrince5"):
        make(f"em_int_{integ}", em, 12, 12, cases.explosion_ic, pre=2, steps=4, time_integrator=integ, CFL=0.3)
    for nq in (2, 3):
        make(f"em_nqp{nq}", em, 12, 12, cases.explosion_ic, pre=2, steps=3, fvm_num_quadrature_points=nq)
    make("em_hlle_cons_rk2", em, 12, 12, cases.explosion_ic, pre=3, steps=5, note="patched oracle (2 edits)",
         fvm_flux_function_type="HLLE", time_integrator="RK2")
    # the remaining shipped examples (BASELINE.json configs[3] = jet; supersonic_step, implosion, shockbox)
    make("jet_hlll_prim_rk2", cases.jet_mesh(6), 36, 6, cases.jet_ic, pre=12, steps=10,
         fvm_flux_function_type="HLLL", time_integrator="RK2", CFL=0.4, reconstruction_type="primitive")
    make("jet_hlle_prim_rk2", cases.jet_mesh(6), 36, 6, cases.jet_ic, pre=12, steps=10, note="patched oracle (2 edits)",
         fvm_flux_function_type="HLLE", time_integrator="RK2", CFL=0.4, reconstruction_type="primitive")
    make("step_hlll_prim_rk2", cases.step_mesh(8), 24, 8, cases.step_ic, pre=15, steps=10,
         fvm_flux_function_type="HLLL", time_integrator="RK2", CFL=0.3, reconstruction_type="primitive")
    one = cases.em_mesh(nbx=1, nby=1, east=10.0, north=10.0)
    make("implosion_roe_cons_rk2", one, 24, 24, cases.implosion_ic, pre=6, steps=8, time_integrator="RK2", CFL=0.4)
    make("shockbox_roe_cons_rk2", one, 24, 24, cases.shockbox_ic, pre=6, steps=8, time_integrator="RK2", CFL=0.4)


if __name__ == "__main__":
    main()

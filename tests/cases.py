"""Shared problem definitions for the parity tests: the same block dictionaries and initial
states are fed to the oracle (``oracle/muscl_oracle.py``) and to the CUDA engine."""
from __future__ import annotations

import numpy as np

from pyhype_b200.mesh.base import QuadMeshGenerator
from pyhype_b200.mesh.rectangular import RectagularMeshGenerator

from pyhype_b200.examples import (  # noqa: F401  (shared with bench.py: the same inputs are benchmarked and tested)
    A_INF, GAMMA, RHO_INF, SIDES, _nd_inlet, dmr_ic, dmr_mesh, em_mesh, explosion_ic, jet_ic, jet_mesh, prim_to_cons_nd,
    smooth_ic, wedge_ic, wedge_mesh, ws_ic, ws_ic_1x1, ws_ic_smooth, ws_mesh,
)


def step_mesh(ny):
    """examples/supersonic_step/mesh.py (step_ten_block): ten 3 x 1 blocks around a forward-facing step,
    Mach 5 Dirichlet inlet on the west side, block numbering NOT a regular grid."""
    inlet = _nd_inlet([1.0, 5.0, 0.0, 1 / 1.4], ny).reshape(ny, 1, 4)
    wall, out = "Slipwall", "OutletDirichlet"
    #      gid: (x0, y0,   E     W     N     S,    bcE   bcW    bcN   bcS)
    spec = {
        0: (0, 0, None, None, 1, None, wall, inlet, None, wall),
        1: (0, 1, 6, None, 2, 0, None, inlet, None, None),
        2: (0, 2, 5, None, 3, 1, None, inlet, None, None),
        3: (0, 3, 4, None, None, 2, None, inlet, wall, None),
        4: (3, 3, 9, 3, None, 5, None, None, wall, None),
        5: (3, 2, 8, 2, 4, 6, None, None, None, None),
        6: (3, 1, 7, 1, 5, None, None, None, None, wall),
        7: (6, 1, None, 6, 8, None, out, None, None, wall),
        8: (6, 2, None, 5, 9, 7, out, None, None, None),
        9: (6, 3, None, 4, None, 8, out, None, wall, None),
    }
    blocks = {}
    for gid, (x0, y0, nE, nW, nN, nS, bE, bW, bN, bS) in spec.items():
        blocks[gid] = dict(
            nBLK=gid, NW=[x0, y0 + 1], NE=[x0 + 3, y0 + 1], SW=[x0, y0], SE=[x0 + 3, y0],
            NeighborE=nE, NeighborW=nW, NeighborN=nN, NeighborS=nS,
            NeighborNE=None, NeighborNW=None, NeighborSE=None, NeighborSW=None,
            BCTypeE=bE, BCTypeW=bW, BCTypeN=bN, BCTypeS=bS,
            BCTypeNE=None, BCTypeNW=None, BCTypeSE=None, BCTypeSW=None,
        )
    return blocks


def step_ic(x, y):
    """pyhype/initial_conditions/supersonic_flood.py with the values of examples/supersonic_step/config.py:8-14."""
    W = np.empty(x.shape + (4,))
    W[...] = np.array([1.0, 5.0, 0.0, 1 / GAMMA])
    return prim_to_cons_nd(W)


def _two_state(cond):
    UL = prim_to_cons_nd(np.array([4.6968, 0.0, 0.0, 404400.0]).reshape(1, 1, 4))
    UR = prim_to_cons_nd(np.array([1.1742, 0.0, 0.0, 101100.0]).reshape(1, 1, 4))
    return np.where(cond[..., None], UR, UL)


def implosion_ic(x, y):
    """examples/implosion/initial_condition.py:33-60 (low-pressure corner x, y <= 5)."""
    return _two_state(np.logical_and(x <= 5, y <= 5))


def shockbox_ic(x, y):
    """examples/shockbox/initial_condition.py:33-63 (two opposite low-pressure quadrants)."""
    return _two_state(np.logical_or(np.logical_and(x <= 5, y <= 5), np.logical_and(x > 5, y > 5)))


def build_oracle(blocks, nx, ny, ic, **kw):
    from oracle import muscl_oracle as mo

    prob = mo.Problem(blocks, nx, ny, **kw)
    for b in prob.blocks.values():
        b.U = ic(b.geom.xc, b.geom.yc)
    prob.apply_bc()
    return prob


def build_engine(blocks, nx, ny, ic, flux="Roe", limiter="Venkatakrishnan", recon="conservative", integrator="RK4",
                 CFL=0.7, device=0, local=None, states=None, nqp=1):
    from pyhype_b200.engine import Engine
    from pyhype_b200.mesh.quad_mesh import QuadMesh
    from pyhype_b200.time_marching import TABLEAUX

    tab = TABLEAUX[integrator] if isinstance(integrator, str) else integrator
    eng = Engine(nx, ny, flux, limiter, recon, tab, GAMMA, CFL, device=device, num_quadrature_points=nqp)
    gids = sorted(blocks) if local is None else sorted(local)
    meshes = {}
    for gid in gids:
        b = blocks[gid]
        m = QuadMesh(nx, ny, NE=b["NE"], NW=b["NW"], SE=b["SE"], SW=b["SW"])
        meshes[gid] = m
        eng.add_block(gid, m, {s: b["Neighbor" + s] for s in SIDES}, {s: b["BCType" + s] for s in SIDES},
                      local_gids=set(gids))
    eng.finalize()
    for gid in gids:
        m = meshes[gid]
        U = states[gid] if states is not None else ic(m.x[:, :, 0], m.y[:, :, 0])
        eng.upload(gid, U)
    eng.apply_bc()
    eng.meshes = meshes
    return eng

"""The reference's example problems as data: block dictionaries (what its mesh generators produce) and initial states as
functions of the cell centroids -- examples/explosion_multi, examples/dmr, examples/supersonic_wedge, examples/jet and the
smooth field of SURVEY.md section 8d.  Shared by bench.py (named-configuration lines), tools/ and tests/cases.py, so that the
benchmark and the parity tests run literally the same inputs.  Host-side numpy only; no arithmetic of the hot path."""
from __future__ import annotations

import numpy as np

from .mesh.base import QuadMeshGenerator
from .mesh.rectangular import RectagularMeshGenerator

GAMMA = 1.4
A_INF = 343.0
RHO_INF = 1.0
SIDES = ("E", "W", "N", "S")


def prim_to_cons_nd(W, a_inf=A_INF, rho_inf=RHO_INF, g=GAMMA):
    """PrimitiveState(...).to_type(ConservativeState) + make_non_dimensional
    (pyhype/states/converter/concrete_defs.py:127-141, pyhype/states/base.py:93-97)."""
    rho, u, v, p = (W[..., k] for k in range(4))
    ek = 0.5 * rho * (u * u + v * v)
    U = np.stack((rho.copy(), rho * u, rho * v, p / (g - 1) + ek), axis=-1)
    U[..., 0] /= rho_inf
    U[..., 1] /= rho_inf * a_inf
    U[..., 2] /= rho_inf * a_inf
    U[..., 3] /= rho_inf * a_inf**2
    return U


def em_mesh(nbx=2, nby=4, east=10.0, north=20.0):
    return RectagularMeshGenerator.generate(
        BCE=["Reflection"], BCW=["Reflection"], BCN=["Reflection"], BCS=["Reflection"],
        east=east, west=0.0, north=north, south=0.0, n_blocks_horizontal=nbx, n_blocks_vertical=nby,
    ).dict


def explosion_ic(x, y, lo=3.0, hi=7.0):
    """examples/explosion/initial_condition.py:35-60"""
    inside = np.logical_and(np.logical_and(x >= lo, x <= hi), np.logical_and(y >= lo, y <= hi))
    WL = np.array([4.6968, 0.0, 0.0, 404400.0])
    WR = np.array([1.1742, 0.0, 0.0, 101100.0])
    UL = prim_to_cons_nd(WL.reshape(1, 1, 4))
    UR = prim_to_cons_nd(WR.reshape(1, 1, 4))
    # the reference fills dimensional conservative states, then non-dimensionalises the whole block
    return np.where(inside[..., None], UL, UR)


def dmr_mesh():
    """examples/dmr/mesh.py"""
    k = 1
    a = 2 / np.sqrt(3)
    d = np.tan(30 * np.pi / 180)
    xs = [0, k, 2 * k, 3 * k, 4 * k]
    return QuadMeshGenerator(
        nx_blk=4, ny_blk=1, BCE=["OutletDirichlet"], BCW=["OutletDirichlet"], BCN=["OutletDirichlet"],
        BCS=["OutletDirichlet", "Slipwall", "Slipwall", "Slipwall"],
        top_x=xs, bot_x=xs, top_y=[a, a, a + d, a + 2 * d, a + 3 * d], bot_y=[0, 0, d, 2 * d, 3 * d],
        left_x=[0, 0], right_x=[4 * k, 4 * k], left_y=[0, a], right_y=[3 * d, a + 3 * d],
    ).dict


def dmr_ic(x, y):
    """examples/dmr/initial_condition.py:36-59"""
    UL = prim_to_cons_nd(np.array([8.0, 8.25, 0.0, 116.5]).reshape(1, 1, 4))
    UR = prim_to_cons_nd(np.array([1.4, 0.0, 0.0, 1.0]).reshape(1, 1, 4))
    return np.where((x <= 0.95)[..., None], UL, UR)


def smooth_ic(x, y):
    """rounding-robust smooth field (IC-B)"""
    rho = 1.2 + 0.3 * np.sin(0.7 * x + 0.3) * np.cos(0.45 * y + 0.1)
    u = 30 * np.cos(0.5 * x) * np.sin(0.35 * y + 0.2)
    v = -25 * np.sin(0.4 * x + 0.5) * np.cos(0.3 * y)
    p = 101325 * (1 + 0.2 * np.cos(0.6 * x - 0.2) * np.sin(0.5 * y + 0.4))
    return prim_to_cons_nd(np.stack((rho, u, v, p), axis=-1))


def wedge_mesh(ny, with_inlet=True):
    """examples/supersonic_wedge/mesh.py renumbered from 0."""
    t15 = 2 * np.tan(15 * np.pi / 180)
    inlet = np.tile(np.array([1.0, 2.0, 0.0, 1 / GAMMA]), (ny, 1))
    inlet = inlet / np.array([RHO_INF, A_INF, A_INF, RHO_INF * A_INF**2])  # make_non_dimensional on a primitive state
    nil = dict(NeighborN=None, NeighborS=None, NeighborNE=None, NeighborNW=None, NeighborSE=None, NeighborSW=None,
               BCTypeNE=None, BCTypeNW=None, BCTypeSE=None, BCTypeSW=None)
    b0 = dict(nBLK=0, NW=[0, 2], NE=[2, 2], SW=[0, 0], SE=[2, 0], NeighborE=1, NeighborW=None, BCTypeE=None,
              BCTypeW=inlet.reshape(ny, 1, 4) if with_inlet else "OutletDirichlet", BCTypeN="OutletDirichlet", BCTypeS="Reflection", **nil)
    b1 = dict(nBLK=1, NW=[2, 2], NE=[4, 2 + t15], SW=[2, 0], SE=[4, t15], NeighborE=None, NeighborW=0,
              BCTypeE="OutletDirichlet", BCTypeW=None, BCTypeN="OutletDirichlet", BCTypeS="Reflection", **nil)
    return {0: b0, 1: b1}


def wedge_ic(x, y):
    W = np.empty(x.shape + (4,))
    W[...] = np.array([1.0, 2.0, 0.0, 1 / GAMMA])
    return prim_to_cons_nd(W)


def _nd_inlet(W, n):
    """PrimitiveDirichletBC state tiled to the ghost strip, non-dimensionalised like
    pyhype/boundary_conditions/base.py:41 does on construction."""
    inlet = np.tile(np.asarray(W, dtype=float), (n, 1))
    return inlet / np.array([RHO_INF, A_INF, A_INF, RHO_INF * A_INF**2])


def jet_mesh(ny):
    """examples/jet/mesh.py: nine blocks stacked south -> north, slip walls on the west side except the
    Dirichlet inlet of the middle block, outflow elsewhere."""
    inlet = _nd_inlet([1.0, 0.1, 0.0, 2.0 / GAMMA], ny).reshape(ny, 1, 4)
    return QuadMeshGenerator(
        nx_blk=1, ny_blk=9, BCE=["OutletDirichlet"] * 9, BCW=["Slipwall"] * 4 + [inlet] + ["Slipwall"] * 4,
        BCN=["OutletDirichlet"], BCS=["OutletDirichlet"], NE=(1, 0.5), SW=(0, 0), NW=(0, 0.5), SE=(1, 0),
    ).dict


def jet_ic(x, y):
    """examples/jet/initial_condition.py:33-46 (air at rest, p = 1/gamma), broadcast to the block."""
    W = np.empty(x.shape + (4,))
    W[...] = np.array([1.0, 0.0, 0.0, 1 / GAMMA])
    return prim_to_cons_nd(W)


# ---- synthetic weak-scaling explosion (BASELINE.json configs[4]) --------------------------
WS_BLOCK_LEN = 1.25   # length units per block side


def ws_mesh(n_rows, blocks_per_row):
    """`blocks_per_row` x `n_rows` rectangular blocks (one row of 8 per GPU in the benchmark), reflection walls all round."""
    return RectagularMeshGenerator.generate(
        BCE=["Reflection"], BCW=["Reflection"], BCN=["Reflection"], BCS=["Reflection"],
        east=WS_BLOCK_LEN * blocks_per_row, west=0.0, north=WS_BLOCK_LEN * n_rows, south=0.0,
        n_blocks_horizontal=blocks_per_row, n_blocks_vertical=n_rows,
    ).dict


def ws_ic(x, y, width, height):
    """Explosion box over the central 40 % of the domain (explosion_multi states), conservative, non-dimensional
    (examples/explosion/initial_condition.py:35-60)."""
    inside = (x >= 0.3 * width) & (x <= 0.7 * width) & (y >= 0.3 * height) & (y <= 0.7 * height)

    def cons(rho, p):
        e = p / (GAMMA - 1) + 0.0
        return np.array([rho / 1.0, 0.0, 0.0, e / (1.0 * A_INF**2)])

    hi, lo = cons(4.6968, 404400.0), cons(1.1742, 101100.0)
    return np.where(inside[..., None], hi, lo)


def ws_ic_smooth(x, y, width, height):
    """Rounding-robust smooth field (IC-B): every face carries a genuine Riemann problem."""
    rho = 1.2 + 0.3 * np.sin(0.7 * x + 0.3) * np.cos(0.45 * y + 0.1)
    u = 30 * np.cos(0.5 * x) * np.sin(0.35 * y + 0.2)
    v = -25 * np.sin(0.4 * x + 0.5) * np.cos(0.3 * y)
    p = 101325 * (1 + 0.2 * np.cos(0.6 * x - 0.2) * np.sin(0.5 * y + 0.4))
    ek = 0.5 * rho * (u * u + v * v)
    U = np.stack((rho, rho * u, rho * v, p / (GAMMA - 1) + ek), axis=-1)
    return U / np.array([1.0, A_INF, A_INF, A_INF**2])


def ws_ic_1x1(x, y):
    """the weak-scaling initial condition on a one-block domain (named fingerprint `ws2048`)"""
    return ws_ic(x, y, WS_BLOCK_LEN, WS_BLOCK_LEN)

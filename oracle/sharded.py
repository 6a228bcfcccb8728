"""Block-sharded driver of the numpy oracle: the CPU analogue of running the reference with
``mpiexec -n N`` (its block -> rank rule, one ghost-strip exchange per RK stage, global dt by a
min-reduction).  TEST INFRASTRUCTURE: used by tests/test_dist_gloo.py (as a stand-in engine that
keeps the C ABI's halo contract) and by bench.py's CPU arms (as the multi-process CPU baseline).
Never imported by pyhype_b200/.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import muscl_oracle as mo

SIDES = ("E", "W", "N", "S")


class OracleShardEngine:
    """Same method surface as pyhype_b200.engine.Engine for the calls pyhype_b200.distributed makes."""

    def __init__(self, blocks, nx, ny, owner, rank, ic, **kw):
        self.device = 0
        self.prob = mo.Problem(blocks, nx, ny, **kw)
        self.local = sorted(g for g, r in owner.items() if r == rank)
        self.owner, self.rank = owner, rank
        for g in list(self.prob.blocks):
            if g not in self.local:
                del self.prob.blocks[g]
        for b in self.prob.blocks.values():
            b.U = ic(b.geom.xc, b.geom.yc)
        self.num_stages = len(self.prob.tableau)
        self._slots, off = [], 0
        for g in self.local:
            b = self.prob.blocks[g]
            for s in SIDES:
                nb = b.nbr[s]
                if nb is not None and owner[nb] != rank and b.bc[s] is None:
                    ln = 4 * (ny if s in ("E", "W") else nx)
                    self._slots.append(dict(gid=g, side=s, nbr=nb, offset=off, length=ln))
                    off += ln
        self._ndoubles = off

    def halo_slots(self):
        return self._slots, self._ndoubles

    @staticmethod
    def _view(ptr, n):
        return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_double)), shape=(n,))

    def pack_halo(self, ptr):
        buf = self._view(ptr, self._ndoubles)
        for s in self._slots:
            b = self.prob.blocks[s["gid"]]
            buf[s["offset"]:s["offset"] + s["length"]] = b.U[b.geom.edge(s["side"])].reshape(-1)

    def unpack_halo(self, ptr):
        buf = self._view(ptr, self._ndoubles)
        for s in self._slots:
            b = self.prob.blocks[s["gid"]]
            b.ghost[s["side"]] = buf[s["offset"]:s["offset"] + s["length"]].reshape(b.ghost[s["side"]].shape).copy()

    def apply_bc(self):
        p = self.prob
        for b in p.blocks.values():
            for d in SIDES:
                if b.bc[d] is not None or b.nbr[d] is None:
                    b.ghost[d] = b.U[b.geom.edge(d)].copy()
                elif b.nbr[d] in p.blocks:
                    nb = p.blocks[b.nbr[d]]
                    b.ghost[d] = nb.U[nb.geom.edge(mo.OPP[d])].copy()
        for b in p.blocks.values():
            for d in SIDES:
                p._bc_func(b, d, b.ghost[d], conservative=True)

    def local_dt(self, ptr):
        self._view(ptr, 1)[0] = min(self.prob.block_dt(b) for b in self.prob.blocks.values())

    def step_begin(self, dt):
        self._dt = dt
        self._U0 = {g: b.U.copy() for g, b in self.prob.blocks.items()}
        self._R = {g: [] for g in self.prob.blocks}

    def step_begin_dev(self, ptr):
        self.step_begin(float(self._view(ptr, 1)[0]))

    def stage(self, s):
        a = self.prob.tableau
        for g, b in self.prob.blocks.items():
            self._R[g].append(self.prob.residual(b))
            x = self._U0[g]
            for k in range(s + 1):
                if a[s][k] != 0:
                    x = x + (self._dt * a[s][k]) * self._R[g][k]
            b.U = x if x is not self._U0[g] else x.copy()

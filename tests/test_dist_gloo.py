"""-m "not gpu": the N>1 path on CPU -- world_size 2, gloo backend.  The host-side sharding logic
(block->rank map, exchange plan, message ordering, stage-wise driving sequence of
``pyhype_b200.distributed``) is exercised with a stand-in engine that keeps the C ABI's halo
contract (slot order, (edge_len, 4) strips) and computes with the oracle, and the sharded result is
required to be bit-identical to the single-process one (sharding must not change bits)."""
import ctypes
import os
import socket

import numpy as np
import pytest

import cases
from oracle import muscl_oracle as mo
from pyhype_b200.distributed import HaloExchanger, advance, distribute_blocks

SIDES = ("E", "W", "N", "S")


class OracleShardEngine:
    """Same method surface as pyhype_b200.engine.Engine for the calls distributed.py makes."""

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


def _worker(rank, world, port, nsteps, ret):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    blocks = cases.em_mesh()
    owner = distribute_blocks(len(blocks), world)
    eng = OracleShardEngine(blocks, 12, 10, owner, rank, cases.explosion_ic)
    hx = HaloExchanger(eng, owner, rank, backend_device=torch.device("cpu"))
    hx.exchange()
    eng.apply_bc()
    dts = []
    for _ in range(nsteps):
        dt = hx.global_dt()
        dts.append(float(dt[0]))
        advance(eng, hx, eng.num_stages, dt_dev_ptr=dt.data_ptr())
    ret[rank] = dict(dts=dts, U={g: b.U for g, b in eng.prob.blocks.items()},
                     ghost={g: b.ghost for g, b in eng.prob.blocks.items()})
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(300)
def test_two_rank_sharded_run_is_bit_identical_to_single_process():
    import torch.multiprocessing as mp

    nsteps = 4
    ref = cases.build_oracle(cases.em_mesh(), 12, 10, cases.explosion_ic)
    _, dts = ref.run(0.0, 1e9, max_steps=nsteps)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), nsteps, ret), nprocs=2, join=True)
    assert set(ret.keys()) == {0, 1}
    for r in (0, 1):
        assert ret[r]["dts"] == dts
        for g, U in ret[r]["U"].items():
            assert np.array_equal(U, ref.blocks[g].U), (r, g)
            for s in SIDES:
                assert np.array_equal(ret[r]["ghost"][g][s], ref.blocks[g].ghost[s]), (r, g, s)
    assert sorted(list(ret[0]["U"]) + list(ret[1]["U"])) == list(range(8))

"""-m "not gpu": the N>1 path on CPU -- world_size 2, gloo backend.  The host-side sharding logic
(block->rank map, exchange plan, message ordering, stage-wise driving sequence of
``pyhype_b200.distributed``) is exercised with a stand-in engine that keeps the C ABI's halo
contract (slot order, (edge_len, 4) strips) and computes with the oracle, and the sharded result is
required to be bit-identical to the single-process one (sharding must not change bits)."""
import os
import socket

import numpy as np
import pytest

import cases
from oracle.sharded import OracleShardEngine
from pyhype_b200.distributed import HaloExchanger, advance, distribute_blocks

SIDES = ("E", "W", "N", "S")


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

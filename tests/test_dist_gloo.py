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


def _case(name):
    """(blocks, nx, ny, ic, scheme) of the two sharded cases: the regular 2 x 4 explosion grid, and the
    ten-block forward-step topology whose neighbour numbering is not a grid (examples/supersonic_step/mesh.py)
    with its Dirichlet inlet, slip walls and the HLLL / primitive / RK2 scheme."""
    if name == "explosion_multi":
        return cases.em_mesh(), 12, 10, cases.explosion_ic, {}
    return cases.step_mesh(6), 9, 6, cases.step_ic, dict(flux="HLLL", recon="primitive", integrator="RK2", CFL=0.3)


def _worker(rank, world, port, nsteps, case, ret):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    blocks, nx, ny, ic, scheme = _case(case)
    owner = distribute_blocks(len(blocks), world)
    eng = OracleShardEngine(blocks, nx, ny, owner, rank, ic, **scheme)
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
@pytest.mark.parametrize("case", ["explosion_multi", "supersonic_step"])
def test_two_rank_sharded_run_is_bit_identical_to_single_process(case):
    import torch.multiprocessing as mp

    nsteps = 4
    blocks, nx, ny, ic, scheme = _case(case)
    ref = cases.build_oracle(blocks, nx, ny, ic, **scheme)
    _, dts = ref.run(0.0, 1e9, max_steps=nsteps)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), nsteps, case, ret), nprocs=2, join=True)
    assert set(ret.keys()) == {0, 1}
    for r in (0, 1):
        assert ret[r]["dts"] == dts
        for g, U in ret[r]["U"].items():
            assert np.array_equal(U, ref.blocks[g].U), (r, g)
            for s in SIDES:
                assert np.array_equal(ret[r]["ghost"][g][s], ref.blocks[g].ghost[s]), (r, g, s)
    assert sorted(list(ret[0]["U"]) + list(ret[1]["U"])) == list(range(len(blocks)))

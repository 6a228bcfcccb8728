"""Launched by tests/test_gpu_multirank.py under torch.distributed.run (one rank per GPU): the
block-sharded run through the Python facade must be bit-identical to the single-process oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from test_gpu_facade import ExplosionInitialCondition, em_config, em_mesh  # noqa: E402


def main():
    import torch
    import torch.distributed as dist

    from pyhype_b200.solvers import Euler2D

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    nx, ny = int(os.environ.get("PYH_TEST_NX", "26")), int(os.environ.get("PYH_TEST_NY", "22"))
    integ = os.environ.get("PYH_TEST_INTEGRATOR", "RK4")
    cfl = 0.3 if integ == "ExplicitEuler1" else 0.7
    config = em_config(nx=nx, ny=ny, t_final=0.004, initial_condition=ExplosionInitialCondition(), time_integrator=integ, CFL=cfl)
    # PYH_TEST_LAYOUT = "NBXxNBY": default 2x4 (rank boundaries between block rows); "2x1" with 2 ranks puts the rank
    # boundary on an east / west edge
    nbx, nby = (int(v) for v in os.environ.get("PYH_TEST_LAYOUT", "2x4").split("x"))
    north = float(os.environ.get("PYH_TEST_NORTH", "20.0"))   # 10.0 with a 2x2 layout puts the explosion box across BOTH interfaces
    mesh = em_mesh() if (nbx, nby) == (2, 4) else cases.em_mesh(nbx=nbx, nby=nby, north=north)
    sim = Euler2D(config=config, mesh_config=mesh)
    if os.environ.get("PYH_TEST_MODE") == "step":
        # the reference's loop body one call at a time (Euler2D.py:199-210): collective get_dt / integrate / realizability check
        sim._pre_process_solve()
        while sim.t < sim.t_final:
            sim.step()
    else:
        sim.solve()
    prob = cases.build_oracle(mesh.dict if hasattr(mesh, "dict") else mesh, nx, ny, cases.explosion_ic, integrator=integ, CFL=cfl)
    t, dts = prob.run(0.0, 0.004 * 343.0)
    assert sim.num_time_step == len(dts) and sim.t == t, (sim.num_time_step, len(dts))
    mine = [b.global_block_num for b in sim.blocks]
    assert len(mine) == (nbx * nby) // world or world > nbx * nby
    if world > nbx * nby:
        assert len(mine) == (1 if rank < nbx * nby else 0)
    for block in sim.blocks:
        ref = prob.blocks[block.global_block_num]
        assert np.array_equal(block.state.data, ref.U), f"rank {rank} block {block.global_block_num}"
        for s in "EWNS":
            assert np.array_equal(getattr(block.ghost, s).state.data, ref.ghost[s]), (rank, block.global_block_num, s)
    dist.barrier()
    if rank == 0:
        print(f"MULTIRANK OK world={world} steps={len(dts)}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

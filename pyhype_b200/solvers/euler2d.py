"""Euler2D -- the drop-in seam.  Same constructor, attributes and ``solve()`` as
``pyhype.solvers.Euler2D`` (pyhype/solvers/Euler2D.py:45-216, pyhype/solvers/base.py:40-172); the
three calls its time loop makes per step (``get_dt``, ``integrate``, realizability check) and the
ghost refresh run on the GPU through the C ABI of ``include/pyh_b200.h``."""
from __future__ import annotations

import logging
import pathlib
from datetime import datetime

import numpy as np

from ..blocks import SIDES, QuadBlock
from ..boundary_conditions.base import BoundaryCondition, PrimitiveDirichletBC
from ..distributed import distribute_blocks, share_unique_id, world
from ..engine import FLUX_IDS, LIMITER_IDS, Engine
from ..mesh.base import MeshGenerator
from ..states import ConservativeState, PrimitiveState, RealizabilityException
from ..time_marching import get_tableau


class _Logger:
    def __init__(self, config, rank):
        self._log = logging.getLogger("pyhype_b200")
        procs = config.show_log_for_procs
        self._on = procs == "all" or rank in procs

    def info(self, msg):
        if self._on:
            self._log.info(msg)

    def error(self, msg):
        self._log.error(msg)


class _AsyncSolutionWriter:
    """``every_n_timesteps`` dumps that do not stall the device loop (solvers/base.py:158-172 layout):
    the D2H copies run on the engine's copy-out stream into page-locked buffers while the next steps
    compute, and ``np.save`` runs on a worker thread.  Two buffer sets; a third dump waits for the
    first to reach the disk."""

    def __init__(self, engine, gids, depth=2):
        import queue
        import threading

        self._engine = engine
        self._sets = [{g: engine.pinned_state_buffer() for g in gids} for _ in range(depth)]
        self._free = queue.Queue()
        for i in range(depth):
            self._free.put(i)
        self._jobs = queue.Queue()
        self._error = None
        self._thread = threading.Thread(target=self._work, name="pyhype_b200-writer", daemon=True)
        self._thread.start()

    def submit(self, files):
        """files: {gid: path}; returns as soon as the copies are enqueued."""
        self._raise_if_failed()
        i = self._free.get()
        for g in files:
            self._engine.download_async(g, self._sets[i][g])
        self._jobs.put((i, dict(files)))

    def _work(self):
        while True:
            job = self._jobs.get()
            if job is None:
                return
            i, files = job
            try:
                self._engine.downloads_sync()
                for g, f in files.items():
                    np.save(file=f, arr=self._sets[i][g])
            except Exception as e:  # surfaced on the solver thread by the next submit / close
                self._error = e
            finally:
                self._free.put(i)

    def _raise_if_failed(self):
        if self._error is not None:
            e, self._error = self._error, None
            raise e

    def close(self):
        if self._thread is not None:
            self._jobs.put(None)
            self._thread.join()
            self._thread = None
        self._raise_if_failed()


def _validate(config):
    """Reject what the reference rejects, with the same exception types."""
    if config.fvm_type != "MUSCL":
        raise ValueError("Specified finite volume method has not been specialized.")
    if config.fvm_spatial_order != 2:
        # order 1 never reaches the FVM in the reference: GradientsFactory only knows orders 2 and 4
        # (pyhype/blocks/base.py:369-376)
        raise ValueError(
            "GradientsFactory.create_gradients(): Error, no gradients container class has been "
            "extended for the given order."
        )
    if config.nghost != 1:
        raise ValueError("Number of ghost cells must be equal to 1 for this method.")  # SecondOrderMUSCL.py:49-52
    if config.fvm_gradient_type != "GreenGauss":
        raise ValueError("Gradient type not specified.")
    if config.fvm_flux_function_type not in FLUX_IDS:
        raise ValueError("Flux function type not specified.")
    if config.fvm_slope_limiter_type not in LIMITER_IDS:
        raise ValueError("Slope limiter type not specified.")
    if config.interface_interpolation != "arithmetic_average":
        raise ValueError("Interface Interpolation method is not defined.")  # quad_block.py:144-152
    if config.reconstruction_type not in (ConservativeState, PrimitiveState):
        raise ValueError("reconstruction_type must be ConservativeState or PrimitiveState")
    if config.fvm_num_quadrature_points not in (1, 2, 3):
        raise KeyError(config.fvm_num_quadrature_points)


class Euler2D:
    def __init__(self, config, mesh_config, device=None) -> None:
        self.config = config
        self.fluid = config.fluid
        self.cpu, self._world, local_rank = world()
        self._logger = _Logger(config, self.cpu)
        _validate(config)
        tableau = get_tableau(config.time_integrator)
        mesh_info = mesh_config.dict if isinstance(mesh_config, MeshGenerator) else mesh_config
        self.mesh_config = mesh_info

        self.t = 0
        self.num_time_step = 0
        self.CFL = config.CFL
        self.t_final = config.t_final * self.fluid.far_field.a  # solvers/base.py:73
        self.profile_data = None
        self.write_path = None
        if config.write_solution:
            self.write_path = pathlib.Path(config.write_solution_base) / config.write_solution_name
            if self.cpu == 0:
                self.write_path.mkdir(exist_ok=True, parents=True)

        # block -> rank map; block ids must be 0..N-1 (pyhype/blocks/base.py:515-534)
        owner = distribute_blocks(len(mesh_info), self._world)
        mine = [g for g, r in owner.items() if r == self.cpu]
        self._owner = owner
        self._blocks = {}
        inputs = {g: mesh_info[g] for g in mine}  # KeyError like the reference if ids are not 0..N-1
        if len(mine) > 1 and config.nx * config.ny >= 1 << 18:
            # host geometry (linspace, arctan, arccos: 1.5 s per 2048^2 block) is numpy-bound and releases the GIL
            from concurrent.futures import ThreadPoolExecutor

            with ThreadPoolExecutor(max_workers=min(8, len(mine))) as pool:
                built = list(pool.map(lambda g: QuadBlock(config, inputs[g], self), mine))
            self._blocks = dict(zip(mine, built))
        else:
            for g in mine:
                self._blocks[g] = QuadBlock(config, inputs[g], self)

        recon = "primitive" if config.reconstruction_type is PrimitiveState else "conservative"
        self._device = local_rank if device is None else device
        self._engine = Engine(
            config.nx, config.ny, config.fvm_flux_function_type, config.fvm_slope_limiter_type, recon, tableau,
            self.fluid.gamma(), config.CFL, device=self._device,
            num_quadrature_points=config.fvm_num_quadrature_points,
        )
        local = set(mine)
        for g, blk in self._blocks.items():
            bcs = {}
            for s in SIDES:
                bc = blk.info.bc[s]
                if isinstance(bc, PrimitiveDirichletBC):
                    inlet = bc.primitive_state.data
                    want = (config.ny, 1, 4) if s in ("E", "W") else (1, config.nx, 4)
                    if inlet.shape != want:  # states/converter/state_converter.py:60-63
                        raise ValueError(
                            f"States must have equal shape, but state has {want} and from_state has {inlet.shape}"
                        )
                    bcs[s] = np.ascontiguousarray(inlet, dtype=np.float64)
                elif isinstance(bc, BoundaryCondition):
                    raise ValueError("Boundary Condition type " + str(bc) + " has not been specialized.")
                else:
                    bcs[s] = bc
            self._engine.add_block(g, blk.mesh, blk.info.neighbors, bcs, local_gids=local)
            blk._attach(self._engine)
        self._engine.finalize()
        self._writer = None
        self._in_solve = False
        if self._world > 1:
            # the transport (NCCL strip exchange + dt all-reduce) lives in the C layer; torch.distributed only carries the
            # 128-byte communicator id (mpi4py's role in the reference: blocks/ghost.py:169-241, solvers/base.py:128-131)
            uid = share_unique_id(Engine.comm_unique_id, self.cpu, self._world)
            self._engine.comm_init(self.cpu, self._world, uid, owner)
        self._logger.info("\n\tFinished setting up solver")

    def __str__(self):
        c = self.config
        return (
            "\tA Solver of type Euler2D for solving the 2D Euler\n\tequations on structured grids using the Finite Volume Method.\n\n"
            f"\t{'Finite Volume Method: ':<40} {c.fvm_type}\n\t{'Gradient Method: ':<40} {c.fvm_gradient_type}\n"
            f"\t{'Flux Function: ':<40} {c.fvm_flux_function_type}\n\t{'Limiter: ':<40} {c.fvm_slope_limiter_type}\n"
            f"\t{'Time Integrator: ':<40} {c.time_integrator}"
        )

    # -- reference surface -----------------------------------------------------------------------------
    @property
    def blocks(self):
        return self._blocks.values()

    def apply_initial_condition(self):
        for block in self.blocks:
            self.config.initial_condition.apply_to_block(block)

    def apply_boundary_condition(self):
        self._flush_host_states()
        self._refresh_ghosts()

    def get_dt(self) -> float:
        """solvers/base.py:114-136"""
        self._flush_host_states()
        return self._engine.get_dt(self.t, self.t_final)   # collective on >1 rank: global minimum

    def solve(self):
        self._pre_process_solve()
        self._in_solve = True
        try:
            self._solve()
        finally:
            self._in_solve = False
            if self._writer is not None:
                self._writer.close()   # every dump is on disk when solve() returns (or raises)
                self._writer = None
        self._post_process_solve()

    # -- internals ---------------------------------------------------------------------------------------
    def _flush_host_states(self):
        for block in self.blocks:
            block.state.push_if_touched()

    def _mark_device_newer(self):
        for block in self.blocks:
            block.state.mark_device_newer()

    def _refresh_ghosts(self):
        self._engine.apply_bc()   # remote strips (if any) are exchanged inside

    def _pre_process_solve(self):
        self._logger.info("\t>>> Setting Initial Conditions")
        self.apply_initial_condition()
        self._logger.info("\t>>> Setting Boundary Conditions")
        self.apply_boundary_condition()
        if self.config.realplot:
            self._logger.info("\t>>> realplot is not available on the GPU path (matplotlib-free); ignored")
        if self.config.write_solution:
            self.write_mesh()
        self._logger.info(f"Date and time: {datetime.today()}")

    def _post_process_solve(self):
        self._logger.info(
            f"Simulation time: {str(self.t / self.fluid.far_field.a)}, Timestep number: {str(self.num_time_step)}"
        )
        self._logger.info("End of simulation")

    def _update_solution_blocks(self, dt: float) -> None:
        """ExplicitRungeKutta.integrate (time_marching/explicit_runge_kutta.py:47-80)"""
        self._flush_host_states()
        self._engine.step(dt)
        self._mark_device_newer()

    def _realizability_check(self):
        ok = self._engine.realizable()   # reduced over all ranks by the library
        if not ok:
            msg = "ConservativeState has unrealizable values (rho <= 0 or e <= 0)"
            self._logger.error(msg)
            raise RealizabilityException(msg)  # the reference logs and calls MPI.Abort (Euler2D.py:144-152)

    def step(self):
        """One pass of the loop body of Euler2D._solve (Euler2D.py:199-210)."""
        dt = self.get_dt()
        self._update_solution_blocks(dt)
        self._realizability_check()
        if self.config.write_solution:
            self.write_solution()
        self.t += dt
        self.num_time_step += 1
        return dt

    def _solve(self) -> None:
        profiler = None
        if self.config.profile:
            import cProfile

            profiler = cProfile.Profile()
            profiler.enable()
        writes = self.config.write_solution and self.config.write_solution_mode == "every_n_timesteps"
        # device-resident loop between output points: dt, t and the step counter stay on the GPU; on >1 rank the strip
        # exchange and the dt all-reduce are part of the same CUDA graph and every rank sees the same t / step count
        self._flush_host_states()
        while self.t < self.t_final:
            self._log_progress()
            if writes:
                every = self.config.write_every_n_timesteps
                nxt = (every - self.num_time_step % every) + 1 if self.num_time_step % every else 1
            else:
                nxt = -1
            # the reference writes *after* the update of a step whose counter is a multiple of
            # `every`, before the counter is incremented (Euler2D.py:206-210)
            t, n, bad, _ = self._engine.run(self.t, self.t_final, max_steps=nxt, poll_every=50)
            self._mark_device_newer()
            if bad:
                msg = "ConservativeState has unrealizable values (rho <= 0 or e <= 0)"
                self._logger.error(msg)
                raise RealizabilityException(msg)
            if n == 0:
                break
            self.num_time_step += n - 1
            if writes:
                self.write_solution()
            self.num_time_step += 1
            self.t = t
        if profiler is not None:
            import pstats

            profiler.disable()
            self.profile_data = pstats.Stats(profiler)
            if self.cpu == 0:
                self.profile_data.sort_stats("tottime").print_stats(50)

    def _log_progress(self):
        self._logger.info(
            f"Simulation time: {self.t / self.fluid.far_field.a}, Timestep number: {self.num_time_step}"
        )

    # -- npy output, same layout as the reference (solvers/base.py:142-172) ---------------------------------
    @staticmethod
    def write_output_nodes(filename: str, array: np.ndarray):
        np.save(file=filename, arr=array)

    def write_mesh(self):
        current_path = self.write_path / "mesh"
        current_path.mkdir(parents=True, exist_ok=True)
        for block in self.blocks:
            self.write_output_nodes(str(current_path / "mesh_x_blk_") + str(block.global_block_num), block.mesh.x)
            self.write_output_nodes(str(current_path / "mesh_y_blk_") + str(block.global_block_num), block.mesh.y)

    def write_solution(self):
        if (
            self.config.write_solution_mode == "every_n_timesteps"
            and self.num_time_step % self.config.write_every_n_timesteps == 0
        ):
            current_path = self.write_path / str(self.num_time_step)
            current_path.mkdir(parents=True, exist_ok=True)
            self._flush_host_states()
            if self._writer is None:
                self._writer = _AsyncSolutionWriter(self._engine, [b.global_block_num for b in self.blocks])
            self._writer.submit({
                block.global_block_num:
                    str(current_path / self.config.write_solution_name) + "_blk_" + str(block.global_block_num) + ".npy"
                for block in self.blocks
            })
            if not self._in_solve:   # called by user code outside solve(): behave synchronously
                self._writer.close()
                self._writer = None

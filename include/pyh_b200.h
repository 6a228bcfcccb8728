/*
 * pyh_b200.h -- C ABI of the B200-native MUSCL residual + explicit-RK time-march engine.
 *
 * Drop-in boundary for ONE hot path of momokhalil/pyHype.  The reference has no FFI layer; its
 * "operator API" is three calls made by Euler2D._solve() per time step
 * (pyhype/solvers/Euler2D.py:195-210) plus one ghost refresh after the initial condition
 * (pyhype/solvers/Euler2D.py:114).  Each entry point below names the reference interface it
 * replaces.  Plain C: opaque handle, plain pointers and sizes, no torch / C++ types.
 *
 * Conventions
 *   - every function returns 0 on success, a negative pyh_status on failure; the message is
 *     available from pyh_last_error() (thread-local);
 *   - host arrays are row-major, fp64, row index i = south->north, column j = west->east;
 *     "aos" state arrays are (ny, nx, 4) with the last axis [rho, rho*u, rho*v, e], exactly the
 *     reference's block.state.data layout (pyhype/fvm/base.py:210-212);
 *   - side order is always E, W, N, S (pyhype/utils/utils.py:229-237);
 *   - one context drives one GPU (one process per GPU); ranks exchange ghost strips and reduce the time step through
 *     the library's own NCCL communicator (pyh_comm_init) -- or, host-driven, through the pack / unpack entry points;
 *   - a context is not thread-safe; the caller owns host buffers, which are only touched during
 *     the call; device pointers handed in must belong to the context's device.
 */
#ifndef PYH_B200_H
#define PYH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PYH_ABI_VERSION 1
#define PYH_MAX_STAGES 6

typedef enum {
    PYH_OK = 0,
    PYH_ERR_INVALID = -1,   /* bad argument / unsupported option (reference raises ValueError) */
    PYH_ERR_CUDA = -2,      /* CUDA runtime failure */
    PYH_ERR_STATE = -3,     /* call order violated (e.g. add_block after finalize) */
    PYH_ERR_NOMEM = -4
} pyh_status;

/* pyhype/flux/factory.py:30-47 */
typedef enum { PYH_FLUX_ROE = 0, PYH_FLUX_HLLE = 1, PYH_FLUX_HLLL = 2 } pyh_flux;
/* pyhype/limiters/factory.py:28-39 */
typedef enum {
    PYH_LIM_VENKATAKRISHNAN = 0, PYH_LIM_VANLEER = 1, PYH_LIM_VANALBADA = 2, PYH_LIM_BARTHJESPERSEN = 3
} pyh_limiter;
/* SolverConfig.reconstruction_type (pyhype/solver_config.py:88) */
typedef enum { PYH_RECON_CONSERVATIVE = 0, PYH_RECON_PRIMITIVE = 1 } pyh_recon;
/* pyhype/blocks/ghost.py:128-143.  Slipwall is the same function as Reflection (ghost.py:256-278),
 * OutletDirichlet is a zero-gradient copy (ghost.py:250-254). */
typedef enum {
    PYH_BC_NONE = 0, PYH_BC_REFLECTION = 1, PYH_BC_SLIPWALL = 2, PYH_BC_OUTLET_DIRICHLET = 3,
    PYH_BC_PRIMITIVE_DIRICHLET = 4
} pyh_bc;
typedef enum { PYH_EAST = 0, PYH_WEST = 1, PYH_NORTH = 2, PYH_SOUTH = 3 } pyh_side;

/* Replaces SolverConfig's numerical fields (pyhype/solver_config.py:64-135) + the Butcher
 * tableau chosen by TimeIntegratorFactory (pyhype/time_marching/factory.py:29-62). */
typedef struct {
    int32_t abi_version;        /* PYH_ABI_VERSION */
    int32_t device;             /* CUDA device ordinal */
    int32_t nx, ny;             /* cells per block (all blocks share them, as in the reference) */
    int32_t flux;               /* pyh_flux */
    int32_t limiter;            /* pyh_limiter */
    int32_t recon;              /* pyh_recon */
    int32_t num_quadrature_points; /* fvm_num_quadrature_points: 1, 2 or 3 (mesh/quadratures.py:33-37) */
    int32_t num_stages;         /* 1..PYH_MAX_STAGES */
    int32_t reserved;
    double  tableau[PYH_MAX_STAGES * PYH_MAX_STAGES]; /* a[s][k] at [s*PYH_MAX_STAGES+k], k<=s */
    double  gamma;              /* Fluid.gamma() (pyhype/fluids/air.py:27-28) */
    double  cfl;                /* SolverConfig.CFL */
} pyh_config;

/* Replaces QuadBlock.__init__ geometry + BlockInfo connectivity
 * (pyhype/blocks/quad_block.py:243-272, pyhype/blocks/base.py:770-843). */
typedef struct {
    int32_t gid;                /* global block number (BlockInfo.nBLK) */
    int32_t is_cartesian;       /* BaseBlockGhost._is_cartesian() (quad_block.py:96-113) */
    int32_t neighbor[4];        /* global block number per side, -1 = none */
    int32_t neighbor_is_local[4]; /* 1: neighbour block lives in this context; 0: on another rank */
    int32_t bc[4];              /* pyh_bc per side */
    const double* nodes_x;      /* (ny+1, nx+1) node coordinates (mesh/quad_mesh.py:79-99) */
    const double* nodes_y;
    const double* area;         /* (ny, nx) QuadMesh.A: needs host libm (arccos/cos), quad_mesh.py:137-170 */
    const double* cos_v;        /* (ny, nx+1) cos(theta) of vertical (E/W) faces, mesh/base.py:66-78 */
    const double* sin_v;
    const double* cos_h;        /* (ny+1, nx) cos(theta) of horizontal (N/S) faces */
    const double* sin_h;
    const double* dirichlet_prim[4]; /* per side: (edge_len, 4) non-dimensional primitive inlet
                                        state for PYH_BC_PRIMITIVE_DIRICHLET, else NULL
                                        (boundary_conditions/base.py:35-44) */
} pyh_block_desc;

const char* pyh_last_error(void);
int pyh_abi_version(void);

/* Euler2D.__init__ -> Blocks.build (pyhype/blocks/base.py:515-572) */
int pyh_create(const pyh_config* cfg, void** ctx_out);
int pyh_add_block(void* ctx, const pyh_block_desc* blk);
int pyh_finalize(void* ctx);
int pyh_destroy(void* ctx);

/* block.state.data setter / getter (pyhype/states/base.py:84-90) */
int pyh_upload_state(void* ctx, int gid, const double* aos);
int pyh_download_state(void* ctx, int gid, double* aos);
/* Uniform initial state without an upload (initial_conditions/supersonic_flood.py:50-59 assigns a (1, 1, 4)
 * state that the setter broadcasts over the block, states/base.py:99-107): every interior cell of block gid is
 * set to the conservative 4-vector `state`. */
int pyh_fill_uniform(void* ctx, int gid, const double* state);
/* Two-state initial conditions without an upload (examples/explosion_multi/initial_condition.py:53-59,
 * examples/dmr/initial_condition.py:55-58: np.where over a condition on the cell centroids block.mesh.x / .y): every interior
 * cell whose centroid lies in the closed box [x0, x1] x [y0, y1] (bounds may be +-inf) is set to the conservative 4-vector
 * `inside`, every other cell to `outside` -- or left untouched when outside == NULL.  The centroids are the device's
 * bit-identical copy of QuadMesh.x / .y (pyhype/mesh/quad_mesh.py:172-184), so the result equals the uploaded numpy fill. */
int pyh_fill_box(void* ctx, int gid, double x0, double x1, double y0, double y1, const double* inside, const double* outside);
/* Asynchronous variants for streaming use (replaces nothing in the reference, which keeps state on the
 * host; serves Solver.write_solution, pyhype/solvers/base.py:158-172, without stalling the time loop,
 * and back-to-back independent runs).  Host buffers should be page-locked; they are read / written
 * until the matching pyh_transfers_sync returns.
 *   pyh_upload_state_async : enqueue the H2D copy of one block's (ny, nx, 4) state into a device staging
 *                            area on the context's copy-in stream; returns immediately.
 *   pyh_commit_uploads     : make the compute stream wait for the staged copies and convert them into the
 *                            current solution buffers (AoS -> SoA planes); returns immediately.
 *   pyh_download_state_async: convert the current solution of one block into a staging area on the compute
 *                            stream, then copy it to `aos` on the copy-out stream; returns immediately.
 *   pyh_transfers_sync     : block until every copy enqueued so far has completed. */
int pyh_upload_state_async(void* ctx, int gid, const double* aos);
int pyh_commit_uploads(void* ctx);
int pyh_download_state_async(void* ctx, int gid, double* aos);
int pyh_transfers_sync(void* ctx);
/* Block until every pyh_download_state_async copy enqueued so far has landed, WITHOUT waiting for
 * compute enqueued after it (may be called from a second host thread, e.g. an output writer). */
int pyh_downloads_sync(void* ctx);
/* Page-locked host memory for the asynchronous copies (cudaHostAlloc / cudaFreeHost). */
int pyh_host_alloc(size_t bytes, void** out);
int pyh_host_free(void* p);
/* ghost strips as the reference keeps them: out is (edge_len, 4) conservative */
int pyh_download_ghost(void* ctx, int gid, int side, double* out);

/* Blocks.apply_boundary_condition (pyhype/blocks/base.py:448-471): local copies + BC functors.
 * Edges whose neighbour is on another rank are filled by pyh_unpack_halo. */
int pyh_apply_bc(void* ctx);

/* Remote ghost exchange, replaces GhostBlock._send_mpi_buffer / recieve_boundary_data /
 * apply_recv_buffers_to_state (pyhype/blocks/ghost.py:169-241).  The send buffer holds, for every
 * (local block, side) whose neighbour is remote, in (gid, side) ascending order, the block's own
 * edge strip as (edge_len, 4) doubles; the recv buffer has the same layout, holding the strips
 * *received for* those same (block, side) slots. */
int pyh_halo_count(void* ctx, int64_t* n_slots, int64_t* n_doubles);
int pyh_halo_slot(void* ctx, int64_t slot, int32_t* gid, int32_t* side, int32_t* nbr_gid,
                  int64_t* offset_doubles, int64_t* len_doubles);
int pyh_pack_halo(void* ctx, double* dev_send);
int pyh_unpack_halo(void* ctx, const double* dev_recv);

/* Solver.get_dt (pyhype/solvers/base.py:114-136) + QuadBlock.get_dt (quad_block.py:423-436).
 * pyh_local_dt writes CFL * min over this context's blocks (over all ranks after pyh_comm_init) to a device double (no
 * clamp; NULL = a scratch double owned by the context, which pyh_step_begin_dev(ctx, NULL) reads back);
 * pyh_get_dt additionally applies the (t_final - t) clamp and returns it on the host. */
int pyh_local_dt(void* ctx, double* dev_dt_out);
int pyh_get_dt(void* ctx, double t, double t_final, double* dt_out);

/* ExplicitRungeKutta.integrate (pyhype/time_marching/explicit_runge_kutta.py:47-80).
 * Stage-wise entry points so that a multi-rank host can interleave the remote ghost exchange:
 *   pyh_step_begin(dt) ; for s in stages: pyh_stage(s) ; [pack/exchange/unpack] ; pyh_apply_bc()
 * pyh_step(dt) is the single-rank convenience wrapper doing all of it. */
int pyh_step_begin(void* ctx, double dt);
int pyh_step_begin_dev(void* ctx, const double* dev_dt);
int pyh_stage(void* ctx, int stage);
int pyh_step(void* ctx, double dt);
/* Multi-rank transport owned by the library: NCCL over NVLink, one context (= one GPU) per rank.  Replaces the
 * mpi4py calls of the reference: Isend / Irecv per ghost strip (pyhype/blocks/ghost.py:169-241), the Waitall of
 * Blocks.apply_boundary_condition (pyhype/blocks/base.py:454-465) and the gather + bcast of the time step
 * (pyhype/solvers/base.py:128-131).  libnccl.so.2 is bound at run time (dlopen; PYH_NCCL_LIB overrides the name).
 *   pyh_comm_unique_id : rank 0 creates the 128-byte NCCL id; the host distributes it (any side channel).
 *   pyh_comm_init      : collective over all ranks, after pyh_finalize.  owner[g] = rank that owns global block g
 *                        (Blocks.distribute_blocks_to_processes, pyhype/blocks/base.py:473-513).
 * Once initialised, pyh_apply_bc, pyh_step, pyh_run, pyh_local_dt, pyh_get_dt and pyh_realizable are COLLECTIVE:
 * every rank must call them in the same order.  Each ghost refresh packs the edge strips remote neighbours need,
 * exchanges them in ONE grouped ncclSend / ncclRecv batch and unpacks them; the CFL minimum and the realizability flag
 * are reduced by one 16-byte ncclAllReduce(min).  Inside pyh_step / pyh_run the exchange is OVERLAPPED with the residual
 * (the reference posts Isend / Irecv and only then applies the local BCs, blocks/base.py:454-465; here the overlap partner is
 * the stage kernel itself): every stage is launched as thin edge strips -- the first / last rows and column strips of each
 * block, whose results the neighbour ranks need -- on a high-priority side stream, followed there by pack -> NCCL -> unpack,
 * while the interior launch runs on the compute stream; both join before the local ghost copies (PYH_NO_HALO_OVERLAP=1
 * selects the blocking order for diagnostics).  A rank may own no blocks. */
#define PYH_COMM_ID_BYTES 128
int pyh_comm_unique_id(void* id_out);
int pyh_comm_init(void* ctx, int32_t rank, int32_t world, const void* id, const int32_t* owner, int32_t nblocks_total);
int pyh_comm_info(void* ctx, int32_t* rank, int32_t* world, int32_t* n_msgs, int64_t* doubles_per_exchange);

/* Euler2D._solve loop (pyhype/solvers/Euler2D.py:195-210), device resident: dt, t and the step
 * counter stay on the GPU; the host only polls every `poll_every` steps.  dts_out (may be NULL)
 * receives up to dts_cap per-step dt values.  *unrealizable is set when rho<=0 or e<=0 was seen
 * (Euler2D._realizability_check, Euler2D.py:144-152).  One time step (CFL reduction, [dt allreduce,]
 * stages, [strip exchange,] ghost refresh) is captured once as a CUDA graph and replayed.  Contexts with
 * remote neighbours need pyh_comm_init first; the call is then collective. */
int pyh_run(void* ctx, double* t_inout, double t_final, int64_t max_steps, int32_t poll_every,
            int64_t* steps_done, int32_t* unrealizable, double* dts_out, int64_t dts_cap);

/* ConservativeState.realizability_conditions on every local block (states/conservative.py:161-165) */
int pyh_realizable(void* ctx, int32_t* ok_out);

/* Test hooks: block.dUdt() (pyhype/blocks/quad_block.py:512-521) and intermediates. */
int pyh_residual(void* ctx, int gid, double* aos_out);
typedef enum { PYH_DBG_GRAD_X = 0, PYH_DBG_GRAD_Y = 1, PYH_DBG_PHI = 2 } pyh_debug_what;
int pyh_debug_fetch(void* ctx, int gid, int what, double* aos_out);

/* Strip shape the stage kernel runs with: lanes per thread block (lanes - 4 output columns) x rows per strip. */
int pyh_march_shape(void* ctx, int32_t* lanes, int32_t* rows);
/* Which kernels run a stage: 0 = the fused row-marching kernel (above), 1 = the three per-cell / per-face kernels
   (pyh_stage_split.cuh) that an eligible context -- one quadrature point, no neighbours on other ranks, at most a few million
   cells -- uses when they are faster on its blocks.  Decided by measurement at the first pyh_run (tuned_ms: milliseconds per
   stage launch it saw, {fused, split}; zeros if nothing was measured); the environment variable PYH_SPLIT=0/1, read by
   pyh_finalize, forces a path.  Same reference calls, same arithmetic, same bits either way.  tuned_ms may be NULL. */
int pyh_stage_path(void* ctx, int32_t* split, double* tuned_ms);
/* Counters for bench.py: kernels launched by this context since creation. */
int pyh_launch_count(void* ctx, int64_t* n);
/* The CUDA stream all of the context's kernels are launched on (as a cudaStream_t value). */
int pyh_stream(void* ctx, uint64_t* stream_out);
/* Time-stamp helpers so the host can time on the launching stream without torch events. */
int pyh_sync(void* ctx);

#ifdef __cplusplus
}
#endif
#endif /* PYH_B200_H */

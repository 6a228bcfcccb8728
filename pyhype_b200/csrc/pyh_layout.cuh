// pyh_layout.cuh -- HBM data layout shared by all kernels.
//
// Every per-block array is a "plane": (ny + 2) rows x pitch doubles.  Cell (i, j), i in [-1, ny],
// j in [-1, nx] lives at  (i + 1) * pitch + PADL + j ; the frame of index -1 / ny / nx is the
// one-cell ghost layer (blocks/ghost.py keeps it in four separate GhostBlock states).  PADL = 2
// puts interior column 0 on a 16-byte boundary.  Vertical-face arrays use column J in [0, nx] of
// row i, horizontal-face arrays use row I in [0, ny] of column j, through the same formula.
//
// All planes of one block live in ONE slab (one cudaMalloc); a plane is addressed as
// base[plane_offset + cell_offset] with 32-bit element offsets taken from the kernel parameter
// space (PlaneOffsets, same for every block), so a kernel needs one pointer per block and no
// descriptor reloads inside its loops.  A conserved-state buffer is 4 consecutive planes (SoA:
// rho, rho*u, rho*v, e).
#pragma once
#include <stdint.h>
#include "../../include/pyh_b200.h"

namespace pyh {

constexpr int PADL = 2;

struct Layout {
    int nx, ny, pitch;
    unsigned plane;  // doubles per plane
    __host__ __device__ inline unsigned at(int i, int j) const { return (unsigned)((i + 1) * pitch + PADL + j); }
};

// element offsets (plane index * plane size) inside a block slab
struct PlaneOffsets {
    unsigned H[3];                 // haloed conserved-state buffers (4 planes each)
    unsigned P[PYH_MAX_STAGES];    // RK partial-sum accumulators (4 planes each), 0 if unused
    unsigned A;                    // cell area
    unsigned dxy;                  // 8 planes: (x_f - x_c, y_f - y_c) for f = E, W, N, S
    unsigned Lv, cv, sv;           // vertical faces (E/W): length, cos(theta), sin(theta)
    unsigned Lh, ch, sh;           // horizontal faces (N/S)
    unsigned cdx, cdy;             // CFL lengths
    unsigned xc, yc;               // cell centroids (block.mesh.x / .y): device-side initial conditions (pyh_fill_box)
    unsigned nplanes;
};

struct RkTarget {
    unsigned src;  // slab offset of the 4-plane source (U0 buffer or accumulator)
    unsigned dst;  // slab offset of the 4-plane destination (next state buffer or accumulator)
    int add;       // 1: out = src + coef * R ; 0: out = src (a[r][s] == 0 but the row ends here)
    int coef;      // index into the device coefficient table (row * PYH_MAX_STAGES + stage)
};

struct StagePlan {
    int ntargets;
    int write_residual;   // test hook: also store R itself into BlkDev::dbg
    int fuse_dt;          // last stage of a step inside pyh_run: CFL minimum + realizability of the NEW state (the next step's dt)
    int push_ghost;       // the thread that writes an edge cell of the stage's output also writes the ghost cells that mirror it
    unsigned cur;         // slab offset of the state buffer this stage reads
    RkTarget t[PYH_MAX_STAGES];
};

// device-resident control block of the time loop
struct Control {
    double t, t_final, dt;
    double coef[PYH_MAX_STAGES * PYH_MAX_STAGES];  // dt * a[s][k]
    unsigned long long dtmin_bits;                 // running min of dx/(|u|+a), dy/(|v|+a) as ordered bits
    unsigned long long allok;                      // 1 while every state seen is realizable, else 0; directly behind dtmin_bits so
                                                   // that ONE ncclAllReduce(min, uint64, 2) reduces both across ranks (pyh_comm.cuh)
    long long nsteps;
    int active;      // 1 while t < t_final
    int bad;         // unrealizable state seen
    int pending_end; // a step has run whose `t += dt` is still to be applied (k_dt_finalize / k_step_end do it: one kernel per step boundary)
    int pad;
    double* dts;                     // pyh_run: optional per-step dt record (device), dts_cap entries
    long long dts_cap;
};

// doubles as order-preserving 64-bit keys (atomicMin on the CFL minimum; ncclAllReduce(min, uint64) across ranks)
__device__ __forceinline__ unsigned long long dkey(double x) {
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dunkey(unsigned long long k) {
    unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}
constexpr unsigned long long DKEY_INF = 0xfff0000000000000ull;  // dkey(+inf)


struct BlkDev {
    double* base;                 // the block's slab
    double* dbg;                  // 4 planes: residual test hook (allocated on demand)
    double* dbgG;                 // 12 planes: gx[4], gy[4], phi[4] (allocated on demand)
    double* aux;                  // 16 planes: limited face states (E, W, N, S x 4 variables) between the kernels of the split stage (pyh_stage_split.cuh), else null
    double* aux_fx;               // 8 planes: face fluxes (vertical, horizontal x 4) of the split stage; the flux planes of all blocks are contiguous
    const double* dir_recon[4];   // Dirichlet strips in reconstruction variables (edge_len x 4, AoS)
    const double* dir_cons[4];    // Dirichlet strips in conservative variables
    int bc[4];
    int nbr[4];                   // local block index of the neighbour or -1
    int remote_slot[4];           // halo slot (>= 0) when the neighbour is on another rank
    double* send[4];              // where this block's edge strip for a remote neighbour goes (pyh_comm_init), (edge_len, 4) doubles
    int cart;
    int gid;
};

}  // namespace pyh

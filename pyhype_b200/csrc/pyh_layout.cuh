// pyh_layout.cuh -- HBM data layout shared by all kernels.
//
// Every per-block array is a "plane": (ny + 2) rows x pitch doubles.  Cell (i, j), i in [-1, ny],
// j in [-1, nx] lives at  (i + 1) * pitch + PADL + j ; the frame of index -1 / ny / nx is the
// one-cell ghost layer (blocks/ghost.py keeps it in four separate GhostBlock states).  PADL = 2
// puts interior column 0 on a 16-byte boundary.  A conserved-state buffer is 4 consecutive
// planes (SoA: rho, rho*u, rho*v, e).  Vertical-face arrays use column J in [0, nx] of row i,
// horizontal-face arrays use row I in [0, ny] of column j, through the same formula.
#pragma once
#include <stdint.h>
#include "../../include/pyh_b200.h"

namespace pyh {

constexpr int PADL = 2;

struct Layout {
    int nx, ny, pitch;
    long long plane;  // doubles per plane
    __host__ __device__ inline long long at(int i, int j) const { return (long long)(i + 1) * pitch + PADL + j; }
};

struct RkTarget {
    int src;   // 0: U0 buffer, 1: accumulator P[row]
    int dst;   // 0: next haloed state buffer, 1: accumulator P[row], 2: debug buffer (writes R itself)
    int row;   // tableau row s' this target belongs to
    int add;   // 1: out = src + coef * R ; 0: out = src (a[s'][s] == 0 but the row ends here)
    int coef;  // index into the device coefficient table (row * PYH_MAX_STAGES + stage)
};

struct StagePlan {
    int ntargets;
    int cur, next, u0;  // indices into BlkDev::H
    RkTarget t[PYH_MAX_STAGES + 1];
};

// device-resident control block of the time loop
struct Control {
    double t, t_final, dt;
    double coef[PYH_MAX_STAGES * PYH_MAX_STAGES];  // dt * a[s][k]
    unsigned long long dtmin_bits;                 // running min of dx/(|u|+a), dy/(|v|+a) as ordered bits
    long long nsteps;
    int active;      // 1 while t < t_final
    int bad;         // unrealizable state seen
    int pad[2];
};

struct BlkDev {
    double* H[3];                 // haloed conserved-state buffers (4 planes each)
    double* P[PYH_MAX_STAGES];    // RK partial-sum accumulators (4 planes each) or nullptr
    double* dbg;                  // 4 planes scratch for test hooks
    double* dbgG;                 // 12 planes: gx[4], gy[4], phi[4] (allocated on demand)
    const double* A;              // cell area
    const double* dxy;            // 8 planes: (x_f - x_c, y_f - y_c) for f = E, W, N, S
    const double* Lv; const double* cv; const double* sv;   // vertical faces (E/W)
    const double* Lh; const double* ch; const double* sh;   // horizontal faces (N/S)
    const double* cdx; const double* cdy;                   // CFL lengths
    const double* dir_recon[4];   // Dirichlet strips in reconstruction variables (edge_len x 4, AoS)
    const double* dir_cons[4];    // Dirichlet strips in conservative variables
    int bc[4];
    int nbr[4];                   // local block index of the neighbour or -1
    int remote_slot[4];           // halo slot (>= 0) when the neighbour is on another rank
    int cart;
    int gid;
};

}  // namespace pyh

// pyh_plan.cuh -- host-side bookkeeping of one explicit Runge-Kutta step: slab layout, buffer roles and the per-stage
// plan the stage kernel executes.  Pure functions (no CUDA calls) shared by pyh_api.cu and by the CPU twin of the kernel
// (tests/host_twin/kernel_twin.cpp), so that the twin drives the kernel through exactly the product's plan logic.
//
// Reference: ExplicitRungeKutta.integrate / _update_state (time_marching/explicit_runge_kutta.py:47-89): stage s
// evaluates R_s on U_s and forms every later stage state as U0 + sum_k (dt a[r][k]) R_k, left to right in k.  Here each
// row r that still needs earlier residuals keeps a running sum P_r, updated by stage s only where a[r][s] != 0 -- the
// same additions in the same order.
#pragma once
#include <string.h>
#include "pyh_layout.cuh"
#include "pyh_march_tu.cuh"

namespace pyh {

// row r needs an accumulator iff some a[r][k], k < r, is non-zero before its own stage
inline void plan_need_acc(const double* a /* PYH_MAX_STAGES^2, row-major */, int S, bool need_acc[PYH_MAX_STAGES]) {
    for (int r = 0; r < PYH_MAX_STAGES; ++r) {
        need_acc[r] = false;
        for (int k = 0; k < r && r < S; ++k)
            if (a[r * PYH_MAX_STAGES + k] != 0.0) need_acc[r] = true;
    }
}

// plane indices -> element offsets inside a block slab
inline PlaneOffsets plan_offsets(unsigned plane, int S, int nq, const bool need_acc[PYH_MAX_STAGES]) {
    PlaneOffsets po;
    memset(&po, 0, sizeof(po));
    unsigned n = 0;
    const int nH = S >= 3 ? 3 : 2;
    for (int h = 0; h < 3; ++h) { po.H[h] = (h < nH) ? n * plane : 0; if (h < nH) n += 4; }
    for (int r = 0; r < S; ++r) if (need_acc[r]) { po.P[r] = n * plane; n += 4; }
    po.A = n++ * plane;
    po.dxy = n * plane; n += 8 * nq;
    po.Lv = n++ * plane; po.cv = n++ * plane; po.sv = n++ * plane;
    po.Lh = n++ * plane; po.ch = n++ * plane; po.sh = n++ * plane;
    po.cdx = n++ * plane; po.cdy = n++ * plane;
    po.xc = n++ * plane; po.yc = n++ * plane;
    po.nplanes = n;
    return po;
}

// buffer that stage s writes its stage state to, given the buffer it reads (cur) and the roles (i0 = solution)
inline int plan_next_buffer(int S, int s, int cur, int i0, int i1, int i2) {
    if (s == S - 1) return (S == 1) ? i1 : i0;
    return (cur == i1) ? i2 : i1;
}

// RK partial-sum plan for stage s
inline StagePlan plan_stage(const double* a, int S, const PlaneOffsets& po, int i0, int s, int cur, int next) {
    StagePlan p;
    memset(&p, 0, sizeof(p));
    p.cur = po.H[cur];
    for (int r = s; r < S; ++r) {
        bool prior = false;
        for (int k = 0; k < s; ++k) if (a[r * PYH_MAX_STAGES + k] != 0.0) prior = true;
        const bool nz = a[r * PYH_MAX_STAGES + s] != 0.0;
        if (r != s && !nz) continue;
        RkTarget t;
        t.src = prior ? po.P[r] : po.H[i0];
        t.dst = (r == s) ? po.H[next] : po.P[r];
        t.add = nz ? 1 : 0;
        t.coef = r * PYH_MAX_STAGES + s;
        p.t[p.ntargets++] = t;
    }
    return p;
}

// ---- which strips each launch of one stage covers ------------------------------------------------------------------
struct TileLaunch {
    MarchTiles tiles;
    unsigned gx, gy;   // grid.x, grid.y (grid.z = number of local blocks)
    int tys;           // rows per strip of this launch
    int edge;          // 1: produces cells that remote neighbours need (runs ahead of the strip exchange)
};

// A context without remote neighbours (split_ns == split_ew == false) covers every block with ONE launch of
// ceil(nx / (nt - 4)) x ceil(ny / tys) strips.  With remote neighbours across north / south edges the first and the last
// `th` rows of every block go into an edge launch of thin strips, with remote neighbours across east / west edges so do the
// first and the last column strip of the remaining rows; the interior launch takes the rest.  The launches of one stage
// write disjoint cells and all read the same input buffer, so they may run concurrently and in any order; together they
// cover every cell exactly once (tests/test_kernel_twin.py runs the split on the CPU).  Returns the number of launches.
constexpr int kEdgeRows = 4;
inline int plan_tiles(int nx, int ny, int nt, int tys, bool split_ns, bool split_ew, TileLaunch out[3], int th = kEdgeRows) {
    const int nsx = (nx + nt - 5) / (nt - 4);
    int n = 0;
    // A requested split that the block is too small for (fewer than 2 th + 1 rows / fewer than 3 column strips) must not
    // leave cells a neighbour rank needs in the interior launch: the exchange would pack them before they are written.
    // Then nothing is split and the one launch counts as "edge" (the exchange follows it: the blocking order).
    const bool can_ns = ny >= 2 * th + 1, can_ew = nsx >= 3;
    if ((split_ns && !can_ns) || (split_ew && !can_ew)) split_ns = split_ew = false;
    const bool any_split = split_ns || split_ew;
    const int mid0 = split_ns ? th : 0, mid1 = split_ns ? ny - th : ny;
    if (split_ns) {       // rows [0, th) and [ny - th, ny), every column strip
        TileLaunch t;
        t.tiles = MarchTiles{0, ny - th, ny, 0, 1};
        t.gx = (unsigned)nsx; t.gy = 2; t.edge = 1; t.tys = th;
        out[n++] = t;
    }
    const int gy_mid = (mid1 - mid0 + tys - 1) / tys;
    if (split_ew) {       // column strips 0 and nsx - 1 of the middle rows
        TileLaunch t;
        t.tiles = MarchTiles{mid0, tys, mid1, 0, nsx - 1};
        t.gx = 2; t.gy = (unsigned)gy_mid; t.edge = 1; t.tys = tys;
        out[n++] = t;
    }
    TileLaunch t;         // the rest: the interior -- or, without any split, the whole block (then the exchange waits for it)
    t.tiles = MarchTiles{mid0, tys, mid1, split_ew ? 1 : 0, 1};
    t.gx = (unsigned)(split_ew ? nsx - 2 : nsx); t.gy = (unsigned)gy_mid; t.edge = any_split ? 0 : 1; t.tys = tys;
    out[n++] = t;
    return n;
}

}  // namespace pyh

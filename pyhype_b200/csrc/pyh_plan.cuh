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

}  // namespace pyh

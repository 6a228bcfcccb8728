// pyh_stage_split.cuh -- the RK stage as THREE kernels (sm_100a, fp64), one thread per cell / face: for problems the fused kernel's strips fit badly.
//
// Same reference path and the same arithmetic, operation for operation, as the fused row-marching kernel of
// pyh_stage_march.cuh (fvm/base.py:108-500, fvm/SecondOrderMUSCL.py, limiters/base.py, gradients/greengauss.py,
// blocks/quad_block.py:120-218, time_marching/explicit_runge_kutta.py:63-89) -- only the decomposition differs.
//
// Why.  The fused kernel keeps a cell's face states and fluxes on chip by marching row strips; a strip pays for two extra
// rows of gradient + limiter and one extra face solve at its boundaries, and a thread block lives for (rows + 1) row times.
// On the reference's own examples (explosion_multi: 8 x 150^2 = 180 k cells, DMR: 4 x 500^2) that means strips of 3-16
// rows, +43 % redundant work and a third of the GPU's warp slots in use (profiles/r02k_em_stage_march_ncu_summary.txt).
// Here the limited face states and the face fluxes (192 B per cell) pass between three kernels through global memory --
// L2-resident at the sizes this path is meant for --, which removes every redundant row and exposes one thread per cell / face:
//
//     k_split_recon  : cell (i, j)  -> gradient, limiter, the four limited face states        -> FS[face][var]   (16 planes)
//     k_split_flux   : face         -> ghost-side state / BC, rotation, Riemann solve, x L   -> FX[dir][var]    (8 planes)
//     k_split_update : cell (i, j)  -> residual, RK partial sums, CFL minimum, ghost push    -> state buffers
//
// (Measured alternative: flux + update in ONE kernel, 32 x 8 thread tiles that hold all four faces of 31 x 7 cells, fluxes in
// shared memory.  explosion_multi 0.161 instead of 0.177 ms/step, but DMR's HLLL solves -- 122 registers, two per thread in
// sequence, +18 % of them redundant -- 0.514 instead of 0.434: profiles/r02p_split_stage_merged_ab.txt.  Not kept.)
//
// Whether this or the fused kernel is faster depends on how the block shape fits the fused kernel's strips, so eligible
// contexts measure both at the first pyh_run (pyh_api.cu: tune_stage_path); PYH_SPLIT=0/1 forces a path.  One quadrature
// point only (every shipped example); 2 / 3 points always take the fused kernel.
#pragma once
#include "pyh_layout.cuh"
#include "pyh_math.cuh"
#include "pyh_march_tu.cuh"
#include "pyh_stage_march.cuh"   // PYH_RO, push_ghost_cell, FastTag / SafeTag

namespace pyh {

// Programmatic dependent launch (sm_90+): the three kernels of a stage -- and the stages of a step -- form a chain in which each
// kernel needs ALL of its predecessor's output.  Launched with cudaLaunchAttributeProgrammaticStreamSerialization, a kernel's
// thread blocks may be scheduled while the predecessor's last wave is still running; pdl_wait() then blocks until the
// predecessor has completed and its writes are visible (everything older is complete transitively: each kernel waits first
// thing, before it touches memory, so there is no write-after-read hazard either), and pdl_trigger() lets the successor start
// its own launch.  What overlaps is launch latency and block scheduling: 0.8 us per boundary measured, 13 boundaries per RK4
// step (explosion_multi 0.168 -> 0.158 ms/step).  Without the launch attribute both are no-ops.
__device__ __forceinline__ void pdl_wait() {
#if !defined(PYH_HOST_TWIN)
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_trigger() {
#if !defined(PYH_HOST_TWIN)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

__device__ __forceinline__ bool split_cell_exists(const Layout& lay, int i, int j) {   // interior or ghost frame without corners
    return (i >= -1) && (i <= lay.ny) && (j >= -1) && (j <= lay.nx) && !((i == -1 || i == lay.ny) && (j == -1 || j == lay.nx));
}

// BaseBlockGhost.from_block (quad_block.py:120-134): conservative -> reconstruction variables
template <int PRIM>
__device__ __forceinline__ void split_to_recon(double q[4], const Consts& C) {
    if (PRIM) {
        double s[4] = {q[0], q[1], q[2], q[3]};
        bool ok = true;
        cons2prim<true>(q, C, ok);
        if (!ok) { q[0] = s[0]; q[1] = s[1]; q[2] = s[2]; q[3] = s[3]; cons2prim<false>(q, C, ok); }
    }
}

// ---- 1: gradient + limiter + limited face states ------------------------------------------------------------------------
// (Requesting the cell's 21 geometry values BEFORE the tile of states is staged, so that their latency overlaps the staging and
// the barrier, was measured and lost: explosion_multi 0.1726 vs 0.1676 ms/step, profiles/r02p_split_stage_merged_ab.txt.)
template <int LIM, int PRIM, int MINB>
__global__ void __launch_bounds__(kSplitReconThreads, MINB)
k_split_recon(const BlkDev* __restrict__ blks, const Layout lay, const PlaneOffsets po, const unsigned cur, const Control* __restrict__ ctl,
              const Consts C) {
    constexpr int TX = kSplitTX, TY = kSplitTY, SX = TX + 2, SY = TY + 2;
    __shared__ double sq[4][SY][SX];                        // reconstruction variables of the tile and its one-cell frame
    const BlkDev& B = blks[blockIdx.z];                     // the block table is constant for the life of the context: read ahead of the wait
    const double* __restrict__ const U = B.base + cur;
    const double* __restrict__ const G = B.base;
    double* __restrict__ const FS = B.aux;
    pdl_wait();
    pdl_trigger();
    if (!ctl->active) return;
    const int nx = lay.nx, ny = lay.ny, pitch = lay.pitch;
    const unsigned PL = lay.plane;
    const int t = threadIdx.x;
    const int j0 = (int)blockIdx.x * TX, i0 = (int)blockIdx.y * TY;
    const int tx = t % TX, ty = t / TX;
    const int i = i0 + ty, j = j0 + tx;
    const bool mine = (i < ny) && (j < nx);
    const unsigned o = lay.at(mine ? i : 0, mine ? j : 0);
    const unsigned oE = o + 1, oN = o + pitch;
    double gL[4], gc[4], gs[4], Acell, dx[4], dy[4];
    auto load_geometry = [&]() {
        gL[0] = PYH_RO(G[po.Lv + oE]); gL[1] = PYH_RO(G[po.Lv + o]); gL[2] = PYH_RO(G[po.Lh + oN]); gL[3] = PYH_RO(G[po.Lh + o]);
        gc[0] = PYH_RO(G[po.cv + oE]); gc[1] = PYH_RO(G[po.cv + o]); gc[2] = PYH_RO(G[po.ch + oN]); gc[3] = PYH_RO(G[po.ch + o]);
        gs[0] = PYH_RO(G[po.sv + oE]); gs[1] = PYH_RO(G[po.sv + o]); gs[2] = PYH_RO(G[po.sh + oN]); gs[3] = PYH_RO(G[po.sh + o]);
        Acell = PYH_RO(G[po.A + o]);
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            dx[f] = PYH_RO(G[po.dxy + (f * 2) * PL + o]);
            dy[f] = PYH_RO(G[po.dxy + (f * 2 + 1) * PL + o]);
        }
    };
    for (int e = t; e < SX * SY; e += kSplitReconThreads) {
        const int li = e / SX, lj = e - li * SX;
        const int ci = i0 - 1 + li, cj = j0 - 1 + lj;
        double q[4] = {1.0, 0.0, 0.0, 1.0};
        if (split_cell_exists(lay, ci, cj)) {
            const unsigned oc = lay.at(ci, cj);
            q[0] = PYH_RO(U[oc]); q[1] = PYH_RO(U[oc + PL]); q[2] = PYH_RO(U[oc + 2 * PL]); q[3] = PYH_RO(U[oc + 3 * PL]);
            split_to_recon<PRIM>(q, C);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) sq[k][li][lj] = q[k];
    }
    __syncthreads();
    if (!mine) return;
    load_geometry();
    // GreenGauss._get_gradinet_JIT (gradients/greengauss.py:110-155); same expressions as phase B of k_stage_march
    double LE = gL[0], LW = gL[1], LN = gL[2], LS = gL[3];
#if PYH_FOLD_POW2
    LE = 0.5 * LE; LW = 0.5 * LW; LN = 0.5 * LN; LS = 0.5 * LS;
#endif
    const double xlE = LE * gc[0], xlW = LW * (-gc[1]);
    const double xlN = LN * gc[2], xlS = LS * (-gc[3]);
    const double ylE = LE * gs[0], ylW = LW * (-gs[1]);
    const double ylN = LN * gs[2], ylS = LS * (-gs[3]);
    bool okA = true;
    double ia = Ar<true>::rcp(Acell, okA);
    if (!okA) ia = 1.0 / Acell;
#pragma unroll 2
    for (int k = 0; k < 4; ++k) {
        const double q = sq[k][ty + 1][tx + 1], qW = sq[k][ty + 1][tx], qE = sq[k][ty + 1][tx + 2];
        const double qS = sq[k][ty][tx + 1], qN = sq[k][ty + 2][tx + 1];
        // face averages (quad_block.py:181-218)
#if PYH_FOLD_POW2
        const double fE = q + qE, fW = qW + q, fN = q + qN, fS = qS + q;
#else
        const double fE = 0.5 * (q + qE), fW = 0.5 * (qW + q), fN = 0.5 * (q + qN), fS = 0.5 * (qS + q);
#endif
        const double gx = (fE * xlE + fW * xlW + fN * xlN + fS * xlS) * ia;
        const double gy = (fE * ylE + fW * ylW + fN * ylN + fS * ylS) * ia;
        double term[4], davg[4];
#pragma unroll
        for (int f = 0; f < 4; ++f) term[f] = gx * dx[f] + gy * dy[f];                 // blocks/base.py:283-288
        // SlopeLimiter._get_slope / _limit (limiters/base.py:47-108, 179-187): 5-point min / max by the 6-comparison network
        const bool wge = qW > qE, sgn = qS > qN;
        const double h1 = wge ? qW : qE, l1 = wge ? qE : qW, h2 = sgn ? qS : qN, l2 = sgn ? qN : qS;
        const double mx = dmax2(dmax2(h1, h2), q);
        const double mn = dmin2(dmin2(l1, l2), q);
        const double dmx = mx - q, dmn = mn - q;
#pragma unroll
        for (int f = 0; f < 4; ++f) davg[f] = (q + term[f]) - q;                       // limiters/base.py:99-102
        double phi;
        if (!limiter4_fast<LIM>(dmx, dmn, davg, phi)) limiter4_safe<LIM>(dmx, dmn, davg, phi);
        if (phi < 0.0) phi = 0.0;                                                      // limiters/base.py:187
#pragma unroll
        for (int f = 0; f < 4; ++f) FS[(f * 4 + k) * (size_t)PL + o] = q + phi * term[f];   // SecondOrderMUSCL.py:124-126
    }
}

// ---- 2: one Riemann problem per face ------------------------------------------------------------------------------------
// One face of one cell: ghost-side state / boundary condition, rotation into the face frame, Riemann solve, rotation back, x L.
// horiz == false: the west face of cell (i, j), i in [0, ny), j in [0, nx]; true: the south face of cell (i, j), i in [0, ny], j in [0, nx).
template <int FLUX, int PRIM>
__device__ __forceinline__ void split_face_flux(const BlkDev& B, const double* __restrict__ U, const double* __restrict__ G,
                                                const double* __restrict__ FS, const Layout& lay, const PlaneOffsets& po, const Consts& C,
                                                const bool horiz, const int i, const int j, double out[4]) {
    const int nx = lay.nx, ny = lay.ny, pitch = lay.pitch;
    const unsigned PL = lay.plane;
    const unsigned o = lay.at(i, j);
    const int cart = B.cart & 1;
    const bool vident = cart || (PYH_SKIP_UNIT_ROT && (B.cart & 2));
    auto ghost_state = [&](int gi, int gj, double q[4]) {     // first-order ghost-side state of a `bc None` edge (fvm/base.py:305-325)
        const unsigned og = lay.at(gi, gj);
        q[0] = PYH_RO(U[og]); q[1] = PYH_RO(U[og + PL]); q[2] = PYH_RO(U[og + 2 * PL]); q[3] = PYH_RO(U[og + 3 * PL]);
        split_to_recon<PRIM>(q, C);
    };
    auto face_state = [&](int f, unsigned oc, double q[4]) {
#pragma unroll
        for (int k = 0; k < 4; ++k) q[k] = FS[(f * 4 + k) * (size_t)PL + oc];
    };
    // GhostBlock.apply_boundary_condition_to_state on an edge state (fvm/base.py:352-362)
    auto apply_bc_edge = [&](int bc, int side, int idx, double c_, double s_, double q[4]) {
        if (bc == PYH_BC_REFLECTION || bc == PYH_BC_SLIPWALL) reflect(q[1], q[2], c_, s_);
        else if (bc == PYH_BC_PRIMITIVE_DIRICHLET) {
            const double* d = B.dir_recon[side] + 4 * (long long)idx;
            q[0] = d[0]; q[1] = d[1]; q[2] = d[2]; q[3] = d[3];
        }
    };
    double QL0[4], QR0[4], cf, sf, Lf;
    if (!horiz) {
        const int bcE = B.bc[PYH_EAST], bcW = B.bc[PYH_WEST];
        cf = PYH_RO(G[po.cv + o]); sf = PYH_RO(G[po.sv + o]); Lf = PYH_RO(G[po.Lv + o]);
        if (j > 0) face_state(0, o - 1, QL0);                                           // east-face state of cell (i, j-1)
        else if (bcW == PYH_BC_NONE) ghost_state(i, -1, QL0);
        else { face_state(1, o, QL0); apply_bc_edge(bcW, PYH_WEST, i, cf, sf, QL0); }
        if (j < nx) face_state(1, o, QR0);                                              // west-face state of cell (i, j)
        else if (bcE == PYH_BC_NONE) ghost_state(i, nx, QR0);
        else { face_state(0, o - 1, QR0); apply_bc_edge(bcE, PYH_EAST, i, cf, sf, QR0); }
        if (!vident) { rot(QL0[1], QL0[2], cf, sf); rot(QR0[1], QR0[2], cf, sf); }      // fvm/base.py:366-376
    } else {
        const int bcN = B.bc[PYH_NORTH], bcS = B.bc[PYH_SOUTH];
        cf = PYH_RO(G[po.ch + o]); sf = PYH_RO(G[po.sh + o]); Lf = PYH_RO(G[po.Lh + o]);
        if (i > 0) face_state(2, o - pitch, QL0);                                       // north-face state of cell (i-1, j)
        else if (bcS == PYH_BC_NONE) ghost_state(-1, j, QL0);
        else { face_state(3, o, QL0); apply_bc_edge(bcS, PYH_SOUTH, j, cf, sf, QL0); }
        if (i < ny) face_state(3, o, QR0);                                              // south-face state of cell (i, j)
        else if (bcN == PYH_BC_NONE) ghost_state(ny, j, QR0);
        else { face_state(2, o - pitch, QR0); apply_bc_edge(bcN, PYH_NORTH, j, cf, sf, QR0); }
        if (cart) { rot90(QL0[1], QL0[2]); rot90(QR0[1], QR0[2]); }                     // fvm/base.py:435-441
        else { rot(QL0[1], QL0[2], cf, sf); rot(QR0[1], QR0[2], cf, sf); }
    }
    double Fq[4];
    auto face = [&](auto tag) -> bool {
        constexpr bool FAST = decltype(tag)::value;
        bool ok = true;
        double QL[4] = {QL0[0], QL0[1], QL0[2], QL0[3]}, QR[4] = {QR0[0], QR0[1], QR0[2], QR0[3]};
#if PYH_COLD_SAFE
        if (!FAST) {
            const Flux4 fc = riemann_flux_cold<FLUX, PRIM>(QL[0], QL[1], QL[2], QL[3], QR[0], QR[1], QR[2], QR[3], C);
            Fq[0] = fc.f[0]; Fq[1] = fc.f[1]; Fq[2] = fc.f[2]; Fq[3] = fc.f[3];
        } else
#endif
        riemann_flux<FLUX, PRIM, FAST>(QL, QR, Fq, C, ok);
        return ok;
    };
    if (!face(FastTag{})) face(SafeTag{});
    if (!horiz) { if (!vident) unrot(Fq[1], Fq[2], cf, sf); }                           // fvm/base.py:388-390
    else if (cart) unrot90(Fq[1], Fq[2]);                                               // fvm/base.py:482-486
    else unrot(Fq[1], Fq[2], cf, sf);
    // integrate_flux (fvm/base.py:188-190), one point: L * (0 + 2 F); riemann_flux returns flux_scale(FLUX) * F
    const double Lf1 = (flux_scale(FLUX) == 2.0) ? Lf : 2.0 * Lf;
#pragma unroll
    for (int k = 0; k < 4; ++k) out[k] = PYH_FOLD_POW2 ? Lf1 * Fq[k] : Lf * (2.0 * Fq[k]);
}

// One thread per face.  blockIdx.y = 0: vertical faces (i, J), i in [0, ny), J in [0, nx] (the west face of cell (i, J));
// blockIdx.y = 1: horizontal faces (I, j), I in [0, ny], j in [0, nx) (the south face of cell (I, j)).
template <int FLUX, int PRIM, int MINB>
__global__ void __launch_bounds__(kSplitFluxThreads, MINB)
k_split_flux(const BlkDev* __restrict__ blks, const Layout lay, const PlaneOffsets po, const unsigned cur, const Control* __restrict__ ctl,
             const Consts C) {
    const BlkDev& B = blks[blockIdx.z];                     // constant for the life of the context: read ahead of the wait
    double* const aux_fx = B.aux_fx;
    const double* const aux_fs = B.aux;
    const double* const slab = B.base;
    pdl_wait();
    pdl_trigger();
    if (!ctl->active) return;
    const int nx = lay.nx, ny = lay.ny;
    const bool horiz = blockIdx.y != 0;
    const int W = horiz ? nx : nx + 1;
    const unsigned n = blockIdx.x * (unsigned)kSplitFluxThreads + threadIdx.x;   // faces of one family per block < 2^31
    const int i = (int)(n / (unsigned)W), j = (int)(n - (unsigned)i * (unsigned)W);
    if (i >= (horiz ? ny + 1 : ny)) return;
    double F[4];
    split_face_flux<FLUX, PRIM>(B, slab + cur, slab, aux_fs, lay, po, C, horiz, i, j, F);
    double* __restrict__ const out = aux_fx + (horiz ? 4 : 0) * (size_t)lay.plane + lay.at(i, j);
#pragma unroll
    for (int k = 0; k < 4; ++k) out[k * (size_t)lay.plane] = F[k];
}

// ---- 3: residual + RK partial sums (+ CFL minimum of the new state, ghost push) ------------------------------------------
// One thread per cell.  The kernel is a chain of dependent memory round trips, not arithmetic, so everything is requested as
// early as it can be: the block table (constant for the life of the context) before the wait on the predecessor, every operand
// -- including the time loop's `active` flag and the RK coefficients -- before the first use, and the early exit of an inactive
// step is taken once all of them are in flight.  All cells of explosion_multi are resident at once (5 x 256 threads per SM).
template <int MINB>   // 5: 48 registers, all of explosion_multi in one wave; 4: 64 registers, every operand of a cell in flight at once
__global__ void __launch_bounds__(kSplitUpdateThreads, MINB)
k_split_update(const BlkDev* __restrict__ blks, const Layout lay, const PlaneOffsets po, const StagePlan plan, const Control* __restrict__ ctl,
               Control* __restrict__ ctl_out, const Consts C) {
    __shared__ double sDT[kSplitUpdateThreads / 32];
    const BlkDev& B = blks[blockIdx.z];
    double* __restrict__ const base = B.base;
    const double* __restrict__ const G = B.base;
    const double* __restrict__ const FX = B.aux_fx;
    pdl_wait();
    pdl_trigger();
    const int active = ctl->active;
    const int nx = lay.nx, ny = lay.ny, pitch = lay.pitch;
    const unsigned PL = lay.plane;
    const unsigned n = blockIdx.x * (unsigned)kSplitUpdateThreads + threadIdx.x;
    const int i = (int)(n / (unsigned)nx), j = (int)(n - (unsigned)i * (unsigned)nx);
    const bool live = i < ny;
    const unsigned o = lay.at(live ? i : 0, live ? j : 0);
    const int nt_ = plan.ntargets;
    constexpr double rscale = PYH_FOLD_POW2 ? 0.5 : 1.0;   // Rk == R / rscale
    double tm = __longlong_as_double(0x7ff0000000000000ll);
    double dI[4] = {0.0, 0.0, 0.0, 0.0}, s0[4] = {0.0, 0.0, 0.0, 0.0}, s1[4] = {0.0, 0.0, 0.0, 0.0};
    double cdx = 1.0, cdy = 1.0, c0 = 0.0, c1 = 0.0, a_cell = 1.0;
    if (live) {
#if !defined(PYH_HOST_TWIN)
        if (plan.push_ghost && (i == 0 || i == ny - 1 || j == 0 || j == nx - 1)) {
            // edge cells: what the ghost push at the end will need -- the block's boundary-condition / neighbour table and the
            // cos / sin of the boundary faces (reflection walls) -- into the L1 now
            asm volatile("prefetch.global.L1 [%0];" ::"l"(&B.bc[0]));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(&B.send[0]));
            if (i == 0 || i == ny - 1) {
                const unsigned of = lay.at(i == 0 ? 0 : ny, j);
                asm volatile("prefetch.global.L1 [%0];" ::"l"(G + po.ch + of));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(G + po.sh + of));
            }
            if (j == 0 || j == nx - 1) {
                const unsigned of = lay.at(i, j == 0 ? 0 : nx);
                asm volatile("prefetch.global.L1 [%0];" ::"l"(G + po.cv + of));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(G + po.sv + of));
            }
        }
#endif
        a_cell = PYH_RO(G[po.A + o]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double IWp = FX[k * (size_t)PL + o], IEp = FX[k * (size_t)PL + o + 1];
            const double ISp = FX[(4 + k) * (size_t)PL + o], INp = FX[(4 + k) * (size_t)PL + o + pitch];
            s0[k] = (nt_ > 0) ? base[plan.t[0].src + k * PL + o] : 0.0;
            s1[k] = (nt_ > 1) ? base[plan.t[1].src + k * PL + o] : 0.0;
            dI[k] = IWp - IEp + ISp - INp;
        }
        if (plan.fuse_dt) { cdx = PYH_RO(G[po.cdx + o]); cdy = PYH_RO(G[po.cdy + o]); }
        // dt * a[s][k] of the two static targets, requested with everything else (k_dt_finalize wrote them before this stage began)
        const double k0 = (nt_ > 0) ? ctl->coef[plan.t[0].coef] : 0.0, k1 = (nt_ > 1) ? ctl->coef[plan.t[1].coef] : 0.0;
        c0 = PYH_FOLD_POW2 ? rscale * k0 : k0;
        c1 = PYH_FOLD_POW2 ? rscale * k1 : k1;
    }
    if (!active) return;   // same for every thread of the grid (the step lies beyond t_final, or the run went bad), taken with all loads in flight
    if (live) {
        const double a = a_cell;
        // D: residual (fvm/base.py:141-165), as in k_stage_march
        double Rk[4];
        auto resid = [&](auto tag) -> bool {
            constexpr bool FAST = decltype(tag)::value;
            bool ok = true;
            typename Ar<FAST>::R ra = Ar<FAST>::recip(a, ok);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#if PYH_FOLD_POW2
                Rk[k] = Ar<FAST>::div(dI[k], ra, ok);                                  // = 2 R; the 0.5 moves into the RK coefficient
#else
                Rk[k] = Ar<FAST>::div(0.5 * dI[k], ra, ok);
#endif
            }
            return ok;
        };
        if (!resid(FastTag{})) resid(SafeTag{});
        // RK partial sums (explicit_runge_kutta.py:66-89)
        double un[4] = {0.0, 0.0, 0.0, 0.0};
        if (nt_ > 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                un[k] = plan.t[0].add ? s0[k] + c0 * Rk[k] : s0[k];
                base[plan.t[0].dst + k * PL + o] = un[k];
            }
        }
        if (nt_ > 1) {   // targets 0 and 1 by static index (no local copy of the plan), the rare rest by a loop
#pragma unroll
            for (int k = 0; k < 4; ++k) base[plan.t[1].dst + k * PL + o] = plan.t[1].add ? s1[k] + c1 * Rk[k] : s1[k];
        }
        if (nt_ > 2) {   // tableaux with more than two live rows (e.g. DormandPrince5); unrolled: a run-time index into the kernel
                         // parameter `plan` would make every thread copy it to local memory first (17 stores at the top of the kernel)
#pragma unroll
            for (int q = 2; q < PYH_MAX_STAGES; ++q) {
                if (q < nt_) {
                    const double cq = PYH_FOLD_POW2 ? rscale * ctl->coef[plan.t[q].coef] : ctl->coef[plan.t[q].coef];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const double src = base[plan.t[q].src + k * PL + o];
                        base[plan.t[q].dst + k * PL + o] = plan.t[q].add ? src + cq * Rk[k] : src;
                    }
                }
            }
        }
        if (plan.fuse_dt && nt_ > 0) {
            // QuadBlock.get_dt (quad_block.py:423-436) + realizability (states/conservative.py:161-165) of the state this step ends with
            double tx_, ty_;
            auto cfl = [&](auto tag) -> bool {
                constexpr bool FAST = decltype(tag)::value;
                bool ok = true;
                typename Ar<FAST>::R rr = Ar<FAST>::recip(un[0], ok);
                const double u = Ar<FAST>::div(un[1], rr, ok), v = Ar<FAST>::div(un[2], rr, ok);
                const double p = C.gm1 * (un[3] - un[0] * (0.5 * (u * u + v * v)));
                const double a_ = Ar<FAST>::sqrt(Ar<FAST>::div(C.g * p, rr, ok), ok);
                tx_ = Ar<FAST>::div(cdx, fabs(u) + a_, ok);
                ty_ = Ar<FAST>::div(cdy, fabs(v) + a_, ok);
                return ok;
            };
            if (!cfl(FastTag{})) cfl(SafeTag{});
            tm = dmin2(tx_, ty_);
            // unrealizable (or NaN): -inf can never be a CFL time, so it doubles as the flag
            if (!(un[0] > 0.0) || !(un[3] > 0.0) || (tm != tm)) tm = __longlong_as_double(0xfff0000000000000ll);
        }
        if (plan.push_ghost && nt_ > 0) {
            // ghost cells mirroring an edge cell this thread has just written (target 0 is the stage's output state): from registers
            // (no re-read of the value just stored); table and wall geometry were prefetched at the top
            const unsigned dst = plan.t[0].dst;
            if (i == 0 || i == ny - 1) push_ghost_values(blks, B, lay, po, dst, i, j, true, un[0], un[1], un[2], un[3]);
            if (j == 0 || j == nx - 1) push_ghost_values(blks, B, lay, po, dst, i, j, false, un[0], un[1], un[2], un[3]);
        }
    }
    if (plan.fuse_dt) {   // block minimum -> one atomicMin per thread block (quad_block.py:436: min over cells; Solver.get_dt: over blocks)
        const int t = threadIdx.x;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) tm = dmin2(tm, __shfl_xor_sync(0xffffffffu, tm, s));
        if ((t & 31) == 0) sDT[t >> 5] = tm;
        __syncthreads();
        if (t == 0) {
            double m = sDT[0];
#pragma unroll
            for (int w = 1; w < kSplitUpdateThreads / 32; ++w) m = dmin2(m, sDT[w]);
            if (m == __longlong_as_double(0xfff0000000000000ll)) { atomicOr(&ctl_out->bad, 1); atomicExch(&ctl_out->allok, 0ull); }
            else atomicMin(&ctl_out->dtmin_bits, dkey(m));
        }
    }
}

}  // namespace pyh

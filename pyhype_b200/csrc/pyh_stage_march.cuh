// pyh_stage_march.cuh -- fused RK-stage kernel of the MUSCL residual (sm_100a, fp64), row-marching.
//
// Reference: fvm/base.py:108-500 (dUdt, flux sweeps), fvm/SecondOrderMUSCL.py (reconstruction),
// limiters/base.py + limiters/limiters.py, gradients/greengauss.py, blocks/quad_block.py:120-218,
// time_marching/explicit_runge_kutta.py:63-89 -- one launch per RK stage for ALL local blocks.
//
// Work decomposition.  grid = (column strips, row strips, blocks) of the strips `tiles` selects.  A CTA of NT threads owns NT-4
// output columns and `tys` rows of one block; thread t <-> column j = strip_origin - 2 + t and
// marches south -> north.  Lane roles: 0 and NT-1 only publish their column's reconstruction
// variables; 1 and NT-2 additionally evaluate gradient + limiter (their face states are the outer
// Riemann states of the first / last output column), NT-2 also the flux of its west face; lanes
// 2..NT-3 produce output cells.  Ghost columns (-1, nx) ride on whatever lane they fall on: they
// publish the ghost value, and the lane at j == nx evaluates the block's east-edge face.
//
// Everything that is carried from row to row lives in shared-memory rings (3 rows of state, 2 rows
// of east-face states / west-face fluxes / north-face states, 1 row of south-face fluxes), so the
// register file only holds one phase's working set and there is ONE block barrier per row:
//     top : publish row r+1 (loaded one iteration ago), issue the loads of row r+2
//     B(r): gradient, limiter, limited face states of cell (r, j)      -> sFE, sQN (+ QW, QS in registers)
//     ---- __syncthreads ----
//     C(r): flux of the west face (needs the east-face state of j-1), flux of the south face
//           (needs the north-face state of row r-1)                    -> sIW, IS
//     D(r-1): residual of cell (r-1, j) (IN = IS(r), IE = sIW[j+1] of row r-1) + RK partial sums
//
// Arithmetic: pyh_math.cuh.  Each phase first runs with the branch-free division/sqrt sequences
// (Ar<true>), accumulating a validity predicate, and is re-evaluated with the plain IEEE operators
// (Ar<false>) in the (never observed) case that an operand fell outside the fast range.
#pragma once
#include <type_traits>
#include "pyh_layout.cuh"
#include "pyh_math.cuh"
#include "pyh_march_tu.cuh"

namespace pyh {

// Build-time switches that remain (the round-1 tuning knobs were measured on the B200 in round 2, profiles/r02a_variants_ab.txt
// and r02b_variants_combined.txt: the winners are now the only code path, the losers are gone):
//   PYH_SKIP_UNIT_ROT (1): skip the rotation by theta == 0 on blocks whose vertical faces are all axis-aligned.
#ifndef PYH_SKIP_UNIT_ROT
#define PYH_SKIP_UNIT_ROT 1
#endif
constexpr int kUnrollB2 = 2;   // two variables' limiter chains in flight (8 division chains): -0.9 % measured; 4 was slower

// the test hooks of the kernel (gradient / limiter / residual stores for pyh_debug_fetch and pyh_residual) sit behind calls to
// out-of-line functions, so that their address arithmetic is not issued (predicated off) for every cell in production
static __device__ __noinline__ void hook_store2(double* p, size_t i0, double v0, size_t i1, double v1) { p[i0] = v0; p[i1] = v1; }
static __device__ __noinline__ void hook_store1(double* p, size_t i0, double v0) { p[i0] = v0; }
static __device__ __noinline__ void hook_store4(double* p, size_t i0, size_t stride, double v0, double v1, double v2, double v3) {
    p[i0] = v0; p[i0 + stride] = v1; p[i0 + 2 * stride] = v2; p[i0 + 3 * stride] = v3;
}

// state and geometry are read-only for the lifetime of a launch (the stage writes other buffers; the ghost frame of its input
// was completed by earlier kernels): ld.global.nc (LDG.E.CONSTANT) instead of the generic LD the compiler emits for pointers
// fetched from the block table
#if !defined(PYH_HOST_TWIN)
#define PYH_RO(expr) __ldg(&(expr))
#else
#define PYH_RO(expr) (expr)
#endif
#define PYH_ROS(expr) PYH_RO(expr)

// Ghost refresh at the source (Blocks.apply_boundary_condition, blocks/base.py:448-471, GhostBlock.* blocks/ghost.py:187-278, as
// a PUSH): a thread block that has written cells on the edge of its mesh block also writes, at its very end, every ghost cell
// that mirrors them: the neighbour block's ghost frame (same GPU), the block's own ghost frame through the boundary-condition
// functor (reflection / slip wall, outlet copy, Dirichlet inlet), or the send buffer of a neighbour on another rank.  Same
// values, same arithmetic (`reflect`) as k_ghost / k_pack_halo, which remain for the refresh after an upload; what disappears
// is one kernel launch per stage (and the pack launch before every NCCL exchange).  Cell (i, j) of the stage's output
// buffer `dst`, sides selected by `ns` (north / south) or east / west.  (Called from the tail of the kernel: a call inside
// the row loop, however rarely taken, cost the loop 12 % through the registers it pinned -- measured.)
static __device__ __forceinline__ void push_ghost_values(const BlkDev* __restrict__ blks, const BlkDev& B, const Layout lay, const PlaneOffsets po,
                                                         const unsigned dst, const int i, const int j, const bool ns,
                                                         const double u0, const double u1, const double u2, const double u3) {
    const int nx = lay.nx, ny = lay.ny;
    const unsigned PL = lay.plane;
#pragma unroll 1
    for (int side = ns ? PYH_NORTH : PYH_EAST; side <= (ns ? PYH_SOUTH : PYH_WEST); ++side) {
        int gi, gj, oi, oj, fi, fj, idx;      // own ghost cell, the neighbour's mirror ghost cell, the boundary face, index along the edge
        if (side == PYH_EAST)       { if (j != nx - 1) continue; idx = i; gi = i; gj = nx; oi = i; oj = -1; fi = i; fj = nx; }
        else if (side == PYH_WEST)  { if (j != 0) continue;      idx = i; gi = i; gj = -1; oi = i; oj = nx; fi = i; fj = 0; }
        else if (side == PYH_NORTH) { if (i != ny - 1) continue; idx = j; gi = ny; gj = j; oi = -1; oj = j; fi = ny; fj = j; }
        else                        { if (i != 0) continue;      idx = j; gi = -1; gj = j; oi = ny; oj = j; fi = 0;  fj = j; }
        const int bc = B.bc[side];
        double q0 = u0, q1 = u1, q2 = u2, q3 = u3;
        if (bc != PYH_BC_NONE || (B.nbr[side] < 0 && B.remote_slot[side] < 0)) {      // the block's own ghost strip (+ BC functor)
            if (bc == PYH_BC_REFLECTION || bc == PYH_BC_SLIPWALL) {
                const unsigned of = lay.at(fi, fj);
                const bool vert = (side == PYH_EAST || side == PYH_WEST);
                const double c_ = B.base[(vert ? po.cv : po.ch) + of], s_ = B.base[(vert ? po.sv : po.sh) + of];
                reflect(q1, q2, c_, s_);
            } else if (bc == PYH_BC_PRIMITIVE_DIRICHLET) {
                const double* d = B.dir_cons[side] + 4 * (long long)idx;
                q0 = d[0]; q1 = d[1]; q2 = d[2]; q3 = d[3];
            }
            double* g = B.base + dst + lay.at(gi, gj);
            g[0] = q0; g[PL] = q1; g[2 * PL] = q2; g[3 * PL] = q3;
        } else if (B.nbr[side] >= 0) {                                                  // neighbour on this GPU: its ghost frame
            double* g = blks[B.nbr[side]].base + dst + lay.at(oi, oj);
            g[0] = q0; g[PL] = q1; g[2 * PL] = q2; g[3 * PL] = q3;
        } else {                                                                        // neighbour on another rank: the send buffer
            double* g = B.send[side] + 4 * (long long)idx;
            g[0] = q0; g[1] = q1; g[2] = q2; g[3] = q3;
        }
    }
}
// ... the same for a cell whose new value is re-read from the output buffer (the fused kernel: its tail runs after the row loop)
static __device__ __forceinline__ void push_ghost_cell(const BlkDev* __restrict__ blks, const BlkDev& B, const Layout lay, const PlaneOffsets po,
                                                       const unsigned dst, const int i, const int j, const bool ns) {
    const unsigned PL = lay.plane;
    const double* u = B.base + dst + lay.at(i, j);
    push_ghost_values(blks, B, lay, po, dst, i, j, ns, u[0], u[PL], u[2 * PL], u[3 * PL]);
}

typedef std::integral_constant<bool, true> FastTag;
typedef std::integral_constant<bool, false> SafeTag;

template <int FLUX, int LIM, int PRIM, int NQ>
__global__ void __launch_bounds__(march_max_threads(NQ), march_min_blocks(NQ))
k_stage_march(const BlkDev* __restrict__ blks, const Layout lay, const PlaneOffsets po, const StagePlan plan,
              const Control* __restrict__ ctl, Control* __restrict__ ctl_out, const Consts C, const int tys, const int want_grad_dbg,
              const MarchTiles tiles) {
    if (!ctl->active) return;
    // Block coordinates: (column strip, row strip, mesh block); `tiles` selects the strips this launch covers (pyh_march_tu.cuh).
    const unsigned bx = (unsigned)(tiles.xfirst + (int)blockIdx.x * tiles.xstride);
    const unsigned by = blockIdx.y;
    const unsigned bz = blockIdx.z;
#ifdef PYH_HOST_TWIN
    extern double smem[];            // tests/host_twin: the emulator's per-block buffer
#else
    extern __shared__ double smem[];
#endif
    const int NT = blockDim.x;
    const int t = threadIdx.x;
    double* const sQ = smem;                       // [3][4][NT]
    double* const sFE = sQ + 12 * NT;              // [2][NQ][4][NT]  east-face states at the quadrature points
    double* const sIW = sFE + 8 * NQ * NT;         // [2][4][NT]      integrated west-face fluxes
    double* const sQN = sIW + 8 * NT;              // [2][NQ][4][NT]  north-face states
    double* const sIS = sQN + 8 * NQ * NT;         // [4][NT]
    double* const sQW = sIS + 4 * NT;              // [NQ][4][NT]  west-face states of (r, j), private
    double* const sQS = sQW + 4 * NQ * NT;         // [NQ][4][NT]  south-face states of (r, j), private
    double dtmin = __longlong_as_double(0x7ff0000000000000ll);   // running CFL minimum of this thread's cells (plan.fuse_dt)

    auto iFE = [&](int par_, int q, int k, int tt) { return ((par_ * NQ + q) * 4 + k) * NT + tt; };   // also sQN
    auto iQW = [&](int q, int k) { return (q * 4 + k) * NT + t; };                                    // also sQS
    const BlkDev& B = blks[bz];
    double* __restrict__ const base = B.base;
    const int bcE = B.bc[PYH_EAST], bcW = B.bc[PYH_WEST], bcN = B.bc[PYH_NORTH], bcS = B.bc[PYH_SOUTH];
    const int cart = B.cart & 1;
    // rotation by theta == 0 (u*1 + v*0, v*1 - u*0) returns its argument by value for every finite state: skip it on
    // blocks whose vertical faces are all axis-aligned (BlkDev::cart bit 1), like the reference does for is_cartesian blocks
    const bool vident = cart || (PYH_SKIP_UNIT_ROT && (B.cart & 2));
    const int nx = lay.nx, ny = lay.ny, pitch = lay.pitch;
    const unsigned PL = lay.plane;
    const int j = (int)bx * (NT - 4) - 2 + t;
    const int i0 = tiles.row0 + (int)by * tiles.rowstride;
    const int i1 = min(i0 + tys, tiles.row1);
    const bool act = (j >= -1) && (j <= nx);
    const bool real = (j >= 0) && (j < nx);
    const bool doB = real && (t >= 1) && (t <= NT - 2);
    const bool doV = (t >= 2) && (t <= NT - 2) && (j >= 0) && (j <= nx);   // west-face flux
    const bool outcol = real && (t >= 2) && (t <= NT - 3);
    const int jc = min(max(j, -1), nx);
    const double* __restrict__ const U = base + plan.cur;
    const double* __restrict__ const G = base;   // geometry planes: read-only for the lifetime of the context

    // ---- helpers ---------------------------------------------------------------------------------
    auto exists = [&](int row) {   // does cell (row, j) exist (interior or ghost frame without corners)?
        return act && (row >= -1) && (row <= ny) && !((row == -1 || row == ny) && (j == -1 || j == nx));
    };
    auto load_raw = [&](int row, double q[4]) {
        if (exists(row)) {
            unsigned o = (unsigned)((row + 1) * pitch + PADL + jc);
            q[0] = PYH_ROS(U[o]); q[1] = PYH_ROS(U[o + PL]); q[2] = PYH_ROS(U[o + 2 * PL]); q[3] = PYH_ROS(U[o + 3 * PL]);
        } else {
            q[0] = 1.0; q[1] = 0.0; q[2] = 0.0; q[3] = 1.0;
        }
    };
    // BaseBlockGhost.from_block (quad_block.py:120-134): conservative -> reconstruction variables
    auto to_recon = [&](double q[4]) {
        if (PRIM) {
            double s[4] = {q[0], q[1], q[2], q[3]};
            bool ok = true;
            cons2prim<true>(q, C, ok);
            if (!ok) { q[0] = s[0]; q[1] = s[1]; q[2] = s[2]; q[3] = s[3]; cons2prim<false>(q, C, ok); }
        }
    };
    auto publish = [&](int slot, const double q[4]) {
#pragma unroll
        for (int k = 0; k < 4; ++k) sQ[(slot * 4 + k) * NT + t] = q[k];
    };
    // GhostBlock.apply_boundary_condition_to_state on an edge state (fvm/base.py:352-362)
    auto apply_bc_edge = [&](int bc, int side, int idx, double c_, double s_, double q[4]) {
        if (bc == PYH_BC_REFLECTION || bc == PYH_BC_SLIPWALL) reflect(q[1], q[2], c_, s_);
        else if (bc == PYH_BC_PRIMITIVE_DIRICHLET) {
            const double* d = B.dir_recon[side] + 4 * (long long)idx;
            q[0] = d[0]; q[1] = d[1]; q[2] = d[2]; q[3] = d[3];
        }
    };

    // ---- prologue: rows r0-1, r0 into the ring, row r0+1 in flight --------------------------------------
    const int r0 = (i0 > 0) ? i0 - 1 : i0;   // first row whose gradient is needed
    int sm = 0, sc = 1, sp = 2;               // ring slots of rows r-1, r, r+1
    double Qn[4];
    load_raw(r0 - 1, Qn); to_recon(Qn); publish(sm, Qn);
    load_raw(r0, Qn);     to_recon(Qn); publish(sc, Qn);
    load_raw(r0 + 1, Qn);
    __syncthreads();

    for (int r = r0; r <= i1; ++r) {
        const int par = r & 1;
        const bool rowreal = r < ny;
        const bool full = (r >= i0) && (r < i1);           // rows this strip outputs
        const unsigned o = (unsigned)((r + 1) * pitch + PADL + jc);

        // top: publish row r+1, issue the loads of row r+2 (L2 prefetch hints were measured and hurt: profiles/r01h_summary.md)
        to_recon(Qn);
        publish(sp, Qn);
        if ((r + 1 <= i1) && (r + 1 < ny)) load_raw(r + 2, Qn);
        else { Qn[0] = 1.0; Qn[1] = 0.0; Qn[2] = 0.0; Qn[3] = 1.0; }

        // ---- B(r) ------------------------------------------------------------------------------------
        {
            const double* qm_ = sQ + sm * 4 * NT;
            const double* qc_ = sQ + sc * 4 * NT;
            const double* qp_ = sQ + sp * 4 * NT;
            if (doB && rowreal) {
                const unsigned oE = o + 1, oN = o + pitch;
                // GreenGauss._get_gradinet_JIT (gradients/greengauss.py:110-155)
                double LE = PYH_RO(G[po.Lv + oE]), LW = PYH_RO(G[po.Lv + o]), LN = PYH_RO(G[po.Lh + oN]), LS = PYH_RO(G[po.Lh + o]);
#if PYH_FOLD_POW2
                // half face weights: (0.5 (q + qE)) * xlE == (q + qE) * (0.5 xlE), both scalings exact (pyh_math.cuh)
                LE = 0.5 * LE; LW = 0.5 * LW; LN = 0.5 * LN; LS = 0.5 * LS;
#endif
                double xlE = LE * PYH_RO(G[po.cv + oE]), xlW = LW * (-PYH_RO(G[po.cv + o]));
                double xlN = LN * PYH_RO(G[po.ch + oN]), xlS = LS * (-PYH_RO(G[po.ch + o]));
                double ylE = LE * PYH_RO(G[po.sv + oE]), ylW = LW * (-PYH_RO(G[po.sv + o]));
                double ylN = LN * PYH_RO(G[po.sh + oN]), ylS = LS * (-PYH_RO(G[po.sh + o]));
                double Acell = PYH_RO(G[po.A + o]);
                double dx[NQ][4], dy[NQ][4];
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
#pragma unroll
                    for (int f = 0; f < 4; ++f) {
                        dx[q][f] = PYH_RO(G[po.dxy + ((q * 4 + f) * 2) * PL + o]);
                        dy[q][f] = PYH_RO(G[po.dxy + ((q * 4 + f) * 2 + 1) * PL + o]);
                    }
                }
                bool okA = true;
                double ia = Ar<true>::rcp(Acell, okA);
                if (!okA) ia = 1.0 / Acell;
                // pass 1: gradient and the high-order terms of the four faces (at every quadrature point); the
                // terms wait in this thread's own face-state slots so that no geometry is live while the limiter divides
#pragma unroll 1
                for (int k = 0; k < 4; ++k) {
                    const double q = qc_[k * NT + t], qW = qc_[k * NT + t - 1], qE = qc_[k * NT + t + 1];
                    const double qS = qm_[k * NT + t], qN = qp_[k * NT + t];
                    // face averages (quad_block.py:181-218)
#if PYH_FOLD_POW2
                    double fE = q + qE, fW = qW + q, fN = q + qN, fS = qS + q;
#else
                    double fE = 0.5 * (q + qE), fW = 0.5 * (qW + q), fN = 0.5 * (q + qN), fS = 0.5 * (qS + q);
#endif
                    double gx = (fE * xlE + fW * xlW + fN * xlN + fS * xlS) * ia;
                    double gy = (fE * ylE + fW * ylW + fN * ylN + fS * ylS) * ia;
#pragma unroll
                    for (int p = 0; p < NQ; ++p) {
                        sFE[iFE(par, p, k, t)] = gx * dx[p][0] + gy * dy[p][0];       // blocks/base.py:283-288
                        sQW[iQW(p, k)] = gx * dx[p][1] + gy * dy[p][1];
                        sQN[iFE(par, p, k, t)] = gx * dx[p][2] + gy * dy[p][2];
                        sQS[iQW(p, k)] = gx * dx[p][3] + gy * dy[p][3];
                    }
                    if (want_grad_dbg) { if (full && outcol) hook_store2(B.dbgG, k * (size_t)PL + o, gx, (4 + k) * (size_t)PL + o, gy); }
                }
                // pass 2: SlopeLimiter._get_slope / _limit (limiters/base.py:47-108, 179-187), four faces side by side;
                // phi is the minimum over every quadrature point of every face
#pragma unroll kUnrollB2
                for (int k = 0; k < 4; ++k) {
                    const double q = qc_[k * NT + t], qW = qc_[k * NT + t - 1], qE = qc_[k * NT + t + 1];
                    const double qS = qm_[k * NT + t], qN = qp_[k * NT + t];
                    double term[NQ][4];
                    // maximum and minimum of the five values with 6 instead of 8 FP64 comparisons: order the two neighbour
                    // pairs once (one comparison yields both the larger and the smaller), then reduce
                    const bool wge = qW > qE, sgn = qS > qN;
                    const double h1 = wge ? qW : qE, l1 = wge ? qE : qW, h2 = sgn ? qS : qN, l2 = sgn ? qN : qS;
                    double mx = dmax2(dmax2(h1, h2), q);
                    double mn = dmin2(dmin2(l1, l2), q);
                    double dmx = mx - q, dmn = mn - q;
                    double phi = 0.0;
#pragma unroll
                    for (int p = 0; p < NQ; ++p) {
                        double davg[4];
                        term[p][0] = sFE[iFE(par, p, k, t)];
                        term[p][1] = sQW[iQW(p, k)];
                        term[p][2] = sQN[iFE(par, p, k, t)];
                        term[p][3] = sQS[iQW(p, k)];
#pragma unroll
                        for (int f = 0; f < 4; ++f) davg[f] = (q + term[p][f]) - q;   // limiters/base.py:99-102
                        double pp;
#if PYH_UNIFORM_SHORTCUT
                        // locally constant reconstruction: slope = 1 on every face (limiters/base.py:213-221); see pyh_math.cuh
                        if (davg[0] == 0.0 && davg[1] == 0.0 && davg[2] == 0.0 && davg[3] == 0.0) pp = limiter_at_one<LIM>();
                        else
#endif
                        if (!limiter4_fast<LIM>(dmx, dmn, davg, pp)) limiter4_safe<LIM>(dmx, dmn, davg, pp);
                        phi = (p == 0) ? pp : dmin2_nan(phi, pp);
                    }
                    if (phi < 0.0) phi = 0.0;                                         // limiters/base.py:187
#pragma unroll
                    for (int p = 0; p < NQ; ++p) {
                        sFE[iFE(par, p, k, t)] = q + phi * term[p][0];                // SecondOrderMUSCL.py:124-126
                        sQW[iQW(p, k)] = q + phi * term[p][1];
                        sQN[iFE(par, p, k, t)] = q + phi * term[p][2];
                        sQS[iQW(p, k)] = q + phi * term[p][3];
                    }
                    if (want_grad_dbg) { if (full && outcol) hook_store1(B.dbgG, (8 + k) * (size_t)PL + o, phi); }
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double q = qc_[k * NT + t];
#pragma unroll
                    for (int p = 0; p < NQ; ++p) {
                        sFE[iFE(par, p, k, t)] = q;
                        sQN[iFE(par, p, k, t)] = q;
                        sQW[iQW(p, k)] = q;
                        sQS[iQW(p, k)] = q;
                    }
                }
            }
        }
        __syncthreads();

        // ---- C(r): west face J = j of row r ----------------------------------------------------------
        double IW[4] = {0.0, 0.0, 0.0, 0.0};
        if (full && doV) {
            const double cf = PYH_RO(G[po.cv + o]), sf = PYH_RO(G[po.sv + o]), Lf = PYH_RO(G[po.Lv + o]);
            // riemann_flux returns flux_scale(FLUX) * F (pyh_math.cuh); the face length absorbs the factor, exactly
            const double Lf1 = (flux_scale(FLUX) == 2.0) ? Lf : 2.0 * Lf;     // one point:  L * (0 + 2 F)
            const double Lfq = (flux_scale(FLUX) == 2.0) ? 0.5 * Lf : Lf;     // 2, 3 points: L * sum_p w_p F_p
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
            for (int p = 0; p < NQ; ++p) {   // one Riemann problem per quadrature point (fvm/base.py:344-351)
                double QL0[4], QR0[4];
                if (j > 0) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) QL0[k] = sFE[iFE(par, p, k, t - 1)];
                } else if (bcW == PYH_BC_NONE) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) QL0[k] = sQ[(sc * 4 + k) * NT + t - 1];      // ghost cell (r, -1), fvm/base.py:305-325
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) QL0[k] = sQW[iQW(p, k)];
                    apply_bc_edge(bcW, PYH_WEST, r, cf, sf, QL0);
                }
                if (j < nx) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) QR0[k] = sQW[iQW(p, k)];
                } else if (bcE == PYH_BC_NONE) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) QR0[k] = sQ[(sc * 4 + k) * NT + t];          // ghost cell (r, nx)
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) QR0[k] = sFE[iFE(par, p, k, t - 1)];         // east-face state of cell (r, nx-1)
                    apply_bc_edge(bcE, PYH_EAST, r, cf, sf, QR0);
                }
                if (!vident) { rot(QL0[1], QL0[2], cf, sf); rot(QR0[1], QR0[2], cf, sf); }    // fvm/base.py:366-376
                double Fq[4];
                auto faceV = [&](auto tag) -> bool {
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
                    if (!vident) unrot(Fq[1], Fq[2], cf, sf);                                // fvm/base.py:388-390
                    return ok;
                };
                if (!faceV(FastTag{})) faceV(SafeTag{});
                // integrate_flux (fvm/base.py:188-190): L * (((0 + w0 F0) + w1 F1) + w2 F2); one point: w0 = 2
                if (NQ == 1) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) IW[k] = PYH_FOLD_POW2 ? Lf1 * Fq[k] : Lf * (2.0 * Fq[k]);
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc[k] = acc[k] + C.qw[p] * Fq[k];
                }
            }
            if (NQ > 1) {
#pragma unroll
                for (int k = 0; k < 4; ++k) IW[k] = Lfq * acc[k];
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) sIW[(par * 4 + k) * NT + t] = IW[k];

        // ---- C(r): south face I = r of column j + D(r-1) ------------------------------------------------
        if (outcol && (r >= i0)) {
            double dA = 1.0, dS0[4] = {0.0, 0.0, 0.0, 0.0}, dS1[4] = {0.0, 0.0, 0.0, 0.0};
            auto load_d = [&]() {
                const unsigned om_ = o - pitch;
                dA = PYH_RO(G[po.A + om_]);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    dS0[k] = (plan.ntargets > 0) ? base[plan.t[0].src + k * PL + om_] : 0.0;
                    dS1[k] = (plan.ntargets > 1) ? base[plan.t[1].src + k * PL + om_] : 0.0;
                }
            };
            const double cf = PYH_RO(G[po.ch + o]), sf = PYH_RO(G[po.sh + o]), Lf = PYH_RO(G[po.Lh + o]);
            const double Lf1 = (flux_scale(FLUX) == 2.0) ? Lf : 2.0 * Lf;
            const double Lfq = (flux_scale(FLUX) == 2.0) ? 0.5 * Lf : Lf;
            double IS[4];
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
            for (int p = 0; p < NQ; ++p) {
                double QL0[4], QR0[4];
                if (r > 0) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) QL0[k] = sQN[iFE(par ^ 1, p, k, t)];
                } else if (bcS == PYH_BC_NONE) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) QL0[k] = sQ[(sm * 4 + k) * NT + t];          // ghost cell (-1, j)
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) QL0[k] = sQS[iQW(p, k)];
                    apply_bc_edge(bcS, PYH_SOUTH, j, cf, sf, QL0);
                }
                if (r < ny) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) QR0[k] = sQS[iQW(p, k)];
                } else if (bcN == PYH_BC_NONE) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) QR0[k] = sQ[(sc * 4 + k) * NT + t];          // ghost cell (ny, j)
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) QR0[k] = sQN[iFE(par ^ 1, p, k, t)];
                    apply_bc_edge(bcN, PYH_NORTH, j, cf, sf, QR0);
                }
                if (cart) { rot90(QL0[1], QL0[2]); rot90(QR0[1], QR0[2]); }                  // fvm/base.py:435-441
                else { rot(QL0[1], QL0[2], cf, sf); rot(QR0[1], QR0[2], cf, sf); }
                double Fq[4];
                auto faceH = [&](auto tag) -> bool {
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
                    if (cart) unrot90(Fq[1], Fq[2]); else unrot(Fq[1], Fq[2], cf, sf);      // fvm/base.py:482-486
                    return ok;
                };
                if (!faceH(FastTag{})) faceH(SafeTag{});
                if (NQ == 1) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) IS[k] = PYH_FOLD_POW2 ? Lf1 * Fq[k] : Lf * (2.0 * Fq[k]);
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc[k] = acc[k] + C.qw[p] * Fq[k];
                }
            }
            if (NQ > 1) {
#pragma unroll
                for (int k = 0; k < 4; ++k) IS[k] = Lfq * acc[k];
            }

            // D(r-1): residual (fvm/base.py:141-165) + RK partial sums (explicit_runge_kutta.py:66-89)
            if (r - 1 >= i0) {
                const unsigned om = o - pitch;
                load_d();
                const double a = dA;
                double Rk[4];
                auto resid = [&](auto tag) -> bool {
                    constexpr bool FAST = decltype(tag)::value;
                    bool ok = true;
                    typename Ar<FAST>::R ra = Ar<FAST>::recip(a, ok);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        double IWp = sIW[((par ^ 1) * 4 + k) * NT + t], IEp = sIW[((par ^ 1) * 4 + k) * NT + t + 1];
                        double ISp = sIS[k * NT + t];
#if PYH_FOLD_POW2
                        Rk[k] = Ar<FAST>::div(IWp - IEp + ISp - IS[k], ra, ok);        // = 2 R; the 0.5 moves into the RK coefficient
#else
                        Rk[k] = Ar<FAST>::div(0.5 * (IWp - IEp + ISp - IS[k]), ra, ok);
#endif
                    }
                    return ok;
                };
                if (!resid(FastTag{})) resid(SafeTag{});
                constexpr double rscale = PYH_FOLD_POW2 ? 0.5 : 1.0;   // Rk == R / rscale
                if (plan.write_residual) hook_store4(B.dbg, om, (size_t)PL, rscale * Rk[0], rscale * Rk[1], rscale * Rk[2], rscale * Rk[3]);
                {
                    // all source loads first, then the updates (targets 0 and 1 by static index: no local copies)
                    const int nt_ = plan.ntargets;
                    double s0[4], s1[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        s0[k] = dS0[k];
                        s1[k] = dS1[k];
                    }
                    if (nt_ > 0) {
                        const double c0 = PYH_FOLD_POW2 ? rscale * ctl->coef[plan.t[0].coef] : ctl->coef[plan.t[0].coef];
                        double un[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) { un[k] = plan.t[0].add ? s0[k] + c0 * Rk[k] : s0[k]; base[plan.t[0].dst + k * PL + om] = un[k]; }
                        if (plan.fuse_dt) {
                            // QuadBlock.get_dt (quad_block.py:423-436) + the realizability conditions (states/conservative.py:161-165) on
                            // the state this step ends with, which is still in registers: the next step's Solver.get_dt costs no pass
                            // over the state (k_dt does the same arithmetic as a kernel of its own for the first step of a run)
                            const double cdx = PYH_RO(G[po.cdx + om]), cdy = PYH_RO(G[po.cdy + om]);
                            double tx, ty;
                            auto cfl = [&](auto tag) -> bool {
                                constexpr bool FAST = decltype(tag)::value;
                                bool ok = true;
                                typename Ar<FAST>::R rr = Ar<FAST>::recip(un[0], ok);
                                const double u = Ar<FAST>::div(un[1], rr, ok), v = Ar<FAST>::div(un[2], rr, ok);
                                const double p = C.gm1 * (un[3] - un[0] * (0.5 * (u * u + v * v)));
                                const double a_ = Ar<FAST>::sqrt(Ar<FAST>::div(C.g * p, rr, ok), ok);
                                tx = Ar<FAST>::div(cdx, fabs(u) + a_, ok);
                                ty = Ar<FAST>::div(cdy, fabs(v) + a_, ok);
                                return ok;
                            };
                            if (!cfl(FastTag{})) cfl(SafeTag{});
                            double tm = dmin2(tx, ty);
                            // unrealizable (or NaN): -inf can never be a CFL time, so it doubles as the flag
                            if (!(un[0] > 0.0) || !(un[3] > 0.0) || (tm != tm)) tm = __longlong_as_double(0xfff0000000000000ll);
                            dtmin = dmin2(dtmin, tm);
                        }
                    }
                    if (nt_ > 1) {
                        const double c1 = PYH_FOLD_POW2 ? rscale * ctl->coef[plan.t[1].coef] : ctl->coef[plan.t[1].coef];
#pragma unroll
                        for (int k = 0; k < 4; ++k) base[plan.t[1].dst + k * PL + om] = plan.t[1].add ? s1[k] + c1 * Rk[k] : s1[k];
                    }
                    for (int q = 2; q < nt_; ++q) {   // tableaux with more than two live rows (e.g. DormandPrince5)
                        const double cq = PYH_FOLD_POW2 ? rscale * ctl->coef[plan.t[q].coef] : ctl->coef[plan.t[q].coef];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            double src = base[plan.t[q].src + k * PL + om];
                            base[plan.t[q].dst + k * PL + om] = plan.t[q].add ? src + cq * Rk[k] : src;
                        }
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) sIS[k * NT + t] = IS[k];
        }

        // rotate the ring
        const int tmp = sm; sm = sc; sc = sp; sp = tmp;
    }
    if (plan.fuse_dt) {   // block minimum -> one atomicMin per thread block (quad_block.py:436: min over cells; Solver.get_dt: over blocks)
        // the rings are dead now: one of their rows serves as the reduction buffer
        __syncthreads();
        double* const sDT = smem;
        sDT[t] = dtmin;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if (t < s && t + s < NT) sDT[t] = dmin2(sDT[t], sDT[t + s]);
            __syncthreads();
        }
        if (t == 0) {
            const double m = sDT[0];
            if (m == __longlong_as_double(0xfff0000000000000ll)) { atomicOr(&ctl_out->bad, 1); atomicExch(&ctl_out->allok, 0ull); }
            else atomicMin(&ctl_out->dtmin_bits, dkey(m));
        }
    }
    if (plan.push_ghost && plan.ntargets > 0) {
        // ghost cells mirroring the block-edge cells this thread block has written (target 0 is the stage's output state)
        const int jlo = (int)bx * (NT - 4), jhi = min(jlo + NT - 4, nx);      // its output columns
        const bool rowS = (i0 == 0) && (i1 > 0), rowN = (i1 == ny) && (i1 > i0);
        const bool colW = (jlo == 0), colE = (jhi == nx);
        if ((rowS || rowN || colW || colE) && (i1 > i0)) {
            __syncthreads();      // the stores of the whole thread block are visible to each of its threads
            const unsigned dst = plan.t[0].dst;
            if (outcol) {
                if (rowS) push_ghost_cell(blks, B, lay, po, dst, 0, j, true);
                if (rowN && !(rowS && ny == 1)) push_ghost_cell(blks, B, lay, po, dst, ny - 1, j, true);
            }
            for (int i = i0 + t; i < i1; i += NT) {
                if (colW) push_ghost_cell(blks, B, lay, po, dst, i, 0, false);
                if (colE && !(colW && nx == 1)) push_ghost_cell(blks, B, lay, po, dst, i, nx - 1, false);
            }
        }
    }
}

}  // namespace pyh

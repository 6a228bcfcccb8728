// pyh_kernels.cuh -- CUDA kernels of the MUSCL residual + RK stage update (sm_100a, fp64).
//
// k_stage_tile : one fused kernel per RK stage for ALL local blocks: stage the halo-padded tile
//                in shared memory -> Green-Gauss gradient + limiter -> limited face states ->
//                rotate -> Riemann flux -> unrotate -> flux integration -> residual -> RK
//                partial-sum updates.  (reference: fvm/base.py:108-500, SecondOrderMUSCL.py,
//                limiters/base.py, gradients/greengauss.py, time_marching/explicit_runge_kutta.py)
// k_ghost      : ghost-strip refresh + boundary conditions (blocks/base.py:448-471, ghost.py)
// k_dt, k_dt_finalize, k_step_end : CFL reduction and the device-resident time loop control
//                (quad_block.py:423-436, solvers/base.py:114-136, Euler2D.py:195-210)
#pragma once
#include "pyh_layout.cuh"
#include "pyh_math.cuh"

namespace pyh {

struct Tableau {
    double a[PYH_MAX_STAGES * PYH_MAX_STAGES];
    int nstages;
};

// ------------------------------------------------------------------------------------------------
// stage kernel, shared-memory tile version
// ------------------------------------------------------------------------------------------------
template <int TX, int TY>
struct TileShape {
    static constexpr int QW = TX + 4, QH = TY + 4, QN = QW * QH;   // state incl. 2-cell halo
    static constexpr int GW = TX + 2, GH = TY + 2, GN = GW * GH;   // gradient/phi incl. 1-cell ring
    static constexpr int NV = TY * (TX + 1);                       // vertical faces
    static constexpr int NH = (TY + 1) * TX;                       // horizontal faces
    static constexpr int SMEM_DOUBLES = 4 * QN + 12 * GN + 4 * NV + 4 * NH;
};

template <int FLUX, int LIM, int PRIM, int TX, int TY, int NT>
__global__ void __launch_bounds__(NT)
k_stage_tile(const BlkDev* __restrict__ blks, Layout lay, StagePlan plan, const Control* __restrict__ ctl,
             Consts C, int want_grad_dbg) {
    if (!ctl->active) return;
    using T = TileShape<TX, TY>;
    extern __shared__ double smem[];
    double* sQ = smem;
    double* sG = sQ + 4 * T::QN;
    double* sIv = sG + 12 * T::GN;
    double* sIh = sIv + 4 * T::NV;

    const BlkDev& B = blks[blockIdx.z];
    const int nx = lay.nx, ny = lay.ny;
    const int i0 = blockIdx.y * TY, j0 = blockIdx.x * TX;
    const int tid = threadIdx.x;
    const long long PL = lay.plane;
    const double* __restrict__ U = B.H[plan.cur];

    // ---- phase A: stage reconstruction variables of the tile + 2-cell halo -----------------------
    for (int c = tid; c < T::QN; c += NT) {
        int li = c / T::QW - 2, lj = c % T::QW - 2;
        int i = i0 + li, j = j0 + lj;
        bool in = (i >= -1) && (i <= ny) && (j >= -1) && (j <= nx) &&
                  !((i == -1 || i == ny) && (j == -1 || j == nx));
        double q[4] = {1.0, 0.0, 0.0, 1.0};
        if (in) {
            long long o = lay.at(i, j);
            q[0] = U[o]; q[1] = U[o + PL]; q[2] = U[o + 2 * PL]; q[3] = U[o + 3 * PL];
            if (PRIM) cons2prim(q, C);   // BaseBlockGhost.from_block (quad_block.py:120-134)
        }
        sQ[c] = q[0]; sQ[T::QN + c] = q[1]; sQ[2 * T::QN + c] = q[2]; sQ[3 * T::QN + c] = q[3];
    }
    __syncthreads();

    // ---- phase B: Green-Gauss gradient + limiter for the tile and its 1-cell ring ----------------
    for (int c = tid; c < T::GN; c += NT) {
        int li = c / T::GW - 1, lj = c % T::GW - 1;
        int i = i0 + li, j = j0 + lj;
        if (i < 0 || i >= ny || j < 0 || j >= nx) continue;
        long long o = lay.at(i, j);
        long long oE = lay.at(i, j + 1), oN = lay.at(i + 1, j);
        // GreenGauss._get_gradinet_JIT (gradients/greengauss.py:110-155)
        double LE = B.Lv[oE], LW = B.Lv[o], LN = B.Lh[oN], LS = B.Lh[o];
        double xlE = LE * B.cv[oE], xlW = LW * (-B.cv[o]), xlN = LN * B.ch[oN], xlS = LS * (-B.ch[o]);
        double ylE = LE * B.sv[oE], ylW = LW * (-B.sv[o]), ylN = LN * B.sh[oN], ylS = LS * (-B.sh[o]);
        double ia = 1.0 / B.A[o];
        double dx[4], dy[4];
#pragma unroll
        for (int f = 0; f < 4; ++f) { dx[f] = B.dxy[(2 * f) * PL + o]; dy[f] = B.dxy[(2 * f + 1) * PL + o]; }
        int cq = (li + 2) * T::QW + (lj + 2);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double* q_ = sQ + k * T::QN;
            double q = q_[cq], qE = q_[cq + 1], qW = q_[cq - 1], qN = q_[cq + T::QW], qS = q_[cq - T::QW];
            // face averages (quad_block.py:181-218)
            double fE = 0.5 * (q + qE), fW = 0.5 * (qW + q), fN = 0.5 * (q + qN), fS = 0.5 * (qS + q);
            double gx = (fE * xlE + fW * xlW + fN * xlN + fS * xlS) * ia;
            double gy = (fE * ylE + fW * ylW + fN * ylN + fS * ylS) * ia;
            // SlopeLimiter._get_slope (limiters/base.py:47-108)
            double mx = dmax2(dmax2(dmax2(dmax2(q, qW), qE), qS), qN);
            double mn = dmin2(dmin2(dmin2(dmin2(q, qW), qE), qS), qN);
            double dmx = mx - q, dmn = mn - q;
            double phi = 0.0;
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                double term = gx * dx[f] + gy * dy[f];        // blocks/base.py:283-288
                double davg = (q + term) - q;                 // limiters/base.py:99-102
                double pf = limiter_fn<LIM>(slope_of(dmx, dmn, davg));
                phi = (f == 0) ? pf : dmin2(phi, pf);          // limiters/base.py:179-186
            }
            if (phi < 0.0) phi = 0.0;                         // limiters/base.py:187
            sG[k * T::GN + c] = gx;
            sG[(4 + k) * T::GN + c] = gy;
            sG[(8 + k) * T::GN + c] = phi;
            if (want_grad_dbg && li >= 0 && li < TY && lj >= 0 && lj < TX) {
                B.dbgG[k * PL + o] = gx; B.dbgG[(4 + k) * PL + o] = gy; B.dbgG[(8 + k) * PL + o] = phi;
            }
        }
    }
    __syncthreads();

    // limited face state of local cell (li, lj) on side f (SecondOrderMUSCL.py:106-126)
    auto face_state = [&](int li, int lj, int f, double out[4]) {
        long long o = lay.at(i0 + li, j0 + lj);
        double dxf = B.dxy[(2 * f) * PL + o], dyf = B.dxy[(2 * f + 1) * PL + o];
        int cq = (li + 2) * T::QW + (lj + 2), cg = (li + 1) * T::GW + (lj + 1);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double term = sG[k * T::GN + cg] * dxf + sG[(4 + k) * T::GN + cg] * dyf;
            out[k] = sQ[k * T::QN + cq] + sG[(8 + k) * T::GN + cg] * term;
        }
    };
    auto ghost_state = [&](int li, int lj, double out[4]) {
        int cq = (li + 2) * T::QW + (lj + 2);
#pragma unroll
        for (int k = 0; k < 4; ++k) out[k] = sQ[k * T::QN + cq];
    };
    // GhostBlock.apply_boundary_condition_to_state on an edge state (fvm/base.py:352-362)
    auto apply_bc_edge = [&](int side, int idx, double c_, double s_, double q[4]) {
        int bc = B.bc[side];
        if (bc == PYH_BC_REFLECTION || bc == PYH_BC_SLIPWALL) reflect(q[1], q[2], c_, s_);
        else if (bc == PYH_BC_PRIMITIVE_DIRICHLET) {
            const double* d = B.dir_recon[side] + 4 * (long long)idx;
            q[0] = d[0]; q[1] = d[1]; q[2] = d[2]; q[3] = d[3];
        }
    };

    // ---- phase C1: vertical (E/W) faces -----------------------------------------------------------
    for (int c = tid; c < T::NV; c += NT) {
        int li = c / (TX + 1), lJ = c % (TX + 1);
        int i = i0 + li, J = j0 + lJ;
        if (i >= ny || J > nx) continue;
        long long of = lay.at(i, J);
        double cf = B.cv[of], sf = B.sv[of], Lf = B.Lv[of];
        double QL[4], QR[4];
        if (J > 0) face_state(li, lJ - 1, PYH_EAST, QL);
        else if (B.bc[PYH_WEST] == PYH_BC_NONE) ghost_state(li, -1, QL);   // fvm/base.py:305-325
        else { face_state(li, 0, PYH_WEST, QL); apply_bc_edge(PYH_WEST, i, cf, sf, QL); }
        if (J < nx) face_state(li, lJ, PYH_WEST, QR);
        else if (B.bc[PYH_EAST] == PYH_BC_NONE) ghost_state(li, lJ, QR);
        else { face_state(li, lJ - 1, PYH_EAST, QR); apply_bc_edge(PYH_EAST, i, cf, sf, QR); }
        if (!B.cart) { rot(QL[1], QL[2], cf, sf); rot(QR[1], QR[2], cf, sf); }   // fvm/base.py:366-376
        if (!PRIM) { cons2prim(QL, C); cons2prim(QR, C); }                        // fvm/base.py:283-303
        double F[4];
        riemann_flux<FLUX>(QL, QR, F, C);
        if (!B.cart) unrot(F[1], F[2], cf, sf);                                   // fvm/base.py:388-390
#pragma unroll
        for (int k = 0; k < 4; ++k) sIv[k * T::NV + c] = Lf * (2.0 * F[k]);       // integrate_flux, fvm/base.py:188-190
    }
    // ---- phase C2: horizontal (N/S) faces --------------------------------------------------------
    for (int c = tid; c < T::NH; c += NT) {
        int lI = c / TX, lj = c % TX;
        int I = i0 + lI, j = j0 + lj;
        if (I > ny || j >= nx) continue;
        long long of = lay.at(I, j);
        double cf = B.ch[of], sf = B.sh[of], Lf = B.Lh[of];
        double QL[4], QR[4];
        if (I > 0) face_state(lI - 1, lj, PYH_NORTH, QL);
        else if (B.bc[PYH_SOUTH] == PYH_BC_NONE) ghost_state(-1, lj, QL);
        else { face_state(0, lj, PYH_SOUTH, QL); apply_bc_edge(PYH_SOUTH, j, cf, sf, QL); }
        if (I < ny) face_state(lI, lj, PYH_SOUTH, QR);
        else if (B.bc[PYH_NORTH] == PYH_BC_NONE) ghost_state(lI, lj, QR);
        else { face_state(lI - 1, lj, PYH_NORTH, QR); apply_bc_edge(PYH_NORTH, j, cf, sf, QR); }
        if (B.cart) { rot90(QL[1], QL[2]); rot90(QR[1], QR[2]); }                 // fvm/base.py:435-441
        else { rot(QL[1], QL[2], cf, sf); rot(QR[1], QR[2], cf, sf); }
        if (!PRIM) { cons2prim(QL, C); cons2prim(QR, C); }
        double F[4];
        riemann_flux<FLUX>(QL, QR, F, C);
        if (B.cart) unrot90(F[1], F[2]); else unrot(F[1], F[2], cf, sf);          // fvm/base.py:482-486
#pragma unroll
        for (int k = 0; k < 4; ++k) sIh[k * T::NH + c] = Lf * (2.0 * F[k]);
    }
    __syncthreads();

    // ---- phase D: residual (fvm/base.py:141-165) + RK partial sums (explicit_runge_kutta.py:66-89)
    for (int c = tid; c < TX * TY; c += NT) {
        int li = c / TX, lj = c % TX;
        int i = i0 + li, j = j0 + lj;
        if (i >= ny || j >= nx) continue;
        long long o = lay.at(i, j);
        double a = B.A[o];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double IW = sIv[k * T::NV + li * (TX + 1) + lj], IE = sIv[k * T::NV + li * (TX + 1) + lj + 1];
            double IS = sIh[k * T::NH + li * TX + lj], IN = sIh[k * T::NH + (li + 1) * TX + lj];
            double R = 0.5 * (IW - IE + IS - IN) / a;
            for (int t = 0; t < plan.ntargets; ++t) {
                const RkTarget& tg = plan.t[t];
                if (tg.dst == 2) { B.dbg[k * PL + o] = R; continue; }
                double src = (tg.src == 0) ? B.H[plan.u0][k * PL + o] : B.P[tg.row][k * PL + o];
                double out = tg.add ? src + ctl->coef[tg.coef] * R : src;
                if (tg.dst == 0) B.H[plan.next][k * PL + o] = out; else B.P[tg.row][k * PL + o] = out;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// ghost strips + boundary conditions (blocks/ghost.py:187-278)
// grid: (ceil(max(nx,ny)/128), 4 sides, nblocks)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_ghost(const BlkDev* __restrict__ blks, Layout lay, int buf, const Control* __restrict__ ctl) {
    if (!ctl->active) return;
    const BlkDev& B = blks[blockIdx.z];
    const int side = blockIdx.y;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int nx = lay.nx, ny = lay.ny;
    const int len = (side == PYH_EAST || side == PYH_WEST) ? ny : nx;
    if (idx >= len) return;
    const int bc = B.bc[side];
    const bool own = (bc != PYH_BC_NONE) || (B.nbr[side] < 0 && B.remote_slot[side] < 0);
    if (!own && B.nbr[side] < 0) return;  // remote neighbour: filled by k_unpack_halo
    int gi, gj, si, sj, fi, fj;            // ghost cell, source cell, edge face
    if (side == PYH_EAST)       { gi = idx; gj = nx; si = idx; sj = own ? nx - 1 : 0;  fi = idx; fj = nx; }
    else if (side == PYH_WEST)  { gi = idx; gj = -1; si = idx; sj = own ? 0 : nx - 1;  fi = idx; fj = 0; }
    else if (side == PYH_NORTH) { gi = ny; gj = idx; si = own ? ny - 1 : 0; sj = idx;  fi = ny;  fj = idx; }
    else                        { gi = -1; gj = idx; si = own ? 0 : ny - 1; sj = idx;  fi = 0;   fj = idx; }
    const double* src = own ? B.H[buf] : blks[B.nbr[side]].H[buf];
    const long long PL = lay.plane;
    long long os = lay.at(si, sj), og = lay.at(gi, gj);
    double q[4] = {src[os], src[os + PL], src[os + 2 * PL], src[os + 3 * PL]};
    if (bc == PYH_BC_REFLECTION || bc == PYH_BC_SLIPWALL) {
        long long of = lay.at(fi, fj);
        bool vert = (side == PYH_EAST || side == PYH_WEST);
        double c_ = vert ? B.cv[of] : B.ch[of], s_ = vert ? B.sv[of] : B.sh[of];
        reflect(q[1], q[2], c_, s_);
    } else if (bc == PYH_BC_PRIMITIVE_DIRICHLET) {
        const double* d = B.dir_cons[side] + 4 * (long long)idx;
        q[0] = d[0]; q[1] = d[1]; q[2] = d[2]; q[3] = d[3];
    }
    double* dst = B.H[buf];
    dst[og] = q[0]; dst[og + PL] = q[1]; dst[og + 2 * PL] = q[2]; dst[og + 3 * PL] = q[3];
}

// remote halo slots: pack own edge strips / unpack received strips (ghost.py:169-241)
struct HaloSlot { int blk; int side; long long offset; };

__global__ void __launch_bounds__(128)
k_pack_halo(const BlkDev* __restrict__ blks, Layout lay, int buf, const HaloSlot* __restrict__ slots, double* __restrict__ out) {
    const HaloSlot s = slots[blockIdx.y];
    const BlkDev& B = blks[s.blk];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int nx = lay.nx, ny = lay.ny;
    const int len = (s.side == PYH_EAST || s.side == PYH_WEST) ? ny : nx;
    if (idx >= len) return;
    int si, sj;
    if (s.side == PYH_EAST) { si = idx; sj = nx - 1; }
    else if (s.side == PYH_WEST) { si = idx; sj = 0; }
    else if (s.side == PYH_NORTH) { si = ny - 1; sj = idx; }
    else { si = 0; sj = idx; }
    long long o = lay.at(si, sj);
    const double* src = B.H[buf];
    double* d = out + s.offset + 4 * (long long)idx;
    d[0] = src[o]; d[1] = src[o + lay.plane]; d[2] = src[o + 2 * lay.plane]; d[3] = src[o + 3 * lay.plane];
}

__global__ void __launch_bounds__(128)
k_unpack_halo(const BlkDev* __restrict__ blks, Layout lay, int buf, const HaloSlot* __restrict__ slots, const double* __restrict__ in) {
    const HaloSlot s = slots[blockIdx.y];
    const BlkDev& B = blks[s.blk];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int nx = lay.nx, ny = lay.ny;
    const int len = (s.side == PYH_EAST || s.side == PYH_WEST) ? ny : nx;
    if (idx >= len) return;
    int gi, gj;
    if (s.side == PYH_EAST) { gi = idx; gj = nx; }
    else if (s.side == PYH_WEST) { gi = idx; gj = -1; }
    else if (s.side == PYH_NORTH) { gi = ny; gj = idx; }
    else { gi = -1; gj = idx; }
    long long o = lay.at(gi, gj);
    double* dst = B.H[buf];
    const double* d = in + s.offset + 4 * (long long)idx;
    dst[o] = d[0]; dst[o + lay.plane] = d[1]; dst[o + 2 * lay.plane] = d[2]; dst[o + 3 * lay.plane] = d[3];
}

// ------------------------------------------------------------------------------------------------
// CFL reduction + realizability (quad_block.py:423-436, states/conservative.py:161-165)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long dkey(double x) {
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dunkey(unsigned long long k) {
    unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}
constexpr unsigned long long DKEY_INF = 0xfff0000000000000ull;  // dkey(+inf)

__global__ void __launch_bounds__(256)
k_dt(const BlkDev* __restrict__ blks, Layout lay, int buf, int nblk, Control* __restrict__ ctl, Consts C, int respect_active) {
    if (respect_active && !ctl->active) return;
    const int nx = lay.nx, ny = lay.ny;
    const long long ncell = (long long)nx * ny;
    const long long total = ncell * nblk;
    const long long PL = lay.plane;
    double m = __longlong_as_double(0x7ff0000000000000ll);
    int bad = 0;
    for (long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x; n < total; n += (long long)gridDim.x * blockDim.x) {
        int b = (int)(n / ncell);
        long long r = n - (long long)b * ncell;
        int i = (int)(r / nx), j = (int)(r - (long long)i * nx);
        const BlkDev& B = blks[b];
        long long o = lay.at(i, j);
        const double* U = B.H[buf];
        double rho = U[o], ru = U[o + PL], rv = U[o + 2 * PL], e = U[o + 3 * PL];
        if (!(rho > 0.0) || !(e > 0.0)) bad = 1;
        double u = ru / rho, v = rv / rho;
        double p = C.gm1 * (e - rho * (0.5 * (u * u + v * v)));
        double a = sqrt(C.g * p / rho);
        double tx = B.cdx[o] / (fabs(u) + a);
        double ty = B.cdy[o] / (fabs(v) + a);
        double tm = dmin2(tx, ty);
        if (tm != tm) bad = 1;
        m = dmin2(m, tm);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        double o = __shfl_xor_sync(0xffffffffu, m, s);
        m = dmin2(m, o);
        bad |= __shfl_xor_sync(0xffffffffu, bad, s);
    }
    __shared__ double sm[8];
    __shared__ int sb[8];
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sm[w] = m; sb[w] = bad; }
    __syncthreads();
    if (w == 0) {
        m = (l < (blockDim.x >> 5)) ? sm[l] : __longlong_as_double(0x7ff0000000000000ll);
        bad = (l < (blockDim.x >> 5)) ? sb[l] : 0;
#pragma unroll
        for (int s = 4; s > 0; s >>= 1) {
            double o = __shfl_xor_sync(0xffffffffu, m, s);
            m = dmin2(m, o);
            bad |= __shfl_xor_sync(0xffffffffu, bad, s);
        }
        if (l == 0) {
            atomicMin(&ctl->dtmin_bits, dkey(m));
            if (bad) atomicOr(&ctl->bad, 1);
        }
    }
}

// mode 0: device-resident loop (Solver.get_dt clamp + while t < t_final)
// mode 1: write CFL*min to *out only (pyh_local_dt / pyh_get_dt)
__global__ void k_dt_finalize(Control* ctl, double cfl, Tableau tab, int mode, double* out) {
    double dt = cfl * dunkey(ctl->dtmin_bits);      // quad_block.py:436 ; min over blocks is exact
    ctl->dtmin_bits = DKEY_INF;
    if (mode == 1) { *out = dt; return; }
    int active = (ctl->t < ctl->t_final) && !ctl->bad;
    ctl->active = active;
    if (!active) return;
    double rem = ctl->t_final - ctl->t;             // solvers/base.py:132-136
    dt = rem < dt ? rem : dt;
    ctl->dt = dt;
    for (int s = 0; s < tab.nstages; ++s)
        for (int k = 0; k <= s; ++k)
            ctl->coef[s * PYH_MAX_STAGES + k] = dt * tab.a[s * PYH_MAX_STAGES + k];   // explicit_runge_kutta.py:71
}

__global__ void k_set_dt(Control* ctl, double dt_host, const double* dt_dev, Tableau tab) {
    double dt = dt_dev ? *dt_dev : dt_host;
    ctl->dt = dt;
    ctl->active = 1;
    for (int s = 0; s < tab.nstages; ++s)
        for (int k = 0; k <= s; ++k)
            ctl->coef[s * PYH_MAX_STAGES + k] = dt * tab.a[s * PYH_MAX_STAGES + k];
}

__global__ void k_step_end(Control* ctl, double* dts, long long dts_cap) {
    if (!ctl->active) return;
    if (dts && ctl->nsteps < dts_cap) dts[ctl->nsteps] = ctl->dt;
    ctl->t += ctl->dt;                               // Euler2D.py:209-210
    ctl->nsteps += 1;
}

// ------------------------------------------------------------------------------------------------
// setup: derived geometry from node coordinates (exact IEEE restatement of mesh/base.py:57-64,
// mesh/quad_mesh.py:133-135,172-184, mesh/quadratures.py:56-95 for the 1-point rule)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_geometry(Layout lay, const double* __restrict__ xn, const double* __restrict__ yn,  // (ny+1, nx+1) dense
           double* __restrict__ dxy, double* __restrict__ Lv, double* __restrict__ Lh,
           double* __restrict__ cdx, double* __restrict__ cdy) {
    const int nx = lay.nx, ny = lay.ny;
    long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long total = (long long)(ny + 1) * (nx + 1);
    if (n >= total) return;
    int i = (int)(n / (nx + 1)), j = (int)(n - (long long)i * (nx + 1));
    const int W = nx + 1;
    // vertical face J=j of row i (needs nodes (i, j) low and (i+1, j) high)
    if (i < ny) {
        double hx = xn[(long long)(i + 1) * W + j], hy = yn[(long long)(i + 1) * W + j];
        double lx = xn[(long long)i * W + j], ly = yn[(long long)i * W + j];
        double ddx = hx - lx, ddy = hy - ly;
        Lv[lay.at(i, j)] = sqrt(ddx * ddx + ddy * ddy);
    }
    // horizontal face I=i of column j (high = west node (i, j), low = east node (i, j+1))
    if (j < nx) {
        double hx = xn[(long long)i * W + j], hy = yn[(long long)i * W + j];
        double lx = xn[(long long)i * W + j + 1], ly = yn[(long long)i * W + j + 1];
        double ddx = hx - lx, ddy = hy - ly;
        Lh[lay.at(i, j)] = sqrt(ddx * ddx + ddy * ddy);
    }
    if (i < ny && j < nx) {
        double xsw = xn[(long long)i * W + j], xse = xn[(long long)i * W + j + 1];
        double xnw = xn[(long long)(i + 1) * W + j], xne = xn[(long long)(i + 1) * W + j + 1];
        double ysw = yn[(long long)i * W + j], yse = yn[(long long)i * W + j + 1];
        double ynw = yn[(long long)(i + 1) * W + j], yne = yn[(long long)(i + 1) * W + j + 1];
        double xc = 0.25 * (xne + xnw + xse + xsw);
        double yc = 0.25 * (yne + ynw + yse + ysw);
        long long o = lay.at(i, j);
        const long long PL = lay.plane;
        // quadrature point of the 1-point rule: 0.5 * ((p2 - p1) * 0 + (p2 + p1)), (p1, p2) per side
        auto qp = [](double p1, double p2) { return 0.5 * ((p2 - p1) * 0.0 + (p2 + p1)); };
        dxy[0 * PL + o] = qp(xne, xse) - xc; dxy[1 * PL + o] = qp(yne, yse) - yc;   // E: (NE, SE)
        dxy[2 * PL + o] = qp(xnw, xsw) - xc; dxy[3 * PL + o] = qp(ynw, ysw) - yc;   // W: (NW, SW)
        dxy[4 * PL + o] = qp(xne, xnw) - xc; dxy[5 * PL + o] = qp(yne, ynw) - yc;   // N: (NE, NW)
        dxy[6 * PL + o] = qp(xse, xsw) - xc; dxy[7 * PL + o] = qp(yse, ysw) - yc;   // S: (SE, SW)
        // dx = E.midpoint.x - W.midpoint.x ; dy = N.midpoint.y - S.midpoint.y  (midpoint = 0.5*(high+low))
        cdx[o] = 0.5 * (xne + xse) - 0.5 * (xnw + xsw);
        cdy[o] = 0.5 * (ynw + yne) - 0.5 * (ysw + yse);
    }
}

// AoS (ny, nx, 4) <-> SoA planes
__global__ void __launch_bounds__(256)
k_aos_to_soa(Layout lay, const double* __restrict__ aos, double* __restrict__ soa) {
    long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long total = (long long)lay.nx * lay.ny;
    if (n >= total) return;
    int i = (int)(n / lay.nx), j = (int)(n - (long long)i * lay.nx);
    long long o = lay.at(i, j);
#pragma unroll
    for (int k = 0; k < 4; ++k) soa[k * lay.plane + o] = aos[4 * n + k];
}
__global__ void __launch_bounds__(256)
k_soa_to_aos(Layout lay, const double* __restrict__ soa, double* __restrict__ aos, int nplanes) {
    long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long total = (long long)lay.nx * lay.ny;
    if (n >= total) return;
    int i = (int)(n / lay.nx), j = (int)(n - (long long)i * lay.nx);
    long long o = lay.at(i, j);
    for (int k = 0; k < nplanes; ++k) aos[(long long)nplanes * n + k] = soa[k * lay.plane + o];
}
// dense (rows, cols) host-layout array -> plane
__global__ void __launch_bounds__(256)
k_dense_to_plane(Layout lay, const double* __restrict__ dense, double* __restrict__ plane, int rows, int cols) {
    long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long total = (long long)rows * cols;
    if (n >= total) return;
    int i = (int)(n / cols), j = (int)(n - (long long)i * cols);
    plane[lay.at(i, j)] = dense[n];
}
// Dirichlet inlet strips: primitive -> reconstruction variables and conservative variables
__global__ void k_dirichlet(const double* __restrict__ prim, double* __restrict__ recon, double* __restrict__ cons,
                            int len, int recon_is_prim, Consts C) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= len) return;
    double w[4] = {prim[4 * n], prim[4 * n + 1], prim[4 * n + 2], prim[4 * n + 3]};
    double U[4];
    prim2cons(w, U, C);
    for (int k = 0; k < 4; ++k) { cons[4 * n + k] = U[k]; recon[4 * n + k] = recon_is_prim ? w[k] : U[k]; }
}
__global__ void k_ghost_strip_fetch(Layout lay, const double* __restrict__ soa, int side, double* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int nx = lay.nx, ny = lay.ny;
    const int len = (side == PYH_EAST || side == PYH_WEST) ? ny : nx;
    if (idx >= len) return;
    int gi, gj;
    if (side == PYH_EAST) { gi = idx; gj = nx; }
    else if (side == PYH_WEST) { gi = idx; gj = -1; }
    else if (side == PYH_NORTH) { gi = ny; gj = idx; }
    else { gi = -1; gj = idx; }
    long long o = lay.at(gi, gj);
    for (int k = 0; k < 4; ++k) out[4 * idx + k] = soa[k * lay.plane + o];
}

}  // namespace pyh

// pyh_stage_march.cuh -- fused RK-stage kernel, row-marching version (sm_100a, fp64).
//
// One thread owns one mesh column of a strip and marches south -> north, keeping the 3-row state
// window, the previous row's north-face state and the previous row's integrated fluxes in
// registers.  Per row the CTA exchanges three small things through shared memory: the row's
// reconstruction variables (for the x-neighbours of the gradient / min-max stencil), the east-face
// states (left Riemann state of the next column) and the west-face integrated fluxes (= east-face
// flux of the previous column).  Two block barriers per row, no div/mod index arithmetic, no tile
// halo recomputation in y (one extra gradient row per strip end), two ring lanes in x.
//
// Thread t of a CTA <-> column j = blockIdx.x * (NT - 2) - 1 + t.  Lanes 1..NT-2 produce output
// cells; lane 0 only supplies the east-face state of column j0-1, lane NT-1 supplies the flux of
// the face between the last output column and its east neighbour.  Ghost columns (-1, nx) are
// carried by a lane like any other column: they publish the ghost value and, for j == nx, compute
// the block's east-edge face flux.
//
// Arithmetic is identical to k_stage_tile (same device functions, same operation order).
#pragma once
#include "pyh_layout.cuh"
#include "pyh_math.cuh"

namespace pyh {

constexpr int MARCH_MAX_THREADS = 192;

template <int FLUX, int LIM, int PRIM>
__global__ void __launch_bounds__(MARCH_MAX_THREADS, 2)
k_stage_march(const BlkDev* __restrict__ blks, Layout lay, StagePlan plan, const Control* __restrict__ ctl,
              Consts C, int tys, int want_grad_dbg) {
    if (!ctl->active) return;
    extern __shared__ double smem[];
    const int NT = blockDim.x;
    const int t = threadIdx.x;
    double* sQ = smem;                 // [2][4][NT]
    double* sFE = sQ + 8 * NT;         // [4][NT]
    double* sIW = sFE + 4 * NT;        // [4][NT]

    const BlkDev& B = blks[blockIdx.z];
    const int nx = lay.nx, ny = lay.ny;
    const long long PL = lay.plane;
    const int pitch = lay.pitch;
    const int j = (int)blockIdx.x * (NT - 2) - 1 + t;
    const int i0 = (int)blockIdx.y * tys;
    const int i1 = min(i0 + tys, ny);
    const bool act = (j >= -1) && (j <= nx);
    const bool real = (j >= 0) && (j < nx);
    const bool outcol = real && (t >= 1) && (t <= NT - 2);
    const int jc = min(max(j, -1), nx);     // clamped column for addressing
    const double* __restrict__ U = B.H[plan.cur];
    const int cart = B.cart;

    // reconstruction variables of cell (row, col); dummy for cells that do not exist (frame corners, outside)
    auto loadQ = [&](int row, int col, double q[4]) {
        bool ok = (row >= -1) && (row <= ny) && (col >= -1) && (col <= nx) &&
                  !((row == -1 || row == ny) && (col == -1 || col == nx));
        if (ok) {
            long long o = (long long)(row + 1) * pitch + PADL + col;
            q[0] = U[o]; q[1] = U[o + PL]; q[2] = U[o + 2 * PL]; q[3] = U[o + 3 * PL];
            if (PRIM) cons2prim(q, C);
        } else {
            q[0] = 1.0; q[1] = 0.0; q[2] = 0.0; q[3] = 1.0;
        }
    };
    auto apply_bc_edge = [&](int side, int idx, double c_, double s_, double q[4]) {
        int bc = B.bc[side];
        if (bc == PYH_BC_REFLECTION || bc == PYH_BC_SLIPWALL) reflect(q[1], q[2], c_, s_);
        else if (bc == PYH_BC_PRIMITIVE_DIRICHLET) {
            const double* d = B.dir_recon[side] + 4 * (long long)idx;
            q[0] = d[0]; q[1] = d[1]; q[2] = d[2]; q[3] = d[3];
        }
    };

    const int r0 = (i0 > 0) ? i0 - 1 : i0;   // first row whose gradient is needed
    double Qm[4], Qc[4], Qp[4];
    double qOut[4], qOutN[4];                // outer x-neighbour of the two ring lanes (rows r, r+1)
    const bool edge_lane = (t == 0) || (t == NT - 1);
    const int jout = (t == 0) ? j - 1 : j + 1;
    if (act) { loadQ(r0 - 1, jc, Qm); loadQ(r0, jc, Qc); loadQ(r0 + 1, jc, Qp); }
    else {
        for (int k = 0; k < 4; ++k) { Qm[k] = Qc[k] = Qp[k] = (k == 0 || k == 3) ? 1.0 : 0.0; }
    }
    if (edge_lane && real) { loadQ(r0, jout, qOut); loadQ(r0 + 1, jout, qOutN); }
    else { for (int k = 0; k < 4; ++k) { qOut[k] = qOutN[k] = (k == 0 || k == 3) ? 1.0 : 0.0; } }
#pragma unroll
    for (int k = 0; k < 4; ++k) sQ[((r0 & 1) * 4 + k) * NT + t] = Qc[k];
    __syncthreads();

    double QNp[4] = {1.0, 0.0, 0.0, 1.0};    // north-face state of row r-1
    double IWp[4] = {0, 0, 0, 0}, IEp[4] = {0, 0, 0, 0}, ISp[4] = {0, 0, 0, 0};

    for (int r = r0; r <= i1; ++r) {
        const bool rowreal = r < ny;
        const bool full = (r >= i0) && (r < i1);          // rows this strip outputs
        const bool needB = rowreal && real;               // gradient / limiter / face states of (r, j)
        const long long o = (long long)(r + 1) * pitch + PADL + jc;

        // prefetch row r+2 (window of the next iteration) and publish row r+1 for the x-neighbours
        double Qn[4], qOutNN[4];
        const bool nextB = (r + 1 <= i1) && (r + 1 < ny);
        if (act && nextB) loadQ(r + 2, jc, Qn);
        else { Qn[0] = 1.0; Qn[1] = 0.0; Qn[2] = 0.0; Qn[3] = 1.0; }
        if (edge_lane && real && nextB) loadQ(r + 2, jout, qOutNN);
        else { qOutNN[0] = 1.0; qOutNN[1] = 0.0; qOutNN[2] = 0.0; qOutNN[3] = 1.0; }
#pragma unroll
        for (int k = 0; k < 4; ++k) sQ[((((r + 1) & 1)) * 4 + k) * NT + t] = Qp[k];

        // ---- B(r): Green-Gauss gradient, limiter, limited face states ---------------------------
        double QE[4], QW[4], QN[4], QS[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { QE[k] = QW[k] = QN[k] = QS[k] = Qc[k]; }
        if (needB) {
            const long long oE = o + 1, oN = o + pitch;
            double LE = B.Lv[oE], LW = B.Lv[o], LN = B.Lh[oN], LS = B.Lh[o];
            double xlE = LE * B.cv[oE], xlW = LW * (-B.cv[o]), xlN = LN * B.ch[oN], xlS = LS * (-B.ch[o]);
            double ylE = LE * B.sv[oE], ylW = LW * (-B.sv[o]), ylN = LN * B.sh[oN], ylS = LS * (-B.sh[o]);
            double ia = 1.0 / B.A[o];
            double dx[4], dy[4];
#pragma unroll
            for (int f = 0; f < 4; ++f) { dx[f] = B.dxy[(2 * f) * PL + o]; dy[f] = B.dxy[(2 * f + 1) * PL + o]; }
            const double* sq = sQ + (r & 1) * 4 * NT;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                double q = Qc[k], qS = Qm[k], qN = Qp[k];
                double qW = (t == 0) ? qOut[k] : sq[k * NT + t - 1];
                double qE = (t == NT - 1) ? qOut[k] : sq[k * NT + t + 1];
                double fE = 0.5 * (q + qE), fW = 0.5 * (qW + q), fN = 0.5 * (q + qN), fS = 0.5 * (qS + q);
                double gx = (fE * xlE + fW * xlW + fN * xlN + fS * xlS) * ia;
                double gy = (fE * ylE + fW * ylW + fN * ylN + fS * ylS) * ia;
                double mx = dmax2(dmax2(dmax2(dmax2(q, qW), qE), qS), qN);
                double mn = dmin2(dmin2(dmin2(dmin2(q, qW), qE), qS), qN);
                double dmx = mx - q, dmn = mn - q;
                double term[4];
                double phi = 0.0;
#pragma unroll
                for (int f = 0; f < 4; ++f) {
                    term[f] = gx * dx[f] + gy * dy[f];
                    double davg = (q + term[f]) - q;
                    double pf = limiter_fn<LIM>(slope_of(dmx, dmn, davg));
                    phi = (f == 0) ? pf : dmin2(phi, pf);
                }
                if (phi < 0.0) phi = 0.0;
                QE[k] = q + phi * term[0];
                QW[k] = q + phi * term[1];
                QN[k] = q + phi * term[2];
                QS[k] = q + phi * term[3];
                if (want_grad_dbg && full && outcol) {
                    B.dbgG[k * PL + o] = gx; B.dbgG[(4 + k) * PL + o] = gy; B.dbgG[(8 + k) * PL + o] = phi;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) sFE[k * NT + t] = QE[k];
        __syncthreads();   // S1: sFE(r) and sQ(r+1) visible

        // ---- C(r): vertical face J = j of row r (lanes 1.., columns 0..nx) ----------------------------
        double IW[4] = {0, 0, 0, 0};
        if (full && (t >= 1) && (j >= 0) && (j <= nx)) {
            double cf = B.cv[o], sf = B.sv[o], Lf = B.Lv[o];
            double QL[4], QR[4];
            if (j > 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k) QL[k] = sFE[k * NT + t - 1];
            } else if (B.bc[PYH_WEST] == PYH_BC_NONE) {
                const double* sq = sQ + (r & 1) * 4 * NT;
#pragma unroll
                for (int k = 0; k < 4; ++k) QL[k] = sq[k * NT + t - 1];      // ghost cell (r, -1)
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) QL[k] = QW[k];
                apply_bc_edge(PYH_WEST, r, cf, sf, QL);
            }
            if (j < nx) {
#pragma unroll
                for (int k = 0; k < 4; ++k) QR[k] = QW[k];
            } else if (B.bc[PYH_EAST] == PYH_BC_NONE) {
#pragma unroll
                for (int k = 0; k < 4; ++k) QR[k] = Qc[k];                   // ghost cell (r, nx)
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) QR[k] = sFE[k * NT + t - 1];     // east-face state of cell (r, nx-1)
                apply_bc_edge(PYH_EAST, r, cf, sf, QR);
            }
            if (!cart) { rot(QL[1], QL[2], cf, sf); rot(QR[1], QR[2], cf, sf); }
            if (!PRIM) { cons2prim(QL, C); cons2prim(QR, C); }
            double F[4];
            riemann_flux<FLUX>(QL, QR, F, C);
            if (!cart) unrot(F[1], F[2], cf, sf);
#pragma unroll
            for (int k = 0; k < 4; ++k) IW[k] = Lf * (2.0 * F[k]);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) sIW[k * NT + t] = IW[k];

        // ---- C(r): horizontal face I = r of column j (south face of row r), output columns only -------
        double IS[4] = {0, 0, 0, 0};
        if (outcol && (r >= i0)) {
            double cf = B.ch[o], sf = B.sh[o], Lf = B.Lh[o];
            double QL[4], QR[4];
            if (r > 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k) QL[k] = QNp[k];
            } else if (B.bc[PYH_SOUTH] == PYH_BC_NONE) {
#pragma unroll
                for (int k = 0; k < 4; ++k) QL[k] = Qm[k];                   // ghost cell (-1, j)
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) QL[k] = QS[k];
                apply_bc_edge(PYH_SOUTH, j, cf, sf, QL);
            }
            if (r < ny) {
#pragma unroll
                for (int k = 0; k < 4; ++k) QR[k] = QS[k];
            } else if (B.bc[PYH_NORTH] == PYH_BC_NONE) {
#pragma unroll
                for (int k = 0; k < 4; ++k) QR[k] = Qc[k];                   // ghost cell (ny, j)
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) QR[k] = QNp[k];
                apply_bc_edge(PYH_NORTH, j, cf, sf, QR);
            }
            if (cart) { rot90(QL[1], QL[2]); rot90(QR[1], QR[2]); }
            else { rot(QL[1], QL[2], cf, sf); rot(QR[1], QR[2], cf, sf); }
            if (!PRIM) { cons2prim(QL, C); cons2prim(QR, C); }
            double F[4];
            riemann_flux<FLUX>(QL, QR, F, C);
            if (cart) unrot90(F[1], F[2]); else unrot(F[1], F[2], cf, sf);
#pragma unroll
            for (int k = 0; k < 4; ++k) IS[k] = Lf * (2.0 * F[k]);
        }

        // ---- D(r-1): residual + RK partial sums of cell (r-1, j) ------------------------------------------
        if (outcol && (r - 1 >= i0)) {
            const long long om = o - pitch;
            double a = B.A[om];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                double R = 0.5 * (IWp[k] - IEp[k] + ISp[k] - IS[k]) / a;
                for (int q = 0; q < plan.ntargets; ++q) {
                    const RkTarget& tg = plan.t[q];
                    if (tg.dst == 2) { B.dbg[k * PL + om] = R; continue; }
                    double src = (tg.src == 0) ? B.H[plan.u0][k * PL + om] : B.P[tg.row][k * PL + om];
                    double out = tg.add ? src + ctl->coef[tg.coef] * R : src;
                    if (tg.dst == 0) B.H[plan.next][k * PL + om] = out; else B.P[tg.row][k * PL + om] = out;
                }
            }
        }
        __syncthreads();   // S2: sIW(r) visible; sFE / sQ slots may be overwritten afterwards

        // carry to the next row
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            IEp[k] = (t < NT - 1) ? sIW[k * NT + t + 1] : 0.0;
            IWp[k] = IW[k];
            ISp[k] = IS[k];
            QNp[k] = QN[k];
            Qm[k] = Qc[k]; Qc[k] = Qp[k]; Qp[k] = Qn[k];
            qOut[k] = qOutN[k]; qOutN[k] = qOutNN[k];
        }
    }
}

}  // namespace pyh

// pyh_kernels.cuh -- CUDA kernels of the MUSCL residual + RK stage update (sm_100a, fp64).
//
// (the fused RK-stage kernel lives in pyh_stage_march.cuh)
// k_ghost      : ghost-strip refresh + boundary conditions (blocks/base.py:448-471, ghost.py)
// k_dt, k_dt_finalize, k_step_end : CFL reduction and the device-resident time loop control
//                (quad_block.py:423-436, solvers/base.py:114-136, Euler2D.py:195-210)
#pragma once
#include "pyh_layout.cuh"
#include <type_traits>
#include "pyh_math.cuh"

namespace pyh {

struct Tableau {
    double a[PYH_MAX_STAGES * PYH_MAX_STAGES];
    int nstages;
};

// ------------------------------------------------------------------------------------------------
// ghost strips + boundary conditions (blocks/ghost.py:187-278)
// grid: (ceil(max(nx,ny)/128), 4 sides, nblocks)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_ghost(const BlkDev* __restrict__ blks, Layout lay, PlaneOffsets po, unsigned buf, const Control* __restrict__ ctl) {
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
    const double* src = (own ? B.base : blks[B.nbr[side]].base) + buf;
    const unsigned PL = lay.plane;
    unsigned os = lay.at(si, sj), og = lay.at(gi, gj);
    double q[4] = {src[os], src[os + PL], src[os + 2 * PL], src[os + 3 * PL]};
    if (bc == PYH_BC_REFLECTION || bc == PYH_BC_SLIPWALL) {
        unsigned of = lay.at(fi, fj);
        bool vert = (side == PYH_EAST || side == PYH_WEST);
        double c_ = B.base[(vert ? po.cv : po.ch) + of], s_ = B.base[(vert ? po.sv : po.sh) + of];
        reflect(q[1], q[2], c_, s_);
    } else if (bc == PYH_BC_PRIMITIVE_DIRICHLET) {
        const double* d = B.dir_cons[side] + 4 * (long long)idx;
        q[0] = d[0]; q[1] = d[1]; q[2] = d[2]; q[3] = d[3];
    }
    double* dst = B.base + buf;
    dst[og] = q[0]; dst[og + PL] = q[1]; dst[og + 2 * PL] = q[2]; dst[og + 3 * PL] = q[3];
}

// remote halo slots: pack own edge strips / unpack received strips (ghost.py:169-241)
struct HaloSlot { int blk; int side; long long offset; };

__global__ void __launch_bounds__(128)
k_pack_halo(const BlkDev* __restrict__ blks, Layout lay, unsigned buf, const HaloSlot* __restrict__ slots, double* __restrict__ out) {
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
    unsigned o = lay.at(si, sj);
    const double* src = B.base + buf;
    double* d = out + s.offset + 4 * (long long)idx;
    d[0] = src[o]; d[1] = src[o + lay.plane]; d[2] = src[o + 2 * lay.plane]; d[3] = src[o + 3 * lay.plane];
}

__global__ void __launch_bounds__(128)
k_unpack_halo(const BlkDev* __restrict__ blks, Layout lay, unsigned buf, const HaloSlot* __restrict__ slots, const double* __restrict__ in) {
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
    unsigned o = lay.at(gi, gj);
    double* dst = B.base + buf;
    const double* d = in + s.offset + 4 * (long long)idx;
    dst[o] = d[0]; dst[o + lay.plane] = d[1]; dst[o + 2 * lay.plane] = d[2]; dst[o + 3 * lay.plane] = d[3];
}

// ------------------------------------------------------------------------------------------------
// CFL reduction + realizability (quad_block.py:423-436, states/conservative.py:161-165)
// ------------------------------------------------------------------------------------------------
constexpr int DT_ROWS = 16;

__global__ void __launch_bounds__(256)
k_dt(const BlkDev* __restrict__ blks, Layout lay, PlaneOffsets po, unsigned buf, int nblk, Control* __restrict__ ctl, Consts C, int respect_active) {
    if (respect_active && !ctl->active) return;
    const int nx = lay.nx, ny = lay.ny;
    const unsigned PL = lay.plane;
    double m = __longlong_as_double(0x7ff0000000000000ll);
    int bad = 0;
    // grid = (column chunks of blockDim.x, row groups of DT_ROWS, blocks): no per-cell index division
    const BlkDev& B = blks[blockIdx.z];
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int ibeg = blockIdx.y * DT_ROWS, iend = min(ibeg + DT_ROWS, ny);
    for (int i = ibeg; i < iend && j < nx; ++i) {
        unsigned o = lay.at(i, j);
        const double* U = B.base + buf;
        double rho = U[o], ru = U[o + PL], rv = U[o + 2 * PL], e = U[o + 3 * PL];
        if (!(rho > 0.0) || !(e > 0.0)) bad = 1;
        const double cdx = B.base[po.cdx + o], cdy = B.base[po.cdy + o];
        double tx, ty;
        auto cfl = [&](auto tag) -> bool {
            constexpr bool FAST = decltype(tag)::value;
            bool ok = true;
            typename Ar<FAST>::R rr = Ar<FAST>::recip(rho, ok);
            double u = Ar<FAST>::div(ru, rr, ok), v = Ar<FAST>::div(rv, rr, ok);
            double p = C.gm1 * (e - rho * (0.5 * (u * u + v * v)));
            double a = Ar<FAST>::sqrt(Ar<FAST>::div(C.g * p, rr, ok), ok);
            tx = Ar<FAST>::div(cdx, fabs(u) + a, ok);
            ty = Ar<FAST>::div(cdy, fabs(v) + a, ok);
            return ok;
        };
        if (!cfl(std::integral_constant<bool, true>{})) cfl(std::integral_constant<bool, false>{});
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
            if (bad) { atomicOr(&ctl->bad, 1); atomicExch(&ctl->allok, 0ull); }
        }
    }
}

// end of a step (Euler2D.py:209-210): t += dt, step counter, optional dt record
__device__ __forceinline__ void apply_step_end(Control* ctl) {
    if (!ctl->pending_end) return;
    ctl->pending_end = 0;
    if (ctl->dts && ctl->nsteps < ctl->dts_cap) ctl->dts[ctl->nsteps] = ctl->dt;
    ctl->t += ctl->dt;
    ctl->nsteps += 1;
}

// mode 0: one STEP BOUNDARY of the device-resident loop: the end of the previous step (if one ran), then Solver.get_dt's clamp
//         and the `while t < t_final` test for the next one
// mode 1: write CFL*min to *out only (pyh_local_dt / pyh_get_dt)
__global__ void k_dt_finalize(Control* ctl, double cfl, Tableau tab, int mode, double* out) {
    if (mode == 0) apply_step_end(ctl);
    double dt = cfl * dunkey(ctl->dtmin_bits);      // quad_block.py:436 ; min over blocks (and ranks) is exact
    ctl->dtmin_bits = DKEY_INF;
    if (!ctl->allok) ctl->bad = 1;                  // a rank saw an unrealizable state (Euler2D.py:144-152 aborts every rank)
    ctl->allok = 1ull;
    if (mode == 1) { *out = dt; return; }
    int active = (ctl->t < ctl->t_final) && !ctl->bad;
    ctl->active = active;
    if (!active) return;
    double rem = ctl->t_final - ctl->t;             // solvers/base.py:132-136
    dt = rem < dt ? rem : dt;
    ctl->dt = dt;
    ctl->pending_end = 1;                           // the stages that follow are this step; its end is applied at the next boundary
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

// the end of the LAST enqueued step (before the host reads the control block)
__global__ void k_step_end(Control* ctl) { apply_step_end(ctl); }

// ------------------------------------------------------------------------------------------------
// setup: derived geometry from node coordinates (exact IEEE restatement of mesh/base.py:57-64,
// mesh/quad_mesh.py:133-135,172-184, mesh/quadratures.py:56-95 for the 1-point rule)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_geometry(Layout lay, const double* __restrict__ xn, const double* __restrict__ yn,  // (ny+1, nx+1) dense
           double* __restrict__ dxy, double* __restrict__ Lv, double* __restrict__ Lh,
           double* __restrict__ cdx, double* __restrict__ cdy, double* __restrict__ pxc, double* __restrict__ pyc, int nq, Consts C) {
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
        unsigned o = lay.at(i, j);
        const size_t PL = lay.plane;
        pxc[o] = xc; pyc[o] = yc;   // QuadMesh.x / .y (quad_mesh.py:172-184), the same doubles the host computes
        // quadrature points (mesh/quadratures.py:56-95): 0.5 * ((p2 - p1) * point + (p2 + p1)), (p1, p2) per side;
        // plane ((q * 4 + f) * 2 + {x, y}) holds the offset of point q of face f from the centroid
        for (int q = 0; q < nq; ++q) {
            const double pt = C.qp[q];
            auto qp = [pt](double p1, double p2) { return 0.5 * ((p2 - p1) * pt + (p2 + p1)); };
            double* d = dxy + (size_t)(q * 8) * PL;
            d[0 * PL + o] = qp(xne, xse) - xc; d[1 * PL + o] = qp(yne, yse) - yc;   // E: (NE, SE)
            d[2 * PL + o] = qp(xnw, xsw) - xc; d[3 * PL + o] = qp(ynw, ysw) - yc;   // W: (NW, SW)
            d[4 * PL + o] = qp(xne, xnw) - xc; d[5 * PL + o] = qp(yne, ynw) - yc;   // N: (NE, NW)
            d[6 * PL + o] = qp(xse, xsw) - xc; d[7 * PL + o] = qp(yse, ysw) - yc;   // S: (SE, SW)
        }
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
    unsigned o = lay.at(i, j);
#pragma unroll
    for (int k = 0; k < 4; ++k) soa[k * (size_t)lay.plane + o] = aos[4 * n + k];
}
__global__ void __launch_bounds__(256)
k_soa_to_aos(Layout lay, const double* __restrict__ soa, double* __restrict__ aos, int nplanes) {
    long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long total = (long long)lay.nx * lay.ny;
    if (n >= total) return;
    int i = (int)(n / lay.nx), j = (int)(n - (long long)i * lay.nx);
    unsigned o = lay.at(i, j);
    for (int k = 0; k < nplanes; ++k) aos[(long long)nplanes * n + k] = soa[k * (size_t)lay.plane + o];
}
__global__ void __launch_bounds__(256)
k_fill_uniform(Layout lay, double* __restrict__ soa, double u0, double u1, double u2, double u3) {
    long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long total = (long long)lay.nx * lay.ny;
    if (n >= total) return;
    int i = (int)(n / lay.nx), j = (int)(n - (long long)i * lay.nx);
    unsigned o = lay.at(i, j);
    const size_t PL = lay.plane;
    soa[o] = u0; soa[PL + o] = u1; soa[2 * PL + o] = u2; soa[3 * PL + o] = u3;
}
// Two-state initial conditions of the reference's examples (examples/explosion_multi/initial_condition.py:53-59: np.where over
// the centroid box 3 <= x <= 7 and 3 <= y <= 7; examples/dmr/initial_condition.py:55-58: x <= 0.95): cells whose centroid lies in
// the closed box [x0, x1] x [y0, y1] get `in`, the others `out` (or stay untouched when has_out == 0).  Comparisons on the
// same centroid doubles as the host's, so the filled state equals the uploaded one bit for bit.
__global__ void __launch_bounds__(256)
k_fill_box(Layout lay, double* __restrict__ soa, const double* __restrict__ xc, const double* __restrict__ yc,
           double x0, double x1, double y0, double y1, double i0, double i1, double i2, double i3,
           int has_out, double o0, double o1, double o2, double o3) {
    long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long total = (long long)lay.nx * lay.ny;
    if (n >= total) return;
    int i = (int)(n / lay.nx), j = (int)(n - (long long)i * lay.nx);
    unsigned o = lay.at(i, j);
    const size_t PL = lay.plane;
    const double x = xc[o], y = yc[o];
    const bool inside = (x >= x0) && (x <= x1) && (y >= y0) && (y <= y1);
    if (inside) { soa[o] = i0; soa[PL + o] = i1; soa[2 * PL + o] = i2; soa[3 * PL + o] = i3; }
    else if (has_out) { soa[o] = o0; soa[PL + o] = o1; soa[2 * PL + o] = o2; soa[3 * PL + o] = o3; }
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
    bool ok = true;
    prim2cons<false>(w, U, C, ok);
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
    unsigned o = lay.at(gi, gj);
    for (int k = 0; k < 4; ++k) out[4 * idx + k] = soa[k * (size_t)lay.plane + o];
}

}  // namespace pyh

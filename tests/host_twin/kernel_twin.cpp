// Host twin of the stage kernel: pyhype_b200/csrc/pyh_stage_march.cuh (and the ghost / geometry / layout kernels of
// pyh_kernels.cuh) compiled by g++ through tests/host_twin/shim and EXECUTED on the CPU -- one OS thread per CUDA
// thread of a thread block, __syncthreads() as a barrier, thread blocks one after another.  The driver below lays a set
// of mesh blocks out exactly as pyh_api.cu does (same Layout / PlaneOffsets / BlkDev / StagePlan), refreshes the ghost
// frame and launches one stage, so tests/test_kernel_twin.py can compare the kernel's residual, gradients, limiter,
// ghost strips and updated state with the reference fixtures without a GPU.  Test infrastructure only.
// Build: g++ -O1 -ffp-contract=off -std=c++20 -pthread -shared -fPIC -DPYH_HOST_TWIN -I tests/host_twin/shim
//        -I pyhype_b200/csrc kernel_twin.cpp
#define PYH_HOST_TWIN 1
#include <cuda_runtime.h>   // the shim
#include <algorithm>
#include <vector>
#include "pyh_kernels.cuh"
#include "pyh_plan.cuh"
#include "pyh_stage_march.cuh"
#include "pyh_stage_split.cuh"

namespace pyh {
double smem[(24 + 24 * 3) * 256];   // what `extern __shared__ double smem[]` of the kernel resolves to
}
using namespace pyh;
using pyh_host_twin::launch;

static unsigned cdivu(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

// Only the instantiations the fixtures use are compiled (each costs ~1 s of g++): every flux x reconstruction mode with
// the Venkatakrishnan limiter and one quadrature point; the other limiters and 2 / 3 points with Roe + conservative.
static MarchFn pick(int f, int l, int p, int nq) {
    if (nq == 1 && l == 0) {
        switch (2 * f + p) {
            case 0: return k_stage_march<0, 0, 0, 1>;
            case 1: return k_stage_march<0, 0, 1, 1>;
            case 2: return k_stage_march<1, 0, 0, 1>;
            case 3: return k_stage_march<1, 0, 1, 1>;
            case 4: return k_stage_march<2, 0, 0, 1>;
            case 5: return k_stage_march<2, 0, 1, 1>;
        }
    }
    if (f == 0 && p == 0 && nq == 1) {
        if (l == 1) return k_stage_march<0, 1, 0, 1>;
        if (l == 2) return k_stage_march<0, 2, 0, 1>;
        if (l == 3) return k_stage_march<0, 3, 0, 1>;
    }
    if (f == 0 && p == 0 && l == 0) {
        if (nq == 2) return k_stage_march<0, 0, 0, 2>;
        if (nq == 3) return k_stage_march<0, 0, 0, 3>;
    }
    return nullptr;
}

// the three-kernel stage of small problems (pyh_stage_split.cuh), same set of instantiations
static SplitReconFn pick_recon(int l, int p) {
    if (l == 0) return p ? k_split_recon<0, 1, 3> : k_split_recon<0, 0, 3>;
    if (p) return nullptr;
    if (l == 1) return k_split_recon<1, 0, 3>;
    if (l == 2) return k_split_recon<2, 0, 3>;
    return k_split_recon<3, 0, 3>;
}
static SplitFluxFn pick_flux(int f, int p) {
    switch (2 * f + p) {
        case 0: return k_split_flux<0, 0, 1>;
        case 1: return k_split_flux<0, 1, 1>;
        case 2: return k_split_flux<1, 0, 1>;
        case 3: return k_split_flux<1, 1, 1>;
        case 4: return k_split_flux<2, 0, 1>;
        default: return k_split_flux<2, 1, 1>;
    }
}

static long long g_split_stages = 0;   // stages that went through the three-kernel path (tests assert the path was really taken)

extern "C" {

long long twin_split_stages() { return g_split_stages; }

int twin_kernel_fold_pow2() { return PYH_FOLD_POW2; }

// Everything pyh_create / pyh_add_block / pyh_upload_state set up, on the host: slabs laid out by plan_offsets, geometry
// planes, Dirichlet strips, BlkDev records, the state in H[0].  All arrays are dense, block after block: nodes
// (ny+1, nx+1); area (ny, nx); cos_v / sin_v (ny, nx+1); cos_h / sin_h (ny+1, nx); nbr / bc (4 per block: E, W, N, S;
// nbr = local block index or -1); dirichlet: 4 strips per block of (max(nx, ny), 4) primitive inlet states (read only
// where bc says so); U (ny, nx, 4).
struct Twin {
    Layout lay;
    Consts C;
    PlaneOffsets po;
    Control ctl;
    int nx, ny, nblk, nq, nt, tys, prim, mlen;
    size_t nc;
    std::vector<std::vector<double>> slabs, dbg, dbgG, dirr, dirc, aux;
    std::vector<BlkDev> blks;
    MarchFn fn;
    SplitReconFn fn_recon = nullptr;
    SplitFluxFn fn_flux = nullptr;
    bool splitpath = false;

    int setup(int flux, int lim, int prim_, int nq_, int nx_, int ny_, int nblk_, int nt_, int tys_, double gamma, int S, const double* tab,
              const double* nodes_x, const double* nodes_y, const double* area, const double* cos_v, const double* sin_v,
              const double* cos_h, const double* sin_h, const int* nbr, const int* bc, const int* is_cart, const double* dirichlet,
              const double* U) {
        nx = nx_; ny = ny_; nblk = nblk_; nq = nq_; nt = nt_; tys = tys_; prim = prim_;
        if (nt < 6 || nt > 256 || nq < 1 || nq > 3 || tys < 1 || S < 1 || S > PYH_MAX_STAGES) return -1;
        fn = pick(flux, lim, prim, nq);
        if (!fn) return -2;   // instantiation not compiled into the twin
        if (const char* e = getenv("PYH_TWIN_SPLITPATH")) splitpath = atoi(e) != 0 && nq == 1;   // pyh_api.cu: choose_split
        if (splitpath) {
            fn_recon = pick_recon(lim, prim);
            fn_flux = pick_flux(flux, prim);
            if (!fn_recon || !fn_flux) return -2;
        }
        lay.nx = nx; lay.ny = ny;
        lay.pitch = ((nx + PADL + 1 + 3) / 4) * 4;                     // pyh_create
        lay.plane = (unsigned)((ny + 2) * lay.pitch);
        C.g = gamma; C.gm1 = gamma - 1.0; C.k = 1.0 / (gamma - 1.0); C.gm = gamma / (gamma - 1.0);
        for (int q = 0; q < 3; ++q) { C.qw[q] = 0.0; C.qp[q] = 0.0; }
        if (nq == 1) { C.qp[0] = 0.0; C.qw[0] = 2.0; }
        else if (nq == 2) { C.qp[0] = -1.0 / std::sqrt(3.0); C.qp[1] = 1.0 / std::sqrt(3.0); C.qw[0] = C.qw[1] = 1.0; }
        else { C.qp[0] = -std::sqrt(3.0 / 5.0); C.qp[1] = 0.0; C.qp[2] = std::sqrt(3.0 / 5.0); C.qw[0] = 5.0 / 9.0; C.qw[1] = 8.0 / 9.0; C.qw[2] = 5.0 / 9.0; }
        bool need_acc[PYH_MAX_STAGES];
        plan_need_acc(tab, S, need_acc);
        po = plan_offsets(lay.plane, S, nq, need_acc);                  // the product's slab layout (pyh_plan.cuh)
        const size_t nn = (size_t)(ny + 1) * (nx + 1), nv = (size_t)ny * (nx + 1), nh = (size_t)(ny + 1) * nx;
        nc = (size_t)ny * nx;
        mlen = std::max(nx, ny);
        aux.assign(nblk, {});
        slabs.assign(nblk, {}); dbg.assign(nblk, {}); dbgG.assign(nblk, {}); dirr.assign(nblk * 4, {}); dirc.assign(nblk * 4, {});
        blks.assign(nblk, BlkDev());
        std::memset(&ctl, 0, sizeof(ctl));
        ctl.active = 1;
        for (int b = 0; b < nblk; ++b) {
            slabs[b].assign((size_t)po.nplanes * lay.plane, 0.0);
            dbg[b].assign(4 * (size_t)lay.plane, 0.0);
            dbgG[b].assign(12 * (size_t)lay.plane, 0.0);
            double* slab = slabs[b].data();
            auto put = [&](const double* host, double* plane, int rows, int cols) {
                const long long n = (long long)rows * cols;
                launch(dim3(cdivu(n, 256)), 256, false, [&] { k_dense_to_plane(lay, host, plane, rows, cols); });
            };
            put(area + b * nc, slab + po.A, ny, nx);
            put(cos_v + b * nv, slab + po.cv, ny, nx + 1);
            put(sin_v + b * nv, slab + po.sv, ny, nx + 1);
            put(cos_h + b * nh, slab + po.ch, ny + 1, nx);
            put(sin_h + b * nh, slab + po.sh, ny + 1, nx);
            launch(dim3(cdivu((long long)nn, 256)), 256, false, [&] {
                k_geometry(lay, nodes_x + b * nn, nodes_y + b * nn, slab + po.dxy, slab + po.Lv, slab + po.Lh, slab + po.cdx, slab + po.cdy, slab + po.xc, slab + po.yc, nq, C);
            });
            BlkDev& D = blks[b];
            std::memset(&D, 0, sizeof(D));
            D.base = slab;
            D.dbg = dbg[b].data();
            D.dbgG = dbgG[b].data();
            if (splitpath) { aux[b].assign((size_t)kSplitPlanes * lay.plane, 0.0); D.aux = aux[b].data(); D.aux_fx = D.aux + (size_t)kSplitStatePlanes * lay.plane; }
            for (int s = 0; s < 4; ++s) {
                D.bc[s] = bc[4 * b + s];
                D.nbr[s] = nbr[4 * b + s];
                D.remote_slot[s] = -1;
                if (D.bc[s] == PYH_BC_PRIMITIVE_DIRICHLET) {
                    const int len = (s == PYH_EAST || s == PYH_WEST) ? ny : nx;
                    const double* pr = dirichlet + ((size_t)(4 * b + s) * mlen) * 4;
                    dirr[4 * b + s].assign(4 * (size_t)len, 0.0);
                    dirc[4 * b + s].assign(4 * (size_t)len, 0.0);
                    double *rr = dirr[4 * b + s].data(), *cc = dirc[4 * b + s].data();
                    launch(dim3(cdivu(len, 128)), 128, false, [&] { k_dirichlet(pr, rr, cc, len, prim, C); });
                    D.dir_recon[s] = rr;
                    D.dir_cons[s] = cc;
                }
            }
            D.cart = is_cart[b] ? 1 : 0;
            {   // pyh_add_block: bit 1 = every vertical face axis-aligned
                bool unit = true;
                for (size_t i = 0; i < nv && unit; ++i) unit = (cos_v[b * nv + i] == 1.0) && (sin_v[b * nv + i] == 0.0);
                if (unit) D.cart |= 2;
            }
            D.gid = b;
            launch(dim3(cdivu((long long)nc, 256)), 256, false, [&] { k_aos_to_soa(lay, U + b * nc * 4, slab + po.H[0]); });
        }
        return 0;
    }
    void ghost(int buf) {   // do_ghost
        launch(dim3(cdivu(mlen, 128), 4, nblk), 128, false, [&] { k_ghost(blks.data(), lay, po, po.H[buf], &ctl); });
    }
    void stage(const StagePlan& plan, int want_grad_dbg) {   // launch_stage: the product's tile plan (PYH_TWIN_SPLIT = 1: north /
        // south edge strips apart, 2: east / west edge columns apart, 3: both -- what a context with remote neighbours launches)
        if (splitpath && !want_grad_dbg && !plan.write_residual) {   // launch_stage_split: recon -> flux -> update
            launch(dim3(cdivu(nx, kSplitTX), cdivu(ny, kSplitTY), nblk), kSplitReconThreads, true,
                   [&] { fn_recon(blks.data(), lay, po, plan.cur, &ctl, C); });
            const long long nfaces = std::max((long long)(nx + 1) * ny, (long long)nx * (ny + 1));
            launch(dim3(cdivu(nfaces, kSplitFluxThreads), 2, nblk), kSplitFluxThreads, false,
                   [&] { fn_flux(blks.data(), lay, po, plan.cur, &ctl, C); });
            launch(dim3(cdivu((long long)nx * ny, kSplitUpdateThreads), 1, nblk), kSplitUpdateThreads, true,
                   [&] { k_split_update<5>(blks.data(), lay, po, plan, &ctl, &ctl, C); });
            ++g_split_stages;
            return;
        }
        const char* e = getenv("PYH_TWIN_SPLIT");
        const int split = e ? atoi(e) : 0;
        TileLaunch tl[3];
        const int n = plan_tiles(nx, ny, nt, tys, (split & 1) != 0, (split & 2) != 0, tl);
        for (int q = n - 1; q >= 0; --q) {   // interior first, edges last: the order must not matter
            const TileLaunch t = tl[q];
            launch(dim3(t.gx, t.gy, nblk), nt, true, [&] { fn(blks.data(), lay, po, plan, &ctl, &ctl, C, t.tys, want_grad_dbg, t.tiles); });
        }
    }
    void fetch_state(int buf, double* out) const {
        for (int b = 0; b < nblk; ++b)
            for (int i = 0; i < ny; ++i)
                for (int j = 0; j < nx; ++j)
                    for (int k = 0; k < 4; ++k)
                        out[((size_t)b * nc + (size_t)i * nx + j) * 4 + k] = slabs[b][po.H[buf] + k * (size_t)lay.plane + lay.at(i, j)];
    }
};

// One ghost refresh + one stage launch (residual test hook on, one RK target: Unew = U + coef * R).
// Outputs: R, Unew (ny, nx, 4); G (12, ny, nx) = gx[4], gy[4], phi[4]; ghost (4 sides, max(nx, ny), 4).
int twin_stage(int flux, int lim, int prim, int nq, int nx, int ny, int nblk, int nt, int tys, double gamma, double coef,
               const double* nodes_x, const double* nodes_y, const double* area, const double* cos_v, const double* sin_v,
               const double* cos_h, const double* sin_h, const int* nbr, const int* bc, const int* is_cart, const double* dirichlet,
               const double* U, double* R, double* Unew, double* G, double* ghost) {
    double tab[PYH_MAX_STAGES * PYH_MAX_STAGES] = {1.0};
    Twin T;
    int rc = T.setup(flux, lim, prim, nq, nx, ny, nblk, nt, tys, gamma, 1, tab, nodes_x, nodes_y, area, cos_v, sin_v, cos_h, sin_h, nbr, bc,
                     is_cart, dirichlet, U);
    if (rc) return rc;
    T.ctl.coef[0] = coef;
    T.ghost(0);
    StagePlan plan = plan_stage(tab, 1, T.po, 0, 0, 0, 1);
    plan.write_residual = 1;
    T.stage(plan, 1);
    const Layout& lay = T.lay;
    const size_t nc = T.nc;
    const int mlen = T.mlen;
    T.fetch_state(1, Unew);
    for (int b = 0; b < nblk; ++b) {
        const double* slab = T.slabs[b].data();
        for (int i = 0; i < ny; ++i)
            for (int j = 0; j < nx; ++j) {
                const unsigned o = lay.at(i, j);
                const size_t c = ((size_t)b * nc + (size_t)i * nx + j) * 4;
                for (int k = 0; k < 4; ++k) R[c + k] = T.dbg[b][k * (size_t)lay.plane + o];
                for (int k = 0; k < 12; ++k) G[((size_t)b * 12 + k) * nc + (size_t)i * nx + j] = T.dbgG[b][k * (size_t)lay.plane + o];
            }
        for (int s = 0; s < 4; ++s) {
            const int len = (s == PYH_EAST || s == PYH_WEST) ? ny : nx;
            for (int idx = 0; idx < len; ++idx) {
                int gi, gj;
                if (s == PYH_EAST) { gi = idx; gj = nx; }
                else if (s == PYH_WEST) { gi = idx; gj = -1; }
                else if (s == PYH_NORTH) { gi = ny; gj = idx; }
                else { gi = -1; gj = idx; }
                for (int k = 0; k < 4; ++k) ghost[(((size_t)b * 4 + s) * mlen + idx) * 4 + k] = slab[T.po.H[0] + k * (size_t)lay.plane + lay.at(gi, gj)];
            }
        }
    }
    return 0;
}

// `nsteps` whole time steps of an S-stage tableau (row-major PYH_MAX_STAGES x PYH_MAX_STAGES) with given dt's, driven like
// pyh_step: k_set_dt's coefficient table, then per stage plan_next_buffer / plan_stage (pyh_plan.cuh), the stage kernel
// and the ghost refresh of the buffer it wrote.  The CFL reduction (warp shuffles) is not emulated: dt comes from the caller.
int twin_steps(int flux, int lim, int prim, int nq, int nx, int ny, int nblk, int nt, int tys, double gamma, int S, const double* tab,
               int nsteps, const double* dts, const double* nodes_x, const double* nodes_y, const double* area, const double* cos_v,
               const double* sin_v, const double* cos_h, const double* sin_h, const int* nbr, const int* bc, const int* is_cart,
               const double* dirichlet, const double* U, double* Uout) {
    Twin T;
    int rc = T.setup(flux, lim, prim, nq, nx, ny, nblk, nt, tys, gamma, S, tab, nodes_x, nodes_y, area, cos_v, sin_v, cos_h, sin_h, nbr, bc,
                     is_cart, dirichlet, U);
    if (rc) return rc;
    int i0 = 0, i1 = 1, i2 = 2, cur = 0;      // Ctx::i0, i1, i2, cur
    T.ghost(i0);                               // pyh_apply_bc after the upload
    for (int n = 0; n < nsteps; ++n) {
        for (int s = 0; s < S; ++s)            // k_set_dt (explicit_runge_kutta.py:71: dt * a[s][k] formed first)
            for (int k = 0; k <= s; ++k) T.ctl.coef[s * PYH_MAX_STAGES + k] = dts[n] * tab[s * PYH_MAX_STAGES + k];
        // ghost refresh: by the stage kernel itself (plan.push_ghost, the product's default) on even steps, by k_ghost behind
        // the stage (topologies the push cannot serve, PYH_NO_PUSH_GHOST) on odd steps -- both must give the reference's state
        const bool push = getenv("PYH_TWIN_NO_PUSH") ? false : (n % 2 == 0);
        for (int s = 0; s < S; ++s) {          // do_stage + do_ghost
            cur = (s == 0) ? i0 : cur;
            const int next = plan_next_buffer(S, s, cur, i0, i1, i2);
            StagePlan pl = plan_stage(tab, S, T.po, i0, s, cur, next);
            pl.push_ghost = push ? 1 : 0;
            T.stage(pl, 0);
            cur = next;
            if (s == S - 1) {
                if (S == 1) std::swap(i0, i1);
                cur = i0;
            }
            if (!push) T.ghost(cur);
        }
    }
    T.fetch_state(i0, Uout);
    return 0;
}

// The device-resident time loop of pyh_run, kernel for kernel: per step k_dt (CFL minimum + realizability, warp shuffles
// and all) -> k_dt_finalize (clamp to t_final - t, dt * a[s][k] table, active flag) -> stages + ghost refreshes ->
// k_step_end; every kernel early-exits once t >= t_final.  Outputs the dt sequence, the final time / step count / flag.
int twin_run(int flux, int lim, int prim, int nq, int nx, int ny, int nblk, int nt, int tys, double gamma, double cfl, int S,
             const double* tab, double t0, double t_final, int max_steps, const double* nodes_x, const double* nodes_y, const double* area,
             const double* cos_v, const double* sin_v, const double* cos_h, const double* sin_h, const int* nbr, const int* bc,
             const int* is_cart, const double* dirichlet, const double* U, double* Uout, double* dts_out, double* t_out, int* nsteps_out,
             int* bad_out) {
    Twin T;
    int rc = T.setup(flux, lim, prim, nq, nx, ny, nblk, nt, tys, gamma, S, tab, nodes_x, nodes_y, area, cos_v, sin_v, cos_h, sin_h, nbr, bc,
                     is_cart, dirichlet, U);
    if (rc) return rc;
    Tableau tb;
    std::memset(&tb, 0, sizeof(tb));
    tb.nstages = S;
    for (int i = 0; i < PYH_MAX_STAGES * PYH_MAX_STAGES; ++i) tb.a[i] = tab[i];
    Control& ctl = T.ctl;
    ctl.t = t0; ctl.t_final = t_final; ctl.dtmin_bits = DKEY_INF; ctl.allok = 1ull; ctl.active = 1; ctl.nsteps = 0; ctl.bad = 0;
    ctl.dts = dts_out; ctl.dts_cap = max_steps;
    int i0 = 0, i1 = 1, i2 = 2, cur = 0;
    T.ghost(i0);
    auto dt_kernel = [&](int buf, int respect_active) {   // launch_dt
        launch(dim3(cdivu(nx, 256), cdivu(ny, DT_ROWS), nblk), 256, true,
               [&] { k_dt(T.blks.data(), T.lay, T.po, T.po.H[buf], nblk, &ctl, T.C, respect_active); });
    };
    // pyh_run: the CFL minimum of the starting state by k_dt, afterwards by the last stage of every step (plan.fuse_dt)
    dt_kernel(i0, 0);
    for (int n = 0; n < max_steps; ++n) {                 // enqueue_step
        k_dt_finalize(&ctl, cfl, tb, 0, nullptr);
        for (int s = 0; s < S; ++s) {
            cur = (s == 0) ? i0 : cur;
            const int next = plan_next_buffer(S, s, cur, i0, i1, i2);
            StagePlan pl = plan_stage(tab, S, T.po, i0, s, cur, next);
            pl.fuse_dt = (s == S - 1) ? 1 : 0;
            pl.push_ghost = 1;                            // no k_ghost between the stages: the stage kernel refreshes the frames
            T.stage(pl, 0);
            cur = next;
            if (s == S - 1) {
                if (S == 1) std::swap(i0, i1);
                cur = i0;
            }
        }
        k_step_end(&ctl);                                 // (the product flushes the step end only before it polls: same arithmetic)
        if (!ctl.active || ctl.bad || !(ctl.t < ctl.t_final)) break;
    }
    double tmp = 0.0;                                     // final realizability check: the flag the last step reduced
    k_dt_finalize(&ctl, cfl, tb, 1, &tmp);
    T.fetch_state(i0, Uout);
    *t_out = ctl.t;
    *nsteps_out = (int)ctl.nsteps;
    *bad_out = ctl.bad;
    return 0;
}

// pyh_plan.cuh: plan_tiles as data, for tests/test_kernel_twin.py (coverage and edge-before-exchange invariants).
// out: per launch {row0, rowstride, row1, xfirst, xstride, gx, gy, tys, edge}; returns the number of launches.
int twin_plan_tiles(int nx, int ny, int nt, int tys, int split_ns, int split_ew, int* out) {
    TileLaunch tl[3];
    const int n = plan_tiles(nx, ny, nt, tys, split_ns != 0, split_ew != 0, tl);
    for (int q = 0; q < n; ++q) {
        const int v[9] = {tl[q].tiles.row0, tl[q].tiles.rowstride, tl[q].tiles.row1, tl[q].tiles.xfirst, tl[q].tiles.xstride,
                          (int)tl[q].gx, (int)tl[q].gy, tl[q].tys, tl[q].edge};
        for (int i = 0; i < 9; ++i) out[9 * q + i] = v[i];
    }
    return n;
}

}  // extern "C"

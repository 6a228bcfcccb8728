// pyh_api.cu -- context, memory management and the extern "C" entry points of include/pyh_b200.h
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdarg>
#include <vector>
#include <map>
#include <algorithm>
#include <string>

#include "pyh_kernels.cuh"
#include "pyh_march_tu.cuh"
#include "pyh_plan.cuh"
#include "pyh_comm.cuh"

using namespace pyh;

static thread_local char g_err[512] = "";
static int set_err(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define NC(call)                                                                                     \
    do {                                                                                             \
        ncclResult_t r_ = (call);                                                                    \
        if (r_ != 0)                                                                                 \
            return set_err(PYH_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,        \
                           nccl_api().GetErrorString ? nccl_api().GetErrorString(r_) : "NCCL error"); \
    } while (0)
#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return set_err(PYH_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,        \
                           cudaGetErrorString(e_));                                                  \
    } while (0)

// The stage kernel is instantiated per (flux, limiter, reconstruction, quadrature points) in three translation
// units (pyh_march_nq{1,2,3}.cu, compiled in parallel); each exports its picker.
namespace pyh {
MarchFn pick_march_nq1(int f, int l, int p);
MarchFn pick_march_nq2(int f, int l, int p);
MarchFn pick_march_nq3(int f, int l, int p);
// pyh_split.cu: the three-kernel stage of small problems (pyh_stage_split.cuh)
SplitReconFn pick_split_recon(int l, int p, bool dense);
SplitFluxFn pick_split_flux(int f, int p, bool dense);
struct SplitLaunchOpts { bool pdl; const void* win_base; size_t win_bytes; float win_hit; };   // PDL chaining + persisting-L2 window over the scratch planes
cudaError_t launch_split_recon(SplitReconFn fn, dim3 grid, cudaStream_t st, SplitLaunchOpts pdl, const BlkDev* blks, const Layout lay, const PlaneOffsets po,
                               const unsigned cur, const Control* ctl, const Consts C);
cudaError_t launch_split_flux(SplitFluxFn fn, dim3 grid, cudaStream_t st, SplitLaunchOpts pdl, const BlkDev* blks, const Layout lay, const PlaneOffsets po,
                              const unsigned cur, const Control* ctl, const Consts C);
cudaError_t launch_split_update(dim3 grid, cudaStream_t st, SplitLaunchOpts pdl, bool dense, const BlkDev* blks, const Layout lay, const PlaneOffsets po,
                                const StagePlan plan, const Control* ctl, Control* ctl_out, const Consts C);
}

namespace {

struct HostBlock {
    pyh_block_desc d;   // pointers inside are NOT retained
    BlkDev dev;
    std::vector<void*> allocs;   // slab + Dirichlet strips (+ on-demand debug buffers)
};

struct Ctx {
    pyh_config cfg;
    Layout lay;
    PlaneOffsets po;
    Consts C;
    Tableau tab;
    cudaStream_t stream = nullptr;
    std::vector<HostBlock> blocks;
    std::map<int, int> gid2idx;
    BlkDev* d_blks = nullptr;
    Control* d_ctl = nullptr;
    HaloSlot* d_slots = nullptr;
    std::vector<HaloSlot> slots;
    std::vector<int> slot_nbr_gid;
    long long halo_doubles = 0;
    bool finalized = false;
    int i0 = 0, i1 = 1, i2 = 2;   // roles of the three haloed buffers; H[i0] holds the current solution
    int cur = 0;                   // buffer read by the next stage
    int stage_next = 0;
    bool need_acc[PYH_MAX_STAGES];
    long long launches = 0;
    double* d_scratch = nullptr;   // staging for uploads/downloads
    size_t scratch_bytes = 0;
    double* d_dts = nullptr;
    long long dts_cap = 0;
    double* d_tmp = nullptr;       // small device scratch (dt etc.)
    int march_nt = 128, march_tys = 64;
    bool march_configured = false;   // cudaFuncSetAttribute done for this context's device
    bool use_split = false;          // every stage as recon / flux / update kernels (pyh_stage_split.cuh) instead of the fused one
    bool split_ready = false;        // scratch planes of the split path allocated (eligible context)
    bool path_tuned = false;         // tune_stage_path has run (first pyh_run)
    double tune_ms[2] = {0.0, 0.0};  // what it measured: ms per stage launch, fused / split
    double* d_aux = nullptr;         // scratch planes of the split path, all blocks in ONE allocation (so that one L2 access-policy window covers them)
    size_t aux_bytes = 0;
    float aux_hit_ratio = 0.f;       // > 0: the split kernels are launched with a persisting-L2 window over [win_base, win_base + win_bytes)
    const void* win_base = nullptr;
    size_t win_bytes = 0;
    bool push_ok = false;            // ghost cells are written by the stage kernel itself (plan.push_ghost): no k_ghost / k_pack_halo per stage
    // asynchronous state streaming (pyh_upload_state_async & co)
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in_done = nullptr, ev_in_consumed = nullptr, ev_out_ready = nullptr;
    std::vector<cudaEvent_t> ev_out_done;   // per block: the D2H copy out of its staging area has finished
    double* d_stage_in = nullptr;           // nblocks x (ny, nx, 4)
    double* d_stage_out = nullptr;
    std::vector<char> staged;
    // overlap of the remote strip exchange with the stage kernel: edge strips + exchange on s_edge, interior on `stream`
    bool split_ns = false, split_ew = false;    // remote neighbours across north / south, east / west block edges
    cudaStream_t s_edge = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_edge_done = nullptr;
    // pyh_run: one period of the time loop (1 step, or 2 for single-stage tableaux whose buffers alternate)
    // captured once as a CUDA graph and replayed
    cudaGraphExec_t run_graph = nullptr;
    int run_graph_steps = 0;
    int run_graph_i0 = 0;          // buffer holding the solution when the captured period starts
    long long run_graph_launches = 0;
    Comm comm;                                  // multi-rank transport (pyh_comm_init); comm.comm == nullptr: single rank
};

int ensure_streaming(Ctx* c) {
    if (c->s_in) return 0;
    const size_t nb = c->blocks.size();
    const size_t per = 4 * (size_t)c->lay.nx * c->lay.ny;
    CU(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->ev_in_done, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_in_consumed, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_out_ready, cudaEventDisableTiming));
    c->ev_out_done.assign(nb, nullptr);
    for (size_t b = 0; b < nb; ++b) CU(cudaEventCreateWithFlags(&c->ev_out_done[b], cudaEventDisableTiming));
    CU(cudaMalloc(&c->d_stage_in, nb * per * sizeof(double)));
    CU(cudaMalloc(&c->d_stage_out, nb * per * sizeof(double)));
    c->staged.assign(nb, 0);
    return 0;
}

int ensure_scratch(Ctx* c, size_t bytes) {
    if (c->scratch_bytes >= bytes) return 0;
    if (c->d_scratch) cudaFree(c->d_scratch);
    c->d_scratch = nullptr;
    c->scratch_bytes = 0;
    CU(cudaMalloc(&c->d_scratch, bytes));
    c->scratch_bytes = bytes;
    return 0;
}

int dalloc(HostBlock& hb, double** p, long long ndoubles, bool zero) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, (size_t)ndoubles * sizeof(double));
    if (e != cudaSuccess) return set_err(PYH_ERR_NOMEM, "cudaMalloc of %lld doubles failed: %s", ndoubles, cudaGetErrorString(e));
    if (zero) {
        e = cudaMemset(q, 0, (size_t)ndoubles * sizeof(double));
        if (e != cudaSuccess) return set_err(PYH_ERR_CUDA, "cudaMemset failed: %s", cudaGetErrorString(e));
    }
    hb.allocs.push_back(q);
    *p = (double*)q;
    return 0;
}

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- stage kernel dispatch ------------------------------------------------------------------------
MarchFn pick_march(int f, int l, int p, int nq) {
    return nq == 1 ? pyh::pick_march_nq1(f, l, p) : (nq == 2 ? pyh::pick_march_nq2(f, l, p) : pyh::pick_march_nq3(f, l, p));
}

// Strip shape of the stage kernel: nt lanes per thread block (nt - 4 output columns) x tys rows.
// A thread block's run time is ~ (tys + 1) row times (one extra row of gradients / limiter in the prologue), and a row
// time (~8 us) hardly depends on how many of the SM's 16 warp slots are taken -- the kernel is latency-bound per warp.  So
// a launch costs about  waves x (tys + 1)  row times, waves = ceil(thread blocks / resident slots).  Large problems (many
// waves) keep the measured optimum nt = widest, tys = 64 (profiles/r01j_tuning_notes.md); problems of a few waves or less
// -- the reference's own examples: explosion_multi is 8 x 150^2, DMR 4 x 500^2 -- search (nt, tys) for the cheapest shape,
// which above all avoids a sparsely filled second wave (DMR: 2 waves x 13 rows -> 1 wave x 18).
void choose_march_shape(Ctx* c) {
    const int nx = c->lay.nx, ny = c->lay.ny;
    const long long nblk = (long long)std::max<size_t>(c->blocks.size(), 1);
    const int nq = c->cfg.num_quadrature_points;
    MarchFn fn = pick_march(c->cfg.flux, c->cfg.limiter, c->cfg.recon, nq);
    const int maxt = march_max_threads(nq);
    cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, march_smem_doubles(nq) * maxt * (int)sizeof(double));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->cfg.device);
    auto slots_for = [&](int n) -> long long {
        int occ = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)fn, n, (size_t)march_smem_doubles(nq) * n * sizeof(double));
        if (e != cudaSuccess || occ < 1) { cudaGetLastError(); occ = std::max(1, 512 / n); }
        return (long long)occ * sms;
    };
    // default: the width that wastes the fewest lanes for this nx, 64 rows (capped so that there are thread blocks for ~6 waves)
    const int cand[] = {64, 96, 128, 160, 192};
    double best = -1.0;
    int nt = 128;
    for (int n : cand) {
        if (n > 128) continue;   // large problems: 4 thread blocks of 128 threads per SM is the measured optimum
        int strips = (nx + n - 5) / (n - 4);
        double util = (double)nx / ((double)strips * n);
        if (util > best + 1e-9 || (util > best - 1e-9 && n > nt)) { best = util; nt = n; }
    }
    int tys = 64;
    {
        long long nsx = (nx + nt - 5) / (nt - 4);
        long long per_row_strip = nsx * nblk;
        long long want_nsy = (sms * 6 + per_row_strip - 1) / per_row_strip;
        if (want_nsy < 1) want_nsy = 1;
        int cap = (int)((ny + want_nsy - 1) / want_nsy);
        tys = std::max(4, std::min(tys, cap));
    }
    auto n_ctas = [&](int n, int ty) { return (long long)((nx + n - 5) / (n - 4)) * nblk * ((ny + ty - 1) / ty); };
    if (tys == 64 && ny > 100) {
        // large problem: rows per strip near 64 such that the thread blocks fill an integer number of waves (8 x 2048^2: 79
        // rows = 5.97 waves of 592 instead of 64 rows = 7.35; measured -1 %, profiles/r02m_rows_per_strip.txt)
        const long long slots = slots_for(nt);
        double best_cost = 1e300;
        int best_ty = tys;
        for (int ty = 48; ty <= std::min(ny, 100); ++ty) {
            const long long ctas = n_ctas(nt, ty);
            const double cost = (double)((ctas + slots - 1) / slots) * (ty + 1.05) + 1e-3 * std::abs(ty - 64);
            if (cost < best_cost) { best_cost = cost; best_ty = ty; }
        }
        tys = best_ty;
    }
    if (n_ctas(nt, tys) <= 3 * slots_for(nt)) {   // small problem: cheapest (nt, tys) by the wave model
        double best_cost = 1e300;
        for (int n : cand) {
            if (n > maxt) continue;
            const long long slots = slots_for(n);
            for (int ty = 2; ty <= std::min(ny, 128); ++ty) {
                const long long ctas = n_ctas(n, ty);
                const long long waves = (ctas + slots - 1) / slots;
                const int last = ny - ((ny + ty - 1) / ty - 1) * ty;          // rows of the last (shortest) strip: ragged strips idle lanes
                const double cost = (double)waves * (ty + 1.05) + 1e-3 * (ty - last) + 1e-4 * (128 - n);
                if (cost < best_cost) { best_cost = cost; nt = n; tys = ty; }
            }
        }
    }
    if (const char* e = getenv("PYH_MARCH_NT")) { int v = atoi(e); if (v >= 32 && v <= maxt && v % 32 == 0) nt = v; }
    if (const char* e = getenv("PYH_MARCH_TYS")) { int v = atoi(e); if (v >= 1) tys = v; }
    c->march_nt = nt;
    c->march_tys = tys;
}

int launch_stage(Ctx* c, const StagePlan& plan, int want_grad_dbg, const TileLaunch& tl, cudaStream_t st) {
    if (c->blocks.empty() || tl.gx == 0 || tl.gy == 0) return 0;   // a rank without blocks only takes part in the reductions (blocks/base.py:473-513 allows it)
    const int nq = c->cfg.num_quadrature_points;
    MarchFn fn = pick_march(c->cfg.flux, c->cfg.limiter, c->cfg.recon, nq);
    const int nt = c->march_nt;
    size_t smem = (size_t)march_smem_doubles(nq) * nt * sizeof(double);
    if (!c->march_configured) {   // per context = per device (the attribute is a per-device property of the function)
        CU(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, march_smem_doubles(nq) * march_max_threads(nq) * (int)sizeof(double)));
        if (const char* e = getenv("PYH_CARVEOUT")) CU(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e)));
        c->march_configured = true;
    }
    dim3 grid(tl.gx, tl.gy, (unsigned)c->blocks.size());
    fn<<<grid, nt, smem, st>>>(c->d_blks, c->lay, c->po, plan, c->d_ctl, c->d_ctl, c->C, tl.tys, want_grad_dbg, tl.tiles);
    CU(cudaGetLastError());
    c->launches++;
    return 0;
}

// the whole block in one launch
int launch_stage(Ctx* c, const StagePlan& plan, int want_grad_dbg) {
    TileLaunch tl[3];
    plan_tiles(c->lay.nx, c->lay.ny, c->march_nt, c->march_tys, false, false, tl);
    return launch_stage(c, plan, want_grad_dbg, tl[0], c->stream);
}

// Small and medium problems may take the three-kernel stage (pyh_stage_split.cuh) instead of the fused kernel: one quadrature
// point, no neighbours on other ranks (their strip exchange is built around the fused kernel's edge / interior launches), and
// few enough cells that the scratch planes are affordable.  Which of the two is faster depends on how well the block shape
// fits the fused kernel's strips (measured, profiles/r02o_split_stage_ab.txt: DMR's 4 x 500^2 blocks run 24 % faster split,
// 8 x 256^2 .. 8 x 1024^2 12-20 % faster fused, explosion_multi's 8 x 150^2 the same), so eligible contexts MEASURE it once, on
// their own state, at the first pyh_run (tune_stage_path); the results are bit-identical either way.  PYH_SPLIT=0 / 1 forces a
// path (A/B runs; the GPU tests run every case once per path).
bool split_eligible(Ctx* c) {
    if (c->cfg.num_quadrature_points != 1 || c->blocks.empty() || !c->slots.empty()) return false;
    if (const char* e = getenv("PYH_SPLIT")) return atoi(e) != 0;
    const long long cells = (long long)c->lay.nx * c->lay.ny * (long long)c->blocks.size();
    return cells <= kSplitMaxCells;
}

int launch_stage_split(Ctx* c, const StagePlan& plan, cudaStream_t st) {
    if (c->blocks.empty()) return 0;
    const int nx = c->lay.nx, ny = c->lay.ny;
    const unsigned nb = (unsigned)c->blocks.size();
    const bool dense = (long long)nx * ny * nb > kSplitDenseCells;   // many waves: the high-occupancy builds of the kernels (pyh_split.cu)
    SplitReconFn k1 = pick_split_recon(c->cfg.limiter, c->cfg.recon, dense);
    SplitFluxFn k2 = pick_split_flux(c->cfg.flux, c->cfg.recon, dense);
    static const bool use_pdl = getenv("PYH_NO_PDL") == nullptr;   // programmatic dependent launch (pyh_stage_split.cuh); PYH_NO_PDL=1: plain stream order (A/B)
    const SplitLaunchOpts pdl = {use_pdl, c->win_base, c->win_bytes, c->aux_hit_ratio};
    CU(launch_split_recon(k1, dim3(cdiv(nx, kSplitTX), cdiv(ny, kSplitTY), nb), st, pdl, c->d_blks, c->lay, c->po, plan.cur, c->d_ctl, c->C));
    const long long nfaces = std::max((long long)(nx + 1) * ny, (long long)nx * (ny + 1));
    CU(launch_split_flux(k2, dim3(cdiv(nfaces, kSplitFluxThreads), 2, nb), st, pdl, c->d_blks, c->lay, c->po, plan.cur, c->d_ctl, c->C));
    CU(launch_split_update(dim3(cdiv((long long)nx * ny, kSplitUpdateThreads), 1, nb), st, pdl, dense, c->d_blks, c->lay, c->po, plan, c->d_ctl, c->d_ctl, c->C));
    c->launches += 3;
    return 0;
}

StagePlan make_plan(Ctx* c, int s, int cur, int next);
int launch_stage(Ctx* c, const StagePlan& plan, int want_grad_dbg);

// Times stage 0 of a step both ways on the context's current state and keeps the faster path.  Stage 0 reads the solution
// buffer and writes only buffers every step rewrites before reading them (the next stage state, the partial sums), without
// the CFL reduction or the ghost push, so the solution, the control block and the results are untouched.
int tune_stage_path(Ctx* c) {
    if (c->path_tuned) return 0;
    c->path_tuned = true;
    if (!c->split_ready || getenv("PYH_SPLIT")) return 0;   // not eligible, or forced
    const int S = c->cfg.num_stages;
    StagePlan p = make_plan(c, 0, c->i0, plan_next_buffer(S, 0, c->i0, c->i0, c->i1, c->i2));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    float ms[2] = {0.f, 0.f};
    const long long l0 = c->launches;
    constexpr int kReps = 5;
    for (int path = 0; path < 2; ++path) {
        auto launch = [&]() { return path ? launch_stage_split(c, p, c->stream) : launch_stage(c, p, 0); };
        int rc = launch();   // eager once: function attributes, caches
        if (rc) return rc;
        // timed the way the time loop runs it: as nodes of a CUDA graph (three eager launches per stage would charge the split
        // path ~4 us of launch gaps it does not have inside the captured step)
        cudaGraph_t g = nullptr;
        cudaGraphExec_t ge = nullptr;
        CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        for (int rep = 0; rep < kReps && !rc; ++rep) rc = launch();
        cudaError_t ec = cudaStreamEndCapture(c->stream, &g);
        if (rc || ec != cudaSuccess) { if (g) cudaGraphDestroy(g); return rc ? rc : set_err(PYH_ERR_CUDA, "stage-path tuning: capture failed: %s", cudaGetErrorString(ec)); }
        ec = cudaGraphInstantiate(&ge, g, 0);
        cudaGraphDestroy(g);
        if (ec != cudaSuccess) return set_err(PYH_ERR_CUDA, "stage-path tuning: cudaGraphInstantiate failed: %s", cudaGetErrorString(ec));
        CU(cudaGraphLaunch(ge, c->stream));   // warm-up
        CU(cudaEventRecord(e0, c->stream));
        CU(cudaGraphLaunch(ge, c->stream));
        CU(cudaEventRecord(e1, c->stream));
        CU(cudaEventSynchronize(e1));
        CU(cudaEventElapsedTime(&ms[path], e0, e1));
        cudaGraphExecDestroy(ge);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    c->launches = l0;                                   // tuning launches are not part of any step
    c->tune_ms[0] = ms[0] / kReps; c->tune_ms[1] = ms[1] / kReps;
    c->use_split = ms[1] < 0.97f * ms[0];               // ties go to the fused kernel
    if (!c->use_split && c->aux_hit_ratio > 0.f) {      // the fused kernel it is: it gets the whole L2 back
        cudaCtxResetPersistingL2Cache();
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
        c->aux_hit_ratio = 0.f;
    }
    return 0;
}

// RK partial-sum plan for stage s (explicit_runge_kutta.py:63-75 restated as running sums:
// row s' accumulates U0 + sum_{k<=s} (dt*a[s'][k]) R_k in k order, exactly the reference's order)
StagePlan make_plan(Ctx* c, int s, int cur, int next) {
    return plan_stage(c->tab.a, c->cfg.num_stages, c->po, c->i0, s, cur, next);   // pyh_plan.cuh
}

int do_ghost(Ctx* c, int buf) {
    if (c->blocks.empty()) return 0;
    int m = std::max(c->lay.nx, c->lay.ny);
    dim3 grid(cdiv(m, 128), 4, (unsigned)c->blocks.size());
    k_ghost<<<grid, 128, 0, c->stream>>>(c->d_blks, c->lay, c->po, c->po.H[buf], c->d_ctl);
    CU(cudaGetLastError());
    c->launches++;
    return 0;
}

// buffer roles of stage s: returns the plan, advances c->cur (and the solution role after the last stage)
StagePlan advance_roles(Ctx* c, int s, bool fuse_dt = false) {
    const int S = c->cfg.num_stages;
    int cur = (s == 0) ? c->i0 : c->cur;
    const int next = plan_next_buffer(S, s, cur, c->i0, c->i1, c->i2);
    StagePlan p = make_plan(c, s, cur, next);
    p.fuse_dt = (fuse_dt && s == S - 1) ? 1 : 0;   // the last stage also reduces the CFL minimum of the state it writes
    p.push_ghost = c->push_ok ? 1 : 0;             // ... and every stage refreshes the ghost cells that mirror the edge cells it writes
    c->cur = next;
    if (s == S - 1) {
        if (S == 1) std::swap(c->i0, c->i1);
        c->cur = c->i0;
    }
    return p;
}

int do_stage(Ctx* c, int s) {
    StagePlan p = advance_roles(c, s);
    return c->use_split ? launch_stage_split(c, p, c->stream) : launch_stage(c, p, 0);
}

int set_active(Ctx* c, int active) {
    CU(cudaMemcpyAsync(&c->d_ctl->active, &active, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    return 0;
}

int launch_dt(Ctx* c, int buf, int respect_active = 0) {
    if (c->blocks.empty()) return 0;   // dtmin_bits stays +inf: Solver.get_dt's np.inf for a rank without blocks
    dim3 grid(cdiv(c->lay.nx, 256), cdiv(c->lay.ny, DT_ROWS), (unsigned)c->blocks.size());
    k_dt<<<grid, 256, 0, c->stream>>>(c->d_blks, c->lay, c->po, c->po.H[buf], (int)c->blocks.size(), c->d_ctl, c->C, respect_active);
    CU(cudaGetLastError());
    c->launches++;
    return 0;
}

// Remote ghost strips of buffer `buf` (GhostBlock.send_boundary_data / recieve_boundary_data / apply_recv_buffers_to_state,
// blocks/ghost.py:169-241, and the Waitall of blocks/base.py:454-465): pack the edge strips the neighbour ranks need, ONE
// grouped ncclSend / ncclRecv batch, unpack into the ghost frames -- all in order on the compute stream (capturable).
int exchange_halo(Ctx* c, int buf, cudaStream_t st = nullptr, bool packed = false) {
    Comm& m = c->comm;
    if (!st) st = c->stream;
    if (!m.comm || c->slots.empty()) return 0;
    NcclApi& N = nccl_api();
    const int mx = std::max(c->lay.nx, c->lay.ny);
    dim3 grid(cdiv(mx, 128), (unsigned)c->slots.size());
    if (!packed) {   // after a stage with plan.push_ghost the stage kernel has already written the strips into the send buffer
        k_pack_halo<<<grid, 128, 0, st>>>(c->d_blks, c->lay, c->po.H[buf], c->d_slots, m.d_send);
        CU(cudaGetLastError());
        c->launches++;
    }
    NC(N.GroupStart());
    for (const HaloMsg& r : m.recvs) NC(N.Recv(m.d_recv + r.offset, (size_t)r.len, kNcclFloat64, r.peer, m.comm, st));
    for (const HaloMsg& q : m.sends) NC(N.Send(m.d_send + q.offset, (size_t)q.len, kNcclFloat64, q.peer, m.comm, st));
    NC(N.GroupEnd());
    k_unpack_halo<<<grid, 128, 0, st>>>(c->d_blks, c->lay, c->po.H[buf], c->d_slots, m.d_recv);
    CU(cudaGetLastError());
    c->launches++;
    return 0;
}

// One Runge-Kutta stage and the ghost refresh behind it (ExplicitRungeKutta.integrate's loop body, explicit_runge_kutta.py:63-80:
// residual + update, then Blocks.apply_boundary_condition).  With remote neighbours the stage is split (pyh_plan.cuh: plan_tiles):
// the thin strips that produce the cells other ranks need run first on a side stream, followed there by pack -> grouped
// ncclSend / ncclRecv -> unpack, while the interior launch runs on the compute stream; the two streams join before the local
// ghost copies.  The exchange is thereby off the critical path and a rank may run up to one stage ahead of its neighbours.
int stage_and_refresh(Ctx* c, int s, bool fuse_dt = false) {
    static const bool no_overlap = getenv("PYH_NO_HALO_OVERLAP") != nullptr;   // diagnostics: blocking exchange behind one launch
    StagePlan p = advance_roles(c, s, fuse_dt);
    int rc;
    static const bool force_split = getenv("PYH_FORCE_EDGE_SPLIT") != nullptr;      // diagnostics: the edge / interior split on ONE rank
    if (force_split && !c->s_edge && !c->blocks.empty()) {
        int lo = 0, hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CU(cudaStreamCreateWithPriority(&c->s_edge, cudaStreamNonBlocking, hi));
        CU(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->ev_edge_done, cudaEventDisableTiming));
        c->split_ns = true;
    }
    if (!force_split && (!c->comm.comm || c->slots.empty() || no_overlap || !c->s_edge)) {
        if ((rc = c->use_split ? launch_stage_split(c, p, c->stream) : launch_stage(c, p, 0))) return rc;
        if ((rc = exchange_halo(c, c->cur, nullptr, c->push_ok))) return rc;
        return c->push_ok ? 0 : do_ghost(c, c->cur);
    }
    TileLaunch tl[3];
    static const int edge_rows = getenv("PYH_EDGE_ROWS") ? std::max(1, atoi(getenv("PYH_EDGE_ROWS"))) : kEdgeRows;   // diagnostics
    const int n = plan_tiles(c->lay.nx, c->lay.ny, c->march_nt, c->march_tys, c->split_ns, c->split_ew, tl, edge_rows);
    CU(cudaEventRecord(c->ev_fork, c->stream));
    CU(cudaStreamWaitEvent(c->s_edge, c->ev_fork, 0));
    for (int q = 0; q < n; ++q)
        if (tl[q].edge && (rc = launch_stage(c, p, 0, tl[q], c->s_edge))) return rc;
    if ((rc = exchange_halo(c, c->cur, c->s_edge, c->push_ok))) return rc;
    CU(cudaEventRecord(c->ev_edge_done, c->s_edge));
    for (int q = 0; q < n; ++q)
        if (!tl[q].edge && (rc = launch_stage(c, p, 0, tl[q], c->stream))) return rc;
    CU(cudaStreamWaitEvent(c->stream, c->ev_edge_done, 0));
    return c->push_ok ? 0 : do_ghost(c, c->cur);   // local ghost cells: written by the stage kernel itself, else by k_ghost
}

// Global CFL minimum and realizability flag (Solver.get_dt gathers and broadcasts the minimum, solvers/base.py:128-131;
// min is exact, so the result does not depend on the rank count): one in-place ncclAllReduce(min) over the two adjacent
// 64-bit words {dtmin_bits, allok} of the device control block.
int reduce_dt(Ctx* c) {
    Comm& m = c->comm;
    if (!m.comm) return 0;
    static_assert(offsetof(Control, allok) == offsetof(Control, dtmin_bits) + sizeof(unsigned long long), "dtmin_bits and allok must be adjacent");
    NC(nccl_api().AllReduce(&c->d_ctl->dtmin_bits, &c->d_ctl->dtmin_bits, 2, kNcclUint64, kNcclMin, m.comm, c->stream));
    return 0;
}

Ctx* as_ctx(void* p) { return reinterpret_cast<Ctx*>(p); }

}  // namespace

extern "C" {

const char* pyh_last_error(void) { return g_err; }
int pyh_abi_version(void) { return PYH_ABI_VERSION; }

int pyh_create(const pyh_config* cfg, void** out) {
    if (!cfg || !out) return set_err(PYH_ERR_INVALID, "null argument");
    if (cfg->abi_version != PYH_ABI_VERSION) return set_err(PYH_ERR_INVALID, "ABI version mismatch: got %d, library is %d", cfg->abi_version, PYH_ABI_VERSION);
    if (cfg->nx < 1 || cfg->ny < 1) return set_err(PYH_ERR_INVALID, "nx, ny must be >= 1");
    if (cfg->flux < 0 || cfg->flux > 2) return set_err(PYH_ERR_INVALID, "unknown flux function %d", cfg->flux);
    if (cfg->limiter < 0 || cfg->limiter > 3) return set_err(PYH_ERR_INVALID, "unknown slope limiter %d", cfg->limiter);
    if (cfg->recon < 0 || cfg->recon > 1) return set_err(PYH_ERR_INVALID, "unknown reconstruction type %d", cfg->recon);
    if (cfg->num_quadrature_points < 1 || cfg->num_quadrature_points > 3) return set_err(PYH_ERR_INVALID, "fvm_num_quadrature_points must be 1, 2 or 3 (got %d)", cfg->num_quadrature_points);
    if (cfg->num_stages < 1 || cfg->num_stages > PYH_MAX_STAGES) return set_err(PYH_ERR_INVALID, "num_stages must be in 1..%d", PYH_MAX_STAGES);
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) return set_err(PYH_ERR_INVALID, "device %d out of range (%d visible)", cfg->device, ndev);
    CU(cudaSetDevice(cfg->device));
    Ctx* c = new Ctx();
    c->cfg = *cfg;
    c->lay.nx = cfg->nx; c->lay.ny = cfg->ny;
    c->lay.pitch = ((cfg->nx + PADL + 1 + 3) / 4) * 4;
    {
        unsigned long long pl = (unsigned long long)(cfg->ny + 2) * (unsigned long long)c->lay.pitch;
        if (pl * 64ull >= (1ull << 32)) { delete c; return set_err(PYH_ERR_INVALID, "block of %d x %d cells is too large for 32-bit slab offsets", cfg->nx, cfg->ny); }
        c->lay.plane = (unsigned)pl;
    }
    c->C.g = cfg->gamma;
    c->C.gm1 = cfg->gamma - 1.0;
    c->C.k = 1.0 / (cfg->gamma - 1.0);
    c->C.gm = cfg->gamma / (cfg->gamma - 1.0);
    {   // mesh/quadratures.py:33-37, same expressions in the same (dict) order
        const int nq = cfg->num_quadrature_points;
        for (int q = 0; q < 3; ++q) { c->C.qw[q] = 0.0; c->C.qp[q] = 0.0; }
        if (nq == 1) { c->C.qp[0] = 0.0; c->C.qw[0] = 2.0; }
        else if (nq == 2) { c->C.qp[0] = -1.0 / std::sqrt(3.0); c->C.qp[1] = 1.0 / std::sqrt(3.0); c->C.qw[0] = c->C.qw[1] = 1.0; }
        else { c->C.qp[0] = -std::sqrt(3.0 / 5.0); c->C.qp[1] = 0.0; c->C.qp[2] = std::sqrt(3.0 / 5.0);
               c->C.qw[0] = 5.0 / 9.0; c->C.qw[1] = 8.0 / 9.0; c->C.qw[2] = 5.0 / 9.0; }
    }
    memset(&c->tab, 0, sizeof(c->tab));
    c->tab.nstages = cfg->num_stages;
    for (int s = 0; s < cfg->num_stages; ++s)
        for (int k = 0; k <= s; ++k) c->tab.a[s * PYH_MAX_STAGES + k] = cfg->tableau[s * PYH_MAX_STAGES + k];
    plan_need_acc(c->tab.a, cfg->num_stages, c->need_acc);                                               // pyh_plan.cuh
    c->po = plan_offsets(c->lay.plane, cfg->num_stages, cfg->num_quadrature_points, c->need_acc);       // slab layout
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CU(cudaMalloc(&c->d_ctl, sizeof(Control)));
    Control h;
    memset(&h, 0, sizeof(h));
    h.dtmin_bits = DKEY_INF;
    h.allok = 1ull;
    h.active = 1;
    CU(cudaMemcpy(c->d_ctl, &h, sizeof(h), cudaMemcpyHostToDevice));
    CU(cudaMalloc(&c->d_tmp, 64 * sizeof(double)));
    *out = c;
    return 0;
}

int pyh_add_block(void* ctx, const pyh_block_desc* b) {
    Ctx* c = as_ctx(ctx);
    if (!c || !b) return set_err(PYH_ERR_INVALID, "null argument");
    if (c->finalized) return set_err(PYH_ERR_STATE, "pyh_add_block after pyh_finalize");
    if (c->gid2idx.count(b->gid)) return set_err(PYH_ERR_INVALID, "block %d added twice", b->gid);
    if (!b->nodes_x || !b->nodes_y || !b->area || !b->cos_v || !b->sin_v || !b->cos_h || !b->sin_h)
        return set_err(PYH_ERR_INVALID, "block %d: null geometry pointer", b->gid);
    for (int s = 0; s < 4; ++s) {
        if (b->bc[s] < 0 || b->bc[s] > PYH_BC_PRIMITIVE_DIRICHLET) return set_err(PYH_ERR_INVALID, "Boundary Condition type %d has not been specialized.", b->bc[s]);
        if (b->bc[s] == PYH_BC_PRIMITIVE_DIRICHLET && !b->dirichlet_prim[s]) return set_err(PYH_ERR_INVALID, "block %d side %d: Dirichlet BC without inlet state", b->gid, s);
    }
    CU(cudaSetDevice(c->cfg.device));
    const Layout& L = c->lay;
    const int nx = L.nx, ny = L.ny;
    HostBlock hb;
    hb.d = *b;
    memset(&hb.dev, 0, sizeof(hb.dev));
    BlkDev& D = hb.dev;
    int rc;
    double* slab = nullptr;
    if ((rc = dalloc(hb, &slab, (long long)c->po.nplanes * L.plane, true))) return rc;
    D.base = slab;
    const PlaneOffsets& po = c->po;
    double *A = slab + po.A, *dxy = slab + po.dxy, *Lv = slab + po.Lv, *cv = slab + po.cv, *sv = slab + po.sv;
    double *Lh = slab + po.Lh, *ch = slab + po.ch, *sh = slab + po.sh, *cdx = slab + po.cdx, *cdy = slab + po.cdy;
    // stage host arrays through the scratch buffer
    size_t nn = (size_t)(ny + 1) * (nx + 1);
    if ((rc = ensure_scratch(c, 2 * nn * sizeof(double)))) return rc;
    auto put = [&](const double* host, double* plane, int rows, int cols) -> int {
        size_t n = (size_t)rows * cols;
        CU(cudaMemcpyAsync(c->d_scratch, host, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        k_dense_to_plane<<<cdiv(n, 256), 256, 0, c->stream>>>(L, c->d_scratch, plane, rows, cols);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(c->stream));
        return 0;
    };
    if ((rc = put(b->area, A, ny, nx))) return rc;
    if ((rc = put(b->cos_v, cv, ny, nx + 1))) return rc;
    if ((rc = put(b->sin_v, sv, ny, nx + 1))) return rc;
    if ((rc = put(b->cos_h, ch, ny + 1, nx))) return rc;
    if ((rc = put(b->sin_h, sh, ny + 1, nx))) return rc;
    CU(cudaMemcpyAsync(c->d_scratch, b->nodes_x, nn * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_scratch + nn, b->nodes_y, nn * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    k_geometry<<<cdiv(nn, 256), 256, 0, c->stream>>>(L, c->d_scratch, c->d_scratch + nn, dxy, Lv, Lh, cdx, cdy, slab + po.xc, slab + po.yc, c->cfg.num_quadrature_points, c->C);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    D.dbg = nullptr;
    D.dbgG = nullptr;
    for (int s = 0; s < 4; ++s) {
        D.bc[s] = b->bc[s];
        D.nbr[s] = -1;
        D.remote_slot[s] = -1;
        D.dir_recon[s] = nullptr;
        D.dir_cons[s] = nullptr;
        if (b->bc[s] == PYH_BC_PRIMITIVE_DIRICHLET) {
            int len = (s == PYH_EAST || s == PYH_WEST) ? ny : nx;
            double *prim, *recon, *cons;
            if ((rc = dalloc(hb, &prim, 4 * len, false))) return rc;
            if ((rc = dalloc(hb, &recon, 4 * len, false))) return rc;
            if ((rc = dalloc(hb, &cons, 4 * len, false))) return rc;
            CU(cudaMemcpy(prim, b->dirichlet_prim[s], 4 * (size_t)len * sizeof(double), cudaMemcpyHostToDevice));
            k_dirichlet<<<cdiv(len, 128), 128, 0, c->stream>>>(prim, recon, cons, len, c->cfg.recon, c->C);
            CU(cudaGetLastError());
            CU(cudaStreamSynchronize(c->stream));
            D.dir_recon[s] = recon;
            D.dir_cons[s] = cons;
        }
    }
    D.cart = b->is_cartesian ? 1 : 0;
    {   // bit 1: every vertical face of the block has theta == 0 exactly (cos == 1, sin == 0), e.g. the rectangular blocks of
        // RectagularMeshGenerator, whose corner coordinates fail the reference's exact is_cartesian test;
        // the rotation into / out of such a face frame is the identity by value (PYH_SKIP_UNIT_ROT, pyh_stage_march.cuh)
        bool unit = true;
        const size_t nv = (size_t)ny * (nx + 1);
        for (size_t i = 0; i < nv && unit; ++i) unit = (b->cos_v[i] == 1.0) && (b->sin_v[i] == 0.0);
        if (unit) D.cart |= 2;
    }
    D.gid = b->gid;
    hb.d.nodes_x = hb.d.nodes_y = hb.d.area = hb.d.cos_v = hb.d.sin_v = hb.d.cos_h = hb.d.sin_h = nullptr;
    for (int s = 0; s < 4; ++s) hb.d.dirichlet_prim[s] = nullptr;
    c->gid2idx[b->gid] = (int)c->blocks.size();
    c->blocks.push_back(hb);
    return 0;
}

int pyh_finalize(void* ctx) {
    Ctx* c = as_ctx(ctx);
    if (!c) return set_err(PYH_ERR_INVALID, "null context");
    if (c->finalized) return set_err(PYH_ERR_STATE, "pyh_finalize called twice");
    CU(cudaSetDevice(c->cfg.device));
    // halo slots in (gid, side) ascending order
    std::vector<std::pair<int, int>> order;
    for (auto& kv : c->gid2idx) order.push_back({kv.first, kv.second});
    std::sort(order.begin(), order.end());
    c->slots.clear();
    c->halo_doubles = 0;
    for (auto& pr : order) {
        HostBlock& hb = c->blocks[pr.second];
        for (int s = 0; s < 4; ++s) {
            int ng = hb.d.neighbor[s];
            hb.dev.nbr[s] = -1;
            hb.dev.remote_slot[s] = -1;
            if (ng < 0) continue;
            if (hb.d.neighbor_is_local[s]) {
                auto it = c->gid2idx.find(ng);
                if (it == c->gid2idx.end()) return set_err(PYH_ERR_INVALID, "block %d: local neighbour %d was not added", hb.d.gid, ng);
                hb.dev.nbr[s] = it->second;
            } else if (hb.d.bc[s] == PYH_BC_NONE) {
                int len = (s == PYH_EAST || s == PYH_WEST) ? c->lay.ny : c->lay.nx;
                HaloSlot hs;
                hs.blk = pr.second; hs.side = s; hs.offset = c->halo_doubles;
                hb.dev.remote_slot[s] = (int)c->slots.size();
                c->slots.push_back(hs);
                c->slot_nbr_gid.push_back(ng);
                c->halo_doubles += 4LL * len;
            }
        }
    }
    // Push-model ghost refresh (pyh_stage_march.cuh: push_ghost_cells) needs a symmetric topology: where block b sees block n
    // across side s without a boundary condition, n must see b across the opposite side without one.  Every mesh generator
    // of the reference builds such dictionaries; anything else keeps the per-stage k_ghost / k_pack_halo kernels.
    {
        static const int opposite[4] = {PYH_WEST, PYH_EAST, PYH_SOUTH, PYH_NORTH};
        bool sym = getenv("PYH_NO_PUSH_GHOST") == nullptr;
        for (auto& hb : c->blocks)
            for (int s = 0; s < 4 && sym; ++s) {
                if (hb.dev.nbr[s] < 0 || hb.d.bc[s] != PYH_BC_NONE) continue;
                const HostBlock& nb = c->blocks[hb.dev.nbr[s]];
                sym = nb.d.bc[opposite[s]] == PYH_BC_NONE && nb.d.neighbor[opposite[s]] == hb.d.gid && nb.d.neighbor_is_local[opposite[s]];
            }
        c->push_ok = sym;
    }
    if (c->halo_doubles > 0) {   // strips for / from neighbours on other ranks (also the target of the stage kernel's pushes)
        CU(cudaMalloc(&c->comm.d_send, (size_t)c->halo_doubles * sizeof(double)));
        CU(cudaMalloc(&c->comm.d_recv, (size_t)c->halo_doubles * sizeof(double)));
        CU(cudaMemset(c->comm.d_send, 0, (size_t)c->halo_doubles * sizeof(double)));
        c->comm.doubles = c->halo_doubles;
        for (const HaloSlot& hs : c->slots) c->blocks[hs.blk].dev.send[hs.side] = c->comm.d_send + hs.offset;
    }
    c->split_ready = split_eligible(c);
    c->use_split = c->split_ready && getenv("PYH_SPLIT") != nullptr;   // forced; otherwise decided by measurement at the first pyh_run
    if (c->split_ready) {
        // limited face states + face fluxes between the three kernels of a stage: written once, read once, then dead.  One
        // allocation for all blocks, marked PERSISTING in the L2 for the split kernels (launch attribute, pyh_split.cu): what a
        // stage streams through (state, geometry) can then not evict what the next kernel is about to read ...
        const size_t nb = c->blocks.size();
        const size_t fs_per = (size_t)kSplitStatePlanes * c->lay.plane, fx_per = (size_t)kSplitFluxPlanes * c->lay.plane;   // doubles
        const size_t fs_bytes = fs_per * nb * sizeof(double), fx_bytes = fx_per * nb * sizeof(double);
        c->aux_bytes = fs_bytes + fx_bytes;
        cudaError_t e = cudaMalloc(&c->d_aux, c->aux_bytes);
        if (e != cudaSuccess) return set_err(PYH_ERR_NOMEM, "cudaMalloc of %zu bytes of stage scratch failed: %s", c->aux_bytes, cudaGetErrorString(e));
        CU(cudaMemset(c->d_aux, 0, c->aux_bytes));
        double* const fx0 = c->d_aux + fs_per * nb;   // layout: [face states of every block][face fluxes of every block]
        for (size_t b = 0; b < nb; ++b) { c->blocks[b].dev.aux = c->d_aux + b * fs_per; c->blocks[b].dev.aux_fx = fx0 + b * fx_per; }
        // ... but only if ALL of it fits the persisting carve-out the device grants (B200: 79 MB of the 126 MB L2).  Measured
        // (profiles/r02y_l2_persist_ab.txt): explosion_multi, 36 MB of scratch, 0.158 -> 0.150 ms/step; a PARTIAL window (DMR: 193 MB,
        // hitRatio 0.43) thrashes, 0.44 -> 0.79; a window over the flux planes alone (64 MB for DMR) changes nothing (-1.8 % / +2.5 %).
        if (!getenv("PYH_NO_L2_PERSIST")) {
            int max_persist = 0, max_window = 0;
            cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, c->cfg.device);
            cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, c->cfg.device);
            const size_t cap = (size_t)std::max(std::min(max_persist, max_window), 0);
            const void* base = nullptr;
            size_t bytes = 0;
            if (c->aux_bytes <= cap) { base = c->d_aux; bytes = c->aux_bytes; }
            if (bytes > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, bytes) == cudaSuccess) {
                size_t got = 0;
                cudaDeviceGetLimit(&got, cudaLimitPersistingL2CacheSize);
                if (got >= bytes) { c->aux_hit_ratio = 1.0f; c->win_base = base; c->win_bytes = bytes; }
                else cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
            } else cudaGetLastError();
        }
    }
    std::vector<BlkDev> tmp;
    for (auto& hb : c->blocks) tmp.push_back(hb.dev);
    CU(cudaMalloc(&c->d_blks, std::max<size_t>(tmp.size(), 1) * sizeof(BlkDev)));
    if (!tmp.empty()) CU(cudaMemcpy(c->d_blks, tmp.data(), tmp.size() * sizeof(BlkDev), cudaMemcpyHostToDevice));
    if (!c->slots.empty()) {
        CU(cudaMalloc(&c->d_slots, c->slots.size() * sizeof(HaloSlot)));
        CU(cudaMemcpy(c->d_slots, c->slots.data(), c->slots.size() * sizeof(HaloSlot), cudaMemcpyHostToDevice));
    }
    choose_march_shape(c);
    c->finalized = true;
    return 0;
}

int pyh_destroy(void* ctx) {
    Ctx* c = as_ctx(ctx);
    if (!c) return 0;
    cudaSetDevice(c->cfg.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (auto& hb : c->blocks) {
        for (void* p : hb.allocs) cudaFree(p);
        if (hb.dev.dbg) cudaFree(hb.dev.dbg);
        if (hb.dev.dbgG) cudaFree(hb.dev.dbgG);
    }
    if (c->d_blks) cudaFree(c->d_blks);
    if (c->d_ctl) cudaFree(c->d_ctl);
    if (c->d_slots) cudaFree(c->d_slots);
    if (c->d_scratch) cudaFree(c->d_scratch);
    if (c->d_dts) cudaFree(c->d_dts);
    if (c->d_tmp) cudaFree(c->d_tmp);
    if (c->aux_hit_ratio > 0.f) {   // give the persisting carve-out back: later contexts of this process want the whole L2
        cudaCtxResetPersistingL2Cache();
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
    }
    if (c->d_aux) cudaFree(c->d_aux);
    if (c->s_in) { cudaStreamSynchronize(c->s_in); cudaStreamDestroy(c->s_in); }
    if (c->s_out) { cudaStreamSynchronize(c->s_out); cudaStreamDestroy(c->s_out); }
    if (c->ev_in_done) cudaEventDestroy(c->ev_in_done);
    if (c->ev_in_consumed) cudaEventDestroy(c->ev_in_consumed);
    if (c->ev_out_ready) cudaEventDestroy(c->ev_out_ready);
    for (cudaEvent_t e : c->ev_out_done) if (e) cudaEventDestroy(e);
    if (c->run_graph) cudaGraphExecDestroy(c->run_graph);
    if (c->comm.comm) nccl_api().CommDestroy(c->comm.comm);
    if (c->s_edge) { cudaStreamSynchronize(c->s_edge); cudaStreamDestroy(c->s_edge); }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_edge_done) cudaEventDestroy(c->ev_edge_done);
    if (c->comm.d_send) cudaFree(c->comm.d_send);
    if (c->comm.d_recv) cudaFree(c->comm.d_recv);
    if (c->d_stage_in) cudaFree(c->d_stage_in);
    if (c->d_stage_out) cudaFree(c->d_stage_out);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

#define GET_BLOCK(c, gid, hb)                                                              \
    if (!(c)) return set_err(PYH_ERR_INVALID, "null context");                             \
    auto it_ = (c)->gid2idx.find(gid);                                                      \
    if (it_ == (c)->gid2idx.end()) return set_err(PYH_ERR_INVALID, "unknown block %d", gid); \
    HostBlock& hb = (c)->blocks[it_->second];                                               \
    CU(cudaSetDevice((c)->cfg.device));

int pyh_upload_state(void* ctx, int gid, const double* aos) {
    Ctx* c = as_ctx(ctx);
    GET_BLOCK(c, gid, hb);
    if (!aos) return set_err(PYH_ERR_INVALID, "null state pointer");
    size_t n = (size_t)c->lay.nx * c->lay.ny;
    int rc = ensure_scratch(c, 4 * n * sizeof(double));
    if (rc) return rc;
    CU(cudaMemcpyAsync(c->d_scratch, aos, 4 * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    k_aos_to_soa<<<cdiv(n, 256), 256, 0, c->stream>>>(c->lay, c->d_scratch, hb.dev.base + c->po.H[c->i0]);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    c->launches++;
    return 0;
}

int pyh_download_state(void* ctx, int gid, double* aos) {
    Ctx* c = as_ctx(ctx);
    GET_BLOCK(c, gid, hb);
    if (!aos) return set_err(PYH_ERR_INVALID, "null state pointer");
    size_t n = (size_t)c->lay.nx * c->lay.ny;
    int rc = ensure_scratch(c, 4 * n * sizeof(double));
    if (rc) return rc;
    k_soa_to_aos<<<cdiv(n, 256), 256, 0, c->stream>>>(c->lay, hb.dev.base + c->po.H[c->i0], c->d_scratch, 4);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(aos, c->d_scratch, 4 * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->launches++;
    return 0;
}

int pyh_fill_uniform(void* ctx, int gid, const double* state) {
    Ctx* c = as_ctx(ctx);
    GET_BLOCK(c, gid, hb);
    if (!state) return set_err(PYH_ERR_INVALID, "null state pointer");
    size_t n = (size_t)c->lay.nx * c->lay.ny;
    k_fill_uniform<<<cdiv(n, 256), 256, 0, c->stream>>>(c->lay, hb.dev.base + c->po.H[c->i0], state[0], state[1], state[2], state[3]);
    CU(cudaGetLastError());
    c->launches++;
    return 0;
}

int pyh_fill_box(void* ctx, int gid, double x0, double x1, double y0, double y1, const double* inside, const double* outside) {
    Ctx* c = as_ctx(ctx);
    GET_BLOCK(c, gid, hb);
    if (!inside) return set_err(PYH_ERR_INVALID, "null state pointer");
    size_t n = (size_t)c->lay.nx * c->lay.ny;
    const double z[4] = {0.0, 0.0, 0.0, 0.0};
    const double* o = outside ? outside : z;
    k_fill_box<<<cdiv(n, 256), 256, 0, c->stream>>>(c->lay, hb.dev.base + c->po.H[c->i0], hb.dev.base + c->po.xc, hb.dev.base + c->po.yc,
                                                    x0, x1, y0, y1, inside[0], inside[1], inside[2], inside[3], outside ? 1 : 0, o[0], o[1], o[2], o[3]);
    CU(cudaGetLastError());
    c->launches++;
    return 0;
}

int pyh_upload_state_async(void* ctx, int gid, const double* aos) {
    Ctx* c = as_ctx(ctx);
    GET_BLOCK(c, gid, hb);
    (void)hb;
    if (!c->finalized) return set_err(PYH_ERR_STATE, "context not finalized");
    if (!aos) return set_err(PYH_ERR_INVALID, "null state pointer");
    int rc = ensure_streaming(c);
    if (rc) return rc;
    const int idx = it_->second;
    const size_t per = 4 * (size_t)c->lay.nx * c->lay.ny;
    // the staging area may still be read by the conversion of the previous batch
    CU(cudaStreamWaitEvent(c->s_in, c->ev_in_consumed, 0));
    CU(cudaMemcpyAsync(c->d_stage_in + idx * per, aos, per * sizeof(double), cudaMemcpyHostToDevice, c->s_in));
    c->staged[idx] = 1;
    return 0;
}

int pyh_commit_uploads(void* ctx) {
    Ctx* c = as_ctx(ctx);
    if (!c || !c->finalized) return set_err(PYH_ERR_STATE, "context not finalized");
    CU(cudaSetDevice(c->cfg.device));
    int rc = ensure_streaming(c);
    if (rc) return rc;
    const size_t n = (size_t)c->lay.nx * c->lay.ny;
    CU(cudaEventRecord(c->ev_in_done, c->s_in));
    CU(cudaStreamWaitEvent(c->stream, c->ev_in_done, 0));
    for (size_t b = 0; b < c->blocks.size(); ++b) {
        if (!c->staged[b]) continue;
        k_aos_to_soa<<<cdiv(n, 256), 256, 0, c->stream>>>(c->lay, c->d_stage_in + b * 4 * n, c->blocks[b].dev.base + c->po.H[c->i0]);
        CU(cudaGetLastError());
        c->launches++;
        c->staged[b] = 0;
    }
    CU(cudaEventRecord(c->ev_in_consumed, c->stream));
    return 0;
}

int pyh_download_state_async(void* ctx, int gid, double* aos) {
    Ctx* c = as_ctx(ctx);
    GET_BLOCK(c, gid, hb);
    if (!c->finalized) return set_err(PYH_ERR_STATE, "context not finalized");
    if (!aos) return set_err(PYH_ERR_INVALID, "null state pointer");
    int rc = ensure_streaming(c);
    if (rc) return rc;
    const int idx = it_->second;
    const size_t n = (size_t)c->lay.nx * c->lay.ny;
    double* stage = c->d_stage_out + idx * 4 * n;
    CU(cudaStreamWaitEvent(c->stream, c->ev_out_done[idx], 0));   // previous copy out of this block's staging area
    k_soa_to_aos<<<cdiv(n, 256), 256, 0, c->stream>>>(c->lay, hb.dev.base + c->po.H[c->i0], stage, 4);
    CU(cudaGetLastError());
    c->launches++;
    CU(cudaEventRecord(c->ev_out_ready, c->stream));
    CU(cudaStreamWaitEvent(c->s_out, c->ev_out_ready, 0));
    CU(cudaMemcpyAsync(aos, stage, 4 * n * sizeof(double), cudaMemcpyDeviceToHost, c->s_out));
    CU(cudaEventRecord(c->ev_out_done[idx], c->s_out));
    return 0;
}

int pyh_transfers_sync(void* ctx) {
    Ctx* c = as_ctx(ctx);
    if (!c) return set_err(PYH_ERR_INVALID, "null context");
    CU(cudaSetDevice(c->cfg.device));
    if (c->s_in) CU(cudaStreamSynchronize(c->s_in));
    CU(cudaStreamSynchronize(c->stream));
    if (c->s_out) CU(cudaStreamSynchronize(c->s_out));
    return 0;
}

int pyh_downloads_sync(void* ctx) {
    Ctx* c = as_ctx(ctx);
    if (!c) return set_err(PYH_ERR_INVALID, "null context");
    CU(cudaSetDevice(c->cfg.device));
    if (c->s_out) CU(cudaStreamSynchronize(c->s_out));
    return 0;
}

int pyh_host_alloc(size_t bytes, void** out) {
    if (!out) return set_err(PYH_ERR_INVALID, "null pointer");
    cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) return set_err(PYH_ERR_NOMEM, "cudaHostAlloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    return 0;
}

int pyh_host_free(void* p) {
    if (p) CU(cudaFreeHost(p));
    return 0;
}

int pyh_download_ghost(void* ctx, int gid, int side, double* out) {
    Ctx* c = as_ctx(ctx);
    GET_BLOCK(c, gid, hb);
    if (side < 0 || side > 3 || !out) return set_err(PYH_ERR_INVALID, "bad side / null pointer");
    int len = (side == PYH_EAST || side == PYH_WEST) ? c->lay.ny : c->lay.nx;
    int rc = ensure_scratch(c, 4 * (size_t)len * sizeof(double));
    if (rc) return rc;
    k_ghost_strip_fetch<<<cdiv(len, 128), 128, 0, c->stream>>>(c->lay, hb.dev.base + c->po.H[c->i0], side, c->d_scratch);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, c->d_scratch, 4 * (size_t)len * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int pyh_apply_bc(void* ctx) {
    Ctx* c = as_ctx(ctx);
    if (!c || !c->finalized) return set_err(PYH_ERR_STATE, "context not finalized");
    CU(cudaSetDevice(c->cfg.device));
    int rc = set_active(c, 1);
    if (rc) return rc;
    if ((rc = exchange_halo(c, c->cur))) return rc;   // no-op without pyh_comm_init (host-driven transport: pyh_pack_halo & co)
    return do_ghost(c, c->cur);
}

int pyh_halo_count(void* ctx, int64_t* n_slots, int64_t* n_doubles) {
    Ctx* c = as_ctx(ctx);
    if (!c || !c->finalized) return set_err(PYH_ERR_STATE, "context not finalized");
    if (n_slots) *n_slots = (int64_t)c->slots.size();
    if (n_doubles) *n_doubles = c->halo_doubles;
    return 0;
}

int pyh_halo_slot(void* ctx, int64_t slot, int32_t* gid, int32_t* side, int32_t* nbr_gid, int64_t* offset, int64_t* len) {
    Ctx* c = as_ctx(ctx);
    if (!c || !c->finalized) return set_err(PYH_ERR_STATE, "context not finalized");
    if (slot < 0 || slot >= (int64_t)c->slots.size()) return set_err(PYH_ERR_INVALID, "slot out of range");
    const HaloSlot& s = c->slots[slot];
    if (gid) *gid = c->blocks[s.blk].d.gid;
    if (side) *side = s.side;
    if (nbr_gid) *nbr_gid = c->slot_nbr_gid[slot];
    if (offset) *offset = s.offset;
    if (len) *len = 4LL * ((s.side == PYH_EAST || s.side == PYH_WEST) ? c->lay.ny : c->lay.nx);
    return 0;
}

int pyh_pack_halo(void* ctx, double* dev_send) {
    Ctx* c = as_ctx(ctx);
    if (!c || !c->finalized) return set_err(PYH_ERR_STATE, "context not finalized");
    if (c->slots.empty()) return 0;
    CU(cudaSetDevice(c->cfg.device));
    int m = std::max(c->lay.nx, c->lay.ny);
    dim3 grid(cdiv(m, 128), (unsigned)c->slots.size());
    k_pack_halo<<<grid, 128, 0, c->stream>>>(c->d_blks, c->lay, c->po.H[c->cur], c->d_slots, dev_send);
    CU(cudaGetLastError());
    c->launches++;
    return 0;
}

int pyh_unpack_halo(void* ctx, const double* dev_recv) {
    Ctx* c = as_ctx(ctx);
    if (!c || !c->finalized) return set_err(PYH_ERR_STATE, "context not finalized");
    if (c->slots.empty()) return 0;
    CU(cudaSetDevice(c->cfg.device));
    int m = std::max(c->lay.nx, c->lay.ny);
    dim3 grid(cdiv(m, 128), (unsigned)c->slots.size());
    k_unpack_halo<<<grid, 128, 0, c->stream>>>(c->d_blks, c->lay, c->po.H[c->cur], c->d_slots, dev_recv);
    CU(cudaGetLastError());
    c->launches++;
    return 0;
}

int pyh_local_dt(void* ctx, double* dev_dt_out) {
    Ctx* c = as_ctx(ctx);
    if (!c || !c->finalized) return set_err(PYH_ERR_STATE, "context not finalized");
    CU(cudaSetDevice(c->cfg.device));
    int rc = launch_dt(c, c->i0);
    if (rc) return rc;
    if ((rc = reduce_dt(c))) return rc;   // with pyh_comm_init: the GLOBAL minimum
    k_dt_finalize<<<1, 1, 0, c->stream>>>(c->d_ctl, c->cfg.cfl, c->tab, 1, dev_dt_out ? dev_dt_out : c->d_tmp);   // NULL: the context's own scratch
    CU(cudaGetLastError());
    c->launches++;
    return 0;
}

int pyh_get_dt(void* ctx, double t, double t_final, double* dt_out) {
    Ctx* c = as_ctx(ctx);
    if (!c || !dt_out) return set_err(PYH_ERR_INVALID, "null pointer");
    int rc = pyh_local_dt(ctx, c->d_tmp);
    if (rc) return rc;
    double dt;
    CU(cudaMemcpyAsync(&dt, c->d_tmp, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *dt_out = (t_final - t < dt) ? (t_final - t) : dt;   // solvers/base.py:132-136
    return 0;
}

int pyh_step_begin(void* ctx, double dt) {
    Ctx* c = as_ctx(ctx);
    if (!c || !c->finalized) return set_err(PYH_ERR_STATE, "context not finalized");
    CU(cudaSetDevice(c->cfg.device));
    k_set_dt<<<1, 1, 0, c->stream>>>(c->d_ctl, dt, nullptr, c->tab);
    CU(cudaGetLastError());
    c->launches++;
    c->stage_next = 0;
    return 0;
}

int pyh_step_begin_dev(void* ctx, const double* dev_dt) {
    Ctx* c = as_ctx(ctx);
    if (!c || !c->finalized) return set_err(PYH_ERR_STATE, "context not finalized");
    if (!dev_dt) dev_dt = c->d_tmp;   // NULL: the value the last pyh_local_dt(ctx, NULL) left in the context's scratch
    CU(cudaSetDevice(c->cfg.device));
    k_set_dt<<<1, 1, 0, c->stream>>>(c->d_ctl, 0.0, dev_dt, c->tab);
    CU(cudaGetLastError());
    c->launches++;
    c->stage_next = 0;
    return 0;
}

int pyh_stage(void* ctx, int stage) {
    Ctx* c = as_ctx(ctx);
    if (!c || !c->finalized) return set_err(PYH_ERR_STATE, "context not finalized");
    if (stage != c->stage_next || stage >= c->cfg.num_stages) return set_err(PYH_ERR_STATE, "stage %d out of order (expected %d)", stage, c->stage_next);
    CU(cudaSetDevice(c->cfg.device));
    int rc = do_stage(c, stage);
    if (rc) return rc;
    c->stage_next = stage + 1;
    return 0;
}

int pyh_step(void* ctx, double dt) {
    Ctx* c = as_ctx(ctx);
    if (c && !c->slots.empty() && !c->comm.comm)
        return set_err(PYH_ERR_STATE, "pyh_step on a context with remote neighbours and no pyh_comm_init; drive the stages from the host");
    int rc = pyh_step_begin(ctx, dt);
    if (rc) return rc;
    for (int s = 0; s < c->cfg.num_stages; ++s)
        if ((rc = stage_and_refresh(c, s))) return rc;
    c->stage_next = c->cfg.num_stages;
    return 0;
}

int pyh_run(void* ctx, double* t_inout, double t_final, int64_t max_steps, int32_t poll_every,
            int64_t* steps_done, int32_t* unrealizable, double* dts_out, int64_t dts_cap) {
    Ctx* c = as_ctx(ctx);
    if (!c || !c->finalized) return set_err(PYH_ERR_STATE, "context not finalized");
    if (!t_inout) return set_err(PYH_ERR_INVALID, "null pointer");
    if (!c->slots.empty() && !c->comm.comm)
        return set_err(PYH_ERR_STATE, "pyh_run on a context with remote neighbours needs pyh_comm_init (or drive the stages from the host)");
    CU(cudaSetDevice(c->cfg.device));
    if (poll_every < 1) poll_every = 64;
    if (max_steps < 0) max_steps = (int64_t)1 << 62;
    if (dts_out && dts_cap > 0) {
        if (c->dts_cap < dts_cap) {
            if (c->d_dts) cudaFree(c->d_dts);
            c->d_dts = nullptr;
            CU(cudaMalloc(&c->d_dts, (size_t)dts_cap * sizeof(double)));
            c->dts_cap = dts_cap;
        }
    }
    Control h;
    memset(&h, 0, sizeof(h));
    h.t = *t_inout; h.t_final = t_final; h.dtmin_bits = DKEY_INF; h.allok = 1ull; h.active = 1; h.nsteps = 0; h.bad = 0;
    double* ddts = (dts_out && dts_cap > 0) ? c->d_dts : nullptr;
    h.dts = ddts; h.dts_cap = ddts ? dts_cap : 0;
    CU(cudaMemcpyAsync(c->d_ctl, &h, sizeof(h), cudaMemcpyHostToDevice, c->stream));
    int rc;
    if ((rc = tune_stage_path(c))) return rc;   // first call only: fused kernel or the three-kernel stage, whichever is faster here
    // one time step: [all-reduce of] the CFL minimum, dt, stages + ghost refresh, t += dt; every kernel early-exits once
    // t >= t_final.  The CFL minimum and the realizability flag of a step's FINAL state are reduced inside its last stage
    // (plan.fuse_dt: the state is still in registers there), so only the first step of a call needs the k_dt pass below.
    static const bool no_fuse = getenv("PYH_NO_FUSED_DT") != nullptr;   // diagnostics: k_dt as a kernel of its own every step
    auto enqueue_step = [&]() -> int {
        int r;
        if (no_fuse && (r = launch_dt(c, c->i0, 1))) return r;
        if ((r = reduce_dt(c))) return r;
        k_dt_finalize<<<1, 1, 0, c->stream>>>(c->d_ctl, c->cfg.cfl, c->tab, 0, nullptr);
        CU(cudaGetLastError());
        for (int s = 0; s < c->cfg.num_stages; ++s)
            if ((r = stage_and_refresh(c, s, !no_fuse))) return r;
        c->launches += 1;   // k_dt_finalize: the step boundary (end of the previous step + dt / stop test of this one)
        return 0;
    };
    auto flush_step_end = [&]() -> int {   // the end of the last enqueued step, before the host looks at t / nsteps
        k_step_end<<<1, 1, 0, c->stream>>>(c->d_ctl);
        CU(cudaGetLastError());
        c->launches += 1;
        return 0;
    };
    if (!no_fuse && (rc = launch_dt(c, c->i0, 0))) return rc;   // CFL minimum of the state the call starts from
    static const bool no_graph = getenv("PYH_NO_GRAPH") != nullptr;
    int64_t enqueued = 0;          // steps handed to the stream in this call (>= steps the device executes)
    const int period = (c->cfg.num_stages == 1) ? 2 : 1;   // steps after which the buffer roles repeat
    if (!no_graph && !c->run_graph && max_steps >= 2 * period) {
        // first steps run eagerly (also configures the kernels' attributes), then one period is captured
        for (int n = 0; n < period; ++n) if ((rc = enqueue_step())) return rc;
        const long long l0 = c->launches;
        cudaGraph_t g = nullptr;
        c->run_graph_i0 = c->i0;
        CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        for (int n = 0; n < period; ++n) {
            if ((rc = enqueue_step())) { cudaStreamEndCapture(c->stream, &g); if (g) cudaGraphDestroy(g); return rc; }
        }
        CU(cudaStreamEndCapture(c->stream, &g));
        c->run_graph_launches = c->launches - l0;
        c->launches = l0;   // nothing of the captured period has run yet
        cudaError_t e = cudaGraphInstantiate(&c->run_graph, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) return set_err(PYH_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
        c->run_graph_steps = period;
        max_steps -= period;   // the eager steps count
        enqueued += period;
        if (steps_done) *steps_done = 0;
    }
    int64_t issued = 0;
    while (issued < max_steps) {
        int64_t chunk = std::min<int64_t>(poll_every, max_steps - issued);
        int64_t n = 0;
        if (c->run_graph && !no_graph) {
            // the captured period has the buffer roles of its first step baked in: realign with one eager step if an odd
            // number of single-stage steps has been taken since (ExplicitEuler1 only; roles of longer tableaux repeat every step)
            if (c->i0 != c->run_graph_i0 && n < chunk) { if ((rc = enqueue_step())) return rc; ++n; }
            for (; n + c->run_graph_steps <= chunk; n += c->run_graph_steps) {
                CU(cudaGraphLaunch(c->run_graph, c->stream));
                c->launches += c->run_graph_launches;
            }
        }
        for (; n < chunk; ++n) if ((rc = enqueue_step())) return rc;
        issued += chunk;
        enqueued += chunk;
        if ((rc = flush_step_end())) return rc;
        CU(cudaMemcpyAsync(&h, c->d_ctl, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        if (!h.active || h.bad || !(h.t < h.t_final)) break;
    }
    if (c->cfg.num_stages == 1) {
        // Single-stage tableaux alternate two buffers and the host flipped their roles once per ENQUEUED step, but the device
        // executed only the first h.nsteps of them (the rest early-exit once t >= t_final or the state went bad): the buffer
        // the device wrote last is the authoritative solution.
        CU(cudaMemcpyAsync(&h, c->d_ctl, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        if ((enqueued - (int64_t)h.nsteps) & 1) { std::swap(c->i0, c->i1); c->cur = c->i0; }
    }
    // final realizability check of the last state (Euler2D.py:204): its flag was reduced by the last executed step
    if (no_fuse && (rc = launch_dt(c, c->i0))) return rc;
    if ((rc = reduce_dt(c))) return rc;
    k_dt_finalize<<<1, 1, 0, c->stream>>>(c->d_ctl, c->cfg.cfl, c->tab, 1, c->d_tmp);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(&h, c->d_ctl, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *t_inout = h.t;
    if (steps_done) *steps_done = h.nsteps;
    if (unrealizable) *unrealizable = h.bad;
    if (ddts) {
        int64_t n = std::min<int64_t>(h.nsteps, dts_cap);
        if (n > 0) CU(cudaMemcpy(dts_out, c->d_dts, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
    }
    rc = set_active(c, 1);
    return rc;
}

int pyh_realizable(void* ctx, int32_t* ok_out) {
    Ctx* c = as_ctx(ctx);
    if (!c || !c->finalized) return set_err(PYH_ERR_STATE, "context not finalized");
    if (!ok_out) return set_err(PYH_ERR_INVALID, "null pointer");
    CU(cudaSetDevice(c->cfg.device));
    int zero = 0;
    CU(cudaMemcpyAsync(&c->d_ctl->bad, &zero, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    int rc = launch_dt(c, c->i0);
    if (rc) return rc;
    if ((rc = reduce_dt(c))) return rc;
    k_dt_finalize<<<1, 1, 0, c->stream>>>(c->d_ctl, c->cfg.cfl, c->tab, 1, c->d_tmp);
    CU(cudaGetLastError());
    int bad = 0;
    CU(cudaMemcpyAsync(&bad, &c->d_ctl->bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMemcpyAsync(&c->d_ctl->bad, &zero, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    *ok_out = bad ? 0 : 1;
    return 0;
}

static int ensure_debug_buffers(Ctx* c, bool grad) {
    bool changed = false;
    for (auto& b : c->blocks) {
        if (!b.dev.dbg) {
            void* q;
            CU(cudaMalloc(&q, 4 * (size_t)c->lay.plane * sizeof(double)));
            CU(cudaMemset(q, 0, 4 * (size_t)c->lay.plane * sizeof(double)));
            b.dev.dbg = (double*)q;
            changed = true;
        }
        if (grad && !b.dev.dbgG) {
            void* q;
            CU(cudaMalloc(&q, 12 * (size_t)c->lay.plane * sizeof(double)));
            CU(cudaMemset(q, 0, 12 * (size_t)c->lay.plane * sizeof(double)));
            b.dev.dbgG = (double*)q;
            changed = true;
        }
    }
    if (changed) {
        std::vector<BlkDev> tmp;
        for (auto& b : c->blocks) tmp.push_back(b.dev);
        CU(cudaMemcpy(c->d_blks, tmp.data(), tmp.size() * sizeof(BlkDev), cudaMemcpyHostToDevice));
    }
    return 0;
}

static int residual_common(Ctx* c, int want_grad) {
    // one stage launch that only stores R itself (and optionally gx, gy, phi) into the debug buffers
    int rc = ensure_debug_buffers(c, want_grad != 0);
    if (rc) return rc;
    StagePlan p;
    memset(&p, 0, sizeof(p));
    p.cur = c->po.H[c->i0];
    p.ntargets = 0;
    p.write_residual = 1;
    if ((rc = set_active(c, 1))) return rc;
    return launch_stage(c, p, want_grad);
}

int pyh_residual(void* ctx, int gid, double* aos_out) {
    Ctx* c = as_ctx(ctx);
    GET_BLOCK(c, gid, hb);
    if (!c->finalized) return set_err(PYH_ERR_STATE, "context not finalized");
    if (!aos_out) return set_err(PYH_ERR_INVALID, "null pointer");
    int rc = residual_common(c, 0);
    if (rc) return rc;
    size_t n = (size_t)c->lay.nx * c->lay.ny;
    if ((rc = ensure_scratch(c, 4 * n * sizeof(double)))) return rc;
    k_soa_to_aos<<<cdiv(n, 256), 256, 0, c->stream>>>(c->lay, hb.dev.dbg, c->d_scratch, 4);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(aos_out, c->d_scratch, 4 * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int pyh_debug_fetch(void* ctx, int gid, int what, double* aos_out) {
    Ctx* c = as_ctx(ctx);
    GET_BLOCK(c, gid, hb);
    if (!c->finalized) return set_err(PYH_ERR_STATE, "context not finalized");
    if (what < 0 || what > 2 || !aos_out) return set_err(PYH_ERR_INVALID, "bad selector / null pointer");
    int rc;
    if ((rc = residual_common(c, 1))) return rc;
    size_t n = (size_t)c->lay.nx * c->lay.ny;
    if ((rc = ensure_scratch(c, 4 * n * sizeof(double)))) return rc;
    k_soa_to_aos<<<cdiv(n, 256), 256, 0, c->stream>>>(c->lay, hb.dev.dbgG + 4 * (size_t)what * c->lay.plane, c->d_scratch, 4);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(aos_out, c->d_scratch, 4 * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int pyh_comm_unique_id(void* id_out) {
    if (!id_out) return set_err(PYH_ERR_INVALID, "null pointer");
    NcclApi& N = nccl_api();
    if (N.error) return set_err(PYH_ERR_STATE, "NCCL unavailable: %s", N.error);
    NcclUniqueId id;
    NC(N.GetUniqueId(&id));
    memcpy(id_out, &id, sizeof(id));
    return 0;
}

int pyh_comm_init(void* ctx, int32_t rank, int32_t world, const void* id, const int32_t* owner, int32_t nblocks_total) {
    Ctx* c = as_ctx(ctx);
    if (!c || !c->finalized) return set_err(PYH_ERR_STATE, "context not finalized");
    if (!id || !owner || world < 1 || rank < 0 || rank >= world) return set_err(PYH_ERR_INVALID, "bad rank / world / null pointer");
    if (c->comm.comm) return set_err(PYH_ERR_STATE, "pyh_comm_init called twice");
    NcclApi& N = nccl_api();
    if (N.error) return set_err(PYH_ERR_STATE, "NCCL unavailable: %s", N.error);
    CU(cudaSetDevice(c->cfg.device));
    Comm& m = c->comm;
    m.rank = rank; m.world = world;
    // message lists: the strip this rank sends for slot (gid, side) is named (gid, side); the strip it receives for that slot
    // is the neighbour's (nbr_gid, opposite side) -- the same ordering rule as pyhype_b200/distributed.py:exchange_plan
    static const int opposite[4] = {PYH_WEST, PYH_EAST, PYH_SOUTH, PYH_NORTH};
    for (size_t i = 0; i < c->slots.size(); ++i) {
        const HaloSlot& hs = c->slots[i];
        const int ng = c->slot_nbr_gid[i];
        if (ng < 0 || ng >= nblocks_total) return set_err(PYH_ERR_INVALID, "neighbour block %d outside the owner table (%d blocks)", ng, nblocks_total);
        const int peer = owner[ng];
        if (peer < 0 || peer >= world || peer == rank) return set_err(PYH_ERR_INVALID, "block %d: remote neighbour %d is owned by rank %d", c->blocks[hs.blk].d.gid, ng, peer);
        const long long len = 4LL * ((hs.side == PYH_EAST || hs.side == PYH_WEST) ? c->lay.ny : c->lay.nx);
        m.sends.push_back(HaloMsg{peer, c->blocks[hs.blk].d.gid, hs.side, hs.offset, len});
        m.recvs.push_back(HaloMsg{peer, ng, opposite[hs.side], hs.offset, len});
    }
    for (const HaloSlot& hs : c->slots) {
        if (hs.side == PYH_NORTH || hs.side == PYH_SOUTH) c->split_ns = true;
        else c->split_ew = true;
    }
    if (!c->slots.empty() && !c->s_edge) {
        int lo = 0, hi = 0;   // edge strips + exchange win the first free thread-block slots
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CU(cudaStreamCreateWithPriority(&c->s_edge, cudaStreamNonBlocking, hi));
        CU(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->ev_edge_done, cudaEventDisableTiming));
    }
    std::sort(m.sends.begin(), m.sends.end(), msg_less);
    std::sort(m.recvs.begin(), m.recvs.end(), msg_less);
    m.doubles = c->halo_doubles;   // send / receive buffers: allocated by pyh_finalize
    NcclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    NC(N.CommInitRank(&m.comm, world, uid, rank));
    // the cached single-rank graph (if any) does not contain the exchange
    if (c->run_graph) { cudaGraphExecDestroy(c->run_graph); c->run_graph = nullptr; }
    return 0;
}

int pyh_comm_info(void* ctx, int32_t* rank, int32_t* world, int32_t* n_msgs, int64_t* doubles_per_exchange) {
    Ctx* c = as_ctx(ctx);
    if (!c) return set_err(PYH_ERR_INVALID, "null context");
    if (rank) *rank = c->comm.comm ? c->comm.rank : 0;
    if (world) *world = c->comm.comm ? c->comm.world : 1;
    if (n_msgs) *n_msgs = (int32_t)c->comm.sends.size();
    if (doubles_per_exchange) *doubles_per_exchange = c->comm.doubles;
    return 0;
}

int pyh_march_shape(void* ctx, int32_t* lanes, int32_t* rows) {
    Ctx* c = as_ctx(ctx);
    if (!c || !c->finalized) return set_err(PYH_ERR_STATE, "context not finalized");
    if (lanes) *lanes = c->march_nt;
    if (rows) *rows = c->march_tys;
    return 0;
}

int pyh_stage_path(void* ctx, int32_t* split, double* tuned_ms) {
    Ctx* c = as_ctx(ctx);
    if (!c || !c->finalized) return set_err(PYH_ERR_STATE, "context not finalized");
    if (split) *split = c->use_split ? 1 : 0;
    if (tuned_ms) { tuned_ms[0] = c->tune_ms[0]; tuned_ms[1] = c->tune_ms[1]; }
    return 0;
}

int pyh_launch_count(void* ctx, int64_t* n) {
    Ctx* c = as_ctx(ctx);
    if (!c || !n) return set_err(PYH_ERR_INVALID, "null argument");
    *n = c->launches;
    return 0;
}

int pyh_stream(void* ctx, uint64_t* out) {
    Ctx* c = as_ctx(ctx);
    if (!c || !out) return set_err(PYH_ERR_INVALID, "null argument");
    *out = (uint64_t)(uintptr_t)c->stream;
    return 0;
}

int pyh_sync(void* ctx) {
    Ctx* c = as_ctx(ctx);
    if (!c) return set_err(PYH_ERR_INVALID, "null context");
    CU(cudaSetDevice(c->cfg.device));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

}  // extern "C"

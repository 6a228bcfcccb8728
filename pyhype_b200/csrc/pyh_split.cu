// Instantiations of the three-kernel stage of small problems (pyh_stage_split.cuh): a translation unit of its own so that it
// compiles in parallel with the fused stage kernels (see __graft_entry__.build).
#include <cstring>
#include "pyh_stage_split.cuh"

namespace pyh {
// Occupancy targets (launch bounds), measured on the B200 (profiles/r02z_split_occupancy_ab.txt).  Problems of many waves are
// throughput-bound and gain from more resident warps even at the price of a few spilled registers (`dense`: recon 4 x 256 threads at
// 64 registers, flux 8 x 128 at 64; DMR 0.416 -> 0.405 ms/step); problems of one or two waves are latency-bound and lose
// (explosion_multi 0.153 -> 0.159), they keep recon 3 x 256 and flux 6 x 128 at 80 registers.  For the HLLL flux the step from
// ptxas' own choice (122 registers, 4 x 128) to 80 registers alone was worth 5 % on DMR.
template <int L>
static SplitReconFn rpick(int p, bool dense) {
    if (dense) return p ? k_split_recon<L, 1, 4> : k_split_recon<L, 0, 4>;
    return p ? k_split_recon<L, 1, 3> : k_split_recon<L, 0, 3>;
}
template <int F>
static SplitFluxFn fpick(int p, bool dense) {
    if (dense) return p ? k_split_flux<F, 1, 8> : k_split_flux<F, 0, 8>;
    return p ? k_split_flux<F, 1, 6> : k_split_flux<F, 0, 6>;
}

SplitReconFn pick_split_recon(int l, int p, bool dense) {
    switch (l) {
        case 0: return rpick<0>(p, dense);
        case 1: return rpick<1>(p, dense);
        case 2: return rpick<2>(p, dense);
        default: return rpick<3>(p, dense);
    }
}
SplitFluxFn pick_split_flux(int f, int p, bool dense) {
    switch (f) {
        case 0: return fpick<0>(p, dense);
        case 1: return fpick<1>(p, dense);
        default: return fpick<2>(p, dense);
    }
}
struct SplitLaunchOpts { bool pdl; const void* win_base; size_t win_bytes; float win_hit; };   // same as in pyh_api.cu
// cudaLaunchKernelEx with / without programmatic stream serialization (pyh_stage_split.cuh: pdl_wait / pdl_trigger) and with / without
// a persisting-L2 access-policy window over the stage scratch (written by one kernel of the stage, read by the next)
template <typename... P, typename... A>
static cudaError_t launch_ex(void (*fn)(P...), dim3 grid, int threads, cudaStream_t st, const SplitLaunchOpts& o, A... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    memset(at, 0, sizeof(at));
    unsigned n = 0;
    if (o.pdl) {
        at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (o.win_hit > 0.f && o.win_bytes > 0) {
        at[n].id = cudaLaunchAttributeAccessPolicyWindow;
        at[n].val.accessPolicyWindow.base_ptr = const_cast<void*>(o.win_base);
        at[n].val.accessPolicyWindow.num_bytes = o.win_bytes;
        at[n].val.accessPolicyWindow.hitRatio = o.win_hit;
        at[n].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        at[n].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        ++n;
    }
    cfg.attrs = at;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, fn, args...);
}
cudaError_t launch_split_recon(SplitReconFn fn, dim3 grid, cudaStream_t st, SplitLaunchOpts pdl, const BlkDev* blks, const Layout lay, const PlaneOffsets po,
                               const unsigned cur, const Control* ctl, const Consts C) {
    return launch_ex(fn, grid, kSplitReconThreads, st, pdl, blks, lay, po, cur, ctl, C);
}
cudaError_t launch_split_flux(SplitFluxFn fn, dim3 grid, cudaStream_t st, SplitLaunchOpts pdl, const BlkDev* blks, const Layout lay, const PlaneOffsets po,
                              const unsigned cur, const Control* ctl, const Consts C) {
    return launch_ex(fn, grid, kSplitFluxThreads, st, pdl, blks, lay, po, cur, ctl, C);
}
cudaError_t launch_split_update(dim3 grid, cudaStream_t st, SplitLaunchOpts pdl, bool dense, const BlkDev* blks, const Layout lay, const PlaneOffsets po,
                                const StagePlan plan, const Control* ctl, Control* ctl_out, const Consts C) {
    if (dense) return launch_ex(k_split_update<4>, grid, kSplitUpdateThreads, st, pdl, blks, lay, po, plan, ctl, ctl_out, C);
    return launch_ex(k_split_update<5>, grid, kSplitUpdateThreads, st, pdl, blks, lay, po, plan, ctl, ctl_out, C);
}
}  // namespace pyh

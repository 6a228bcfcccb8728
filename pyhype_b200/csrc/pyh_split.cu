// Instantiations of the three-kernel stage of small problems (pyh_stage_split.cuh): a translation unit of its own so that it
// compiles in parallel with the fused stage kernels (see __graft_entry__.build).
#include "pyh_stage_split.cuh"

namespace pyh {
template <int L>
static SplitReconFn rpick(int p) { return p ? k_split_recon<L, 1> : k_split_recon<L, 0>; }
template <int F>
static SplitFluxFn fpick(int p) { return p ? k_split_flux<F, 1> : k_split_flux<F, 0>; }

SplitReconFn pick_split_recon(int l, int p) {
    switch (l) {
        case 0: return rpick<0>(p);
        case 1: return rpick<1>(p);
        case 2: return rpick<2>(p);
        default: return rpick<3>(p);
    }
}
SplitFluxFn pick_split_flux(int f, int p) {
    switch (f) {
        case 0: return fpick<0>(p);
        case 1: return fpick<1>(p);
        default: return fpick<2>(p);
    }
}
void launch_split_update(dim3 grid, cudaStream_t st, const BlkDev* blks, const Layout lay, const PlaneOffsets po, const StagePlan plan,
                         const Control* ctl, Control* ctl_out, const Consts C) {
    k_split_update<<<grid, kSplitUpdateThreads, 0, st>>>(blks, lay, po, plan, ctl, ctl_out, C);
}
}  // namespace pyh

// Stage-kernel instantiations for fvm_num_quadrature_points == 2 (one translation unit per value so that
// the three compile in parallel; see __graft_entry__.build).
#include "pyh_stage_march.cuh"

namespace pyh {
template <int F, int L>
static MarchFn mpick_p(int p) { return p ? k_stage_march<F, L, 1, 2> : k_stage_march<F, L, 0, 2>; }
template <int F>
static MarchFn mpick_l(int l, int p) {
    switch (l) {
        case 0: return mpick_p<F, 0>(p);
        case 1: return mpick_p<F, 1>(p);
        case 2: return mpick_p<F, 2>(p);
        default: return mpick_p<F, 3>(p);
    }
}
MarchFn pick_march_nq2(int f, int l, int p) {
#ifdef PYH_ONLY_ROE_VENKAT_CONS   // kernel-tuning builds (tools/build_variant.sh): one instantiation
    (void)f; (void)l; (void)p;
#ifndef PYH_ONLY_F
#define PYH_ONLY_F 0
#define PYH_ONLY_L 0
#define PYH_ONLY_P 0
#endif
    return k_stage_march<PYH_ONLY_F, PYH_ONLY_L, PYH_ONLY_P, 2>;
#else
    switch (f) {
        case 0: return mpick_l<0>(l, p);
        case 1: return mpick_l<1>(l, p);
        default: return mpick_l<2>(l, p);
    }
#endif
}
}  // namespace pyh

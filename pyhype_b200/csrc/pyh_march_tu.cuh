// pyh_march_tu.cuh -- what pyh_api.cu needs to know about the stage kernel without instantiating it.
#pragma once
#include "pyh_layout.cuh"
#include "pyh_math.cuh"

namespace pyh {
#ifndef PYH_MARCH_MAXT
#define PYH_MARCH_MAXT 128
#endif
#ifndef PYH_MARCH_MINB
#define PYH_MARCH_MINB 4
#endif
constexpr int MARCH_MAX_THREADS = PYH_MARCH_MAXT;
// shared-memory doubles per thread for NQ quadrature points per face:
// sQ[3][4], sFE[2][NQ][4], sIW[2][4], sQN[2][NQ][4], sIS[4], sQW[NQ][4], sQS[NQ][4]
constexpr int march_smem_doubles(int nq) { return 24 + 24 * nq; }
// resident CTAs per SM the launch bounds ask for (shared memory is the limiter for NQ > 1)
constexpr int march_min_blocks(int nq) { return nq == 1 ? PYH_MARCH_MINB : (nq == 2 ? 3 : 2); }
}  // namespace pyh

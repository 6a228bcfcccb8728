// pyh_march_tu.cuh -- what pyh_api.cu needs to know about the stage kernel without instantiating it.
#pragma once
#include "pyh_layout.cuh"
#include "pyh_math.cuh"

namespace pyh {
// Launch bounds.  One quadrature point: up to 160 threads (156 output columns: one strip covers explosion_multi's 150-column
// blocks) at >= 3 thread blocks per SM.  The SAME code also runs as 4 thread blocks of 128 threads -- the shape of large
// problems -- as long as ptxas stays within 128 registers, which it does (126-128) and which tests/test_abi.py enforces
// (-maxrregcount is ignored for kernels with launch bounds).  2 / 3 points need more shared memory per thread: 128 threads.
#ifndef PYH_MARCH_MAXT
#define PYH_MARCH_MAXT 160
#endif
#ifndef PYH_MARCH_MINB
#define PYH_MARCH_MINB 3
#endif
constexpr int march_max_threads(int nq) { return nq == 1 ? PYH_MARCH_MAXT : 128; }

// Which strips of every block one launch of the stage kernel covers.  A plain launch covers the whole block; a context with remote
// neighbours splits every stage into an EDGE launch (the thin strips that produce the cells other ranks need) and an INTERIOR
// launch, so that the strip exchange runs behind the interior (pyh_api.cu: stage_and_refresh).
struct MarchTiles {
    int row0, rowstride, row1;   // row strip blockIdx.y covers rows [row0 + y * rowstride, min(that + tys, row1))
    int xfirst, xstride;         // column strip = xfirst + blockIdx.x * xstride
};
struct BlkDev;
struct Control;
typedef void (*MarchFn)(const BlkDev*, const Layout, const PlaneOffsets, const StagePlan, const Control*, Control*, const Consts, const int, const int, const MarchTiles);
// the three-kernel stage (pyh_stage_split.cuh)
typedef void (*SplitReconFn)(const BlkDev*, const Layout, const PlaneOffsets, const unsigned, const Control*, const Consts);
typedef void (*SplitFluxFn)(const BlkDev*, const Layout, const PlaneOffsets, const unsigned, const Control*, const Consts);
constexpr int kSplitTX = 32, kSplitTY = 8;                 // cell tile of k_split_recon (one thread per cell)
constexpr int kSplitReconThreads = kSplitTX * kSplitTY;
constexpr int kSplitFluxThreads = 128;
constexpr int kSplitUpdateThreads = 256;
constexpr int kSplitStatePlanes = 16, kSplitFluxPlanes = 8;  // FS: E, W, N, S face states x 4 variables; FX: vertical, horizontal face fluxes x 4
constexpr int kSplitPlanes = kSplitStatePlanes + kSplitFluxPlanes;
constexpr long long kSplitDenseCells = 500000;               // above: the split kernels' high-occupancy builds (explosion_multi is 180 k cells, DMR 1 M)
constexpr long long kSplitMaxCells = 4500000;               // above: always the fused kernel (the scratch planes cost 192 B per cell; 8 x 1024^2 was 20 % slower split)
// shared-memory doubles per thread for NQ quadrature points per face:
// sQ[3][4], sFE[2][NQ][4], sIW[2][4], sQN[2][NQ][4], sIS[4], sQW[NQ][4], sQS[NQ][4]
// (48 doubles per thread for NQ = 1: 4 thread blocks of 128 threads are exactly the 196 KB shared-memory carve-out, which
// leaves 60 KB of L1; ONE more double per thread selects the 228 KB carve-out and costs 10 % -- measured, profiles/r02b)
constexpr int march_smem_doubles(int nq) { return 24 + 24 * nq; }
// resident CTAs per SM the launch bounds ask for (shared memory is the limiter for NQ > 1)
constexpr int march_min_blocks(int nq) { return nq == 1 ? PYH_MARCH_MINB : (nq == 2 ? 3 : 2); }
}  // namespace pyh

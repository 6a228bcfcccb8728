// Host twin of the device arithmetic: pyhype_b200/csrc/pyh_math.cuh + pyh_fastdiv.cuh compiled by g++ through the
// shim in tests/host_twin/shim (see there).  extern "C" entry points evaluate the SAME source the stage kernel
// inlines -- the plain-operator policy Ar<false> and the branch-free fast policy Ar<true> -- on arrays, so that
// tests/test_host_twin.py can compare them with oracle/muscl_oracle.py bit for bit without a GPU.
// Build: g++ -O2 -ffp-contract=off -std=c++20 -pthread -shared -fPIC -I tests/host_twin/shim -I pyhype_b200/csrc twin.cpp
#define PYH_HOST_TWIN 1
#include <cuda_runtime.h>   // the shim
#include "pyh_math.cuh"

using namespace pyh;

static Consts make_consts(double g) {
    Consts C;
    C.g = g;
    C.gm1 = g - 1;                // Python: g - 1
    C.k = 1.0 / (g - 1.0);
    C.gm = g / (g - 1.0);
    for (int i = 0; i < 3; ++i) { C.qw[i] = 0.0; C.qp[i] = 0.0; }
    C.qw[0] = 2.0;
    return C;
}

template <int FLUX, int PRIM, bool FAST>
static void riemann_n(long n, const double* QL, const double* QR, double* F, int* okf, const Consts& C) {
    for (long i = 0; i < n; ++i) {
        double l[4], r[4], f[4] = {0, 0, 0, 0};
        for (int k = 0; k < 4; ++k) { l[k] = QL[4 * i + k]; r[k] = QR[4 * i + k]; }
        bool ok = true;
        riemann_flux<FLUX, PRIM, FAST>(l, r, f, C, ok);
        for (int k = 0; k < 4; ++k) F[4 * i + k] = f[k];
        okf[i] = ok ? 1 : 0;
    }
}

template <int LIM>
static void limiter_n(int fast, long n, const double* dmx, const double* dmn, const double* davg, double* phi, int* okf) {
    for (long i = 0; i < n; ++i) {
        double d[4] = {davg[4 * i], davg[4 * i + 1], davg[4 * i + 2], davg[4 * i + 3]};
        double p = 0.0;
        bool ok = true;
        if (fast) ok = limiter4_fast<LIM>(dmx[i], dmn[i], d, p);
        else limiter4_safe<LIM>(dmx[i], dmn[i], d, p);
        phi[i] = p;
        okf[i] = ok ? 1 : 0;
    }
}

extern "C" {

// scale of the flux returned by twin_riemann relative to the reference's F (2 when the power-of-two folding of
// the Roe solver is compiled in, see PYH_FOLD_POW2 in pyh_math.cuh)
double twin_flux_scale(int flux, int fast) { return flux_scale(flux, fast != 0); }
int twin_fold_pow2() { return PYH_FOLD_POW2; }

int twin_riemann(int flux, int prim, int fast, double gamma, long n, const double* QL, const double* QR, double* F, int* okf) {
    const Consts C = make_consts(gamma);
#define CASE(FL, PR)                                                                      \
    if (flux == FL && prim == PR) {                                                       \
        if (fast) riemann_n<FL, PR, true>(n, QL, QR, F, okf, C);                           \
        else riemann_n<FL, PR, false>(n, QL, QR, F, okf, C);                               \
        return 0;                                                                         \
    }
    CASE(0, 0) CASE(0, 1) CASE(1, 0) CASE(1, 1) CASE(2, 0) CASE(2, 1)
#undef CASE
    return -1;
}

int twin_limiter4(int lim, int fast, long n, const double* dmx, const double* dmn, const double* davg, double* phi, int* okf) {
    switch (lim) {
        case 0: limiter_n<0>(fast, n, dmx, dmn, davg, phi, okf); return 0;
        case 1: limiter_n<1>(fast, n, dmx, dmn, davg, phi, okf); return 0;
        case 2: limiter_n<2>(fast, n, dmx, dmn, davg, phi, okf); return 0;
        case 3: limiter_n<3>(fast, n, dmx, dmn, davg, phi, okf); return 0;
    }
    return -1;
}

// mode 0: integer emulation of the x87 sequence; mode 1: double-double emulation, two vectors at a time
void twin_nrm2(int mode, long n, const double* x, double* out, int* okf) {
    if (mode == 0) {
        for (long i = 0; i < n; ++i) { out[i] = nrm2_x87(x + 4 * i); okf[i] = 1; }
        return;
    }
    for (long i = 0; i + 1 < n; i += 2) {
        double o[2];
        bool ok[2];
        nrm2_x87_dd2(x + 4 * i, x + 4 * (i + 1), o, ok);
        out[i] = o[0]; out[i + 1] = o[1];
        okf[i] = ok[0]; okf[i + 1] = ok[1];
    }
    if (n & 1) {
        double o[2];
        bool ok[2];
        nrm2_x87_dd2(x + 4 * (n - 1), x + 4 * (n - 1), o, ok);
        out[n - 1] = o[0]; okf[n - 1] = ok[0];
    }
}

// op 0: a / b, 1: 1 / b, 2: sqrt(b); the fast sequences of pyh_fastdiv.cuh
void twin_arith(int op, long n, const double* a, const double* b, double* out, int* okf) {
    for (long i = 0; i < n; ++i) {
        bool ok = true;
        out[i] = op == 0 ? div_fast(a[i], b[i], ok) : op == 1 ? rcp_fast(b[i], ok) : sqrt_fast(b[i], ok);
        okf[i] = ok ? 1 : 0;
    }
}

void twin_cons2prim(int fast, double gamma, long n, const double* U, double* W, int* okf) {
    const Consts C = make_consts(gamma);
    for (long i = 0; i < n; ++i) {
        double q[4] = {U[4 * i], U[4 * i + 1], U[4 * i + 2], U[4 * i + 3]};
        bool ok = true;
        if (fast) cons2prim<true>(q, C, ok); else cons2prim<false>(q, C, ok);
        for (int k = 0; k < 4; ++k) W[4 * i + k] = q[k];
        okf[i] = ok ? 1 : 0;
    }
}

}  // extern "C"

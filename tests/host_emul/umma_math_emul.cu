// umma_math_emul.cu — DEVELOPMENT / TEST HARNESS ONLY (never shipped, never loaded by the product).
// Compiles the arithmetic helpers of the tcgen05 ensemble integrator (csrc/umma_math.cuh: branch-free sincos,
// 52-bit fixed-point split, recombination of the radix-256 digit sums) for the HOST, so that their accuracy and
// exactness can be checked in the CPU-only container (tests/test_umma_math_on_host.py).
#include <cstdint>

#include "../../mcbb.jl_b200/csrc/umma_math.cuh"

using namespace mcbb;

extern "C" void emul_um_sincos(const double* x, int64_t n, double* s, double* c) {
    int64_t i = 0;
    for (; i + 4 <= n; i += 4) {  // the kernel evaluates groups side by side; same arithmetic per element
        const double xs[4] = {x[i], x[i + 1], x[i + 2], x[i + 3]};
        um_sincos_vec<4>(xs, s + i, c + i);
    }
    for (; i < n; i++) {
        const double xs[1] = {x[i]};
        um_sincos_vec<1>(xs, s + i, c + i);
    }
}

extern "C" void emul_um_fixed52(const double* x, int64_t n, uint32_t* lo, uint32_t* hi) {
    for (int64_t i = 0; i < n; i++) {
        unsigned a, b;
        um_fixed52(x[i], a, b);
        lo[i] = a;
        hi[i] = b;
    }
}

// v: n x 7 digit sums (row-major), degh = deg << 19
extern "C" void emul_um_recombine(const uint32_t* v, const int32_t* degh, int64_t n, double* out) {
    for (int64_t i = 0; i < n; i++) {
        const uint32_t* w = v + 7 * i;
        out[i] = um_recombine(w[0], w[1], w[2], w[3], w[4], w[5], w[6], degh[i]);
    }
}

// peaks.cu — in-process roofline denominators the driver-written MEASURED_PEAKS.json does not carry:
// the FP64 FMA-pipe peak (DFMA micro-benchmark), which bounds the ensemble integrator.
#include "common.cuh"

namespace mcbb {

// 16 independent FMA chains per thread, 4 warps per scheduler: the FP64 pipe never waits on a dependency.
__global__ void __launch_bounds__(512) k_dfma_peak(double* out, int iters, double seed) {
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = seed + i + threadIdx.x * 1e-3;
    const double m = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = fma(a[i], m, c);
    }
    double s = 0.;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
    if (s == 12345.678) out[0] = s;  // keeps the chains alive without a store in practice
}

}  // namespace mcbb

using namespace mcbb;

extern "C" int mcbb_measure_fp64_fma_peak(double* tflops_out, double* sm_mhz_est_out) {
    if (!tflops_out) MCBB_FAIL("measure_fp64_fma_peak: NULL argument");
    int dev = 0, sms = 0;
    MCBB_CUDA_OK(cudaGetDevice(&dev));
    MCBB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double* d = (double*)scratch_get(60, 64);
    if (!d) MCBB_FAIL("device allocation failed");
    cudaEvent_t e0, e1;
    MCBB_CUDA_OK(cudaEventCreate(&e0));
    MCBB_CUDA_OK(cudaEventCreate(&e1));
    const int iters = 1 << 15, grid = sms * 4, block = 512;
    double best = 0.;
    for (int rep = 0; rep < 6; rep++) {
        MCBB_CUDA_OK(cudaEventRecord(e0, 0));
        k_dfma_peak<<<grid, block>>>(d, iters, 1.0 + rep);
        MCBB_LAUNCHED();
        MCBB_CUDA_OK(cudaEventRecord(e1, 0));
        MCBB_CUDA_OK(cudaEventSynchronize(e1));
        float ms = 0.f;
        MCBB_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 16.0 * (double)iters * (double)grid * (double)block;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops_out = best;
    // 64 DFMA / clk / SM on sm_100: implied SM clock
    if (sm_mhz_est_out) *sm_mhz_est_out = best * 1e12 / (2.0 * 64.0 * sms) / 1e6;
    return 0;
}

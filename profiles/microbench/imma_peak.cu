// Micro-benchmark: throughput of legacy mma.sync INT8 (m16n8k32) and FP64 DMMA (m8n8k4) on sm_100a,
// to decide whether the 0/1-adjacency batch GEMM of the integrator should move to tensor cores.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_imma(int* out, int iters) {
    int a0 = threadIdx.x, a1 = 2, a2 = 3, a3 = 4, b0 = 5, b1 = 6;
    int c[8][4] = {};
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(c[u][0]), "+r"(c[u][1]), "+r"(c[u][2]), "+r"(c[u][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    int s = 0;
    for (int u = 0; u < 8; u++) s += c[u][0] + c[u][1] + c[u][2] + c[u][3];
    if (s == 123456789) out[0] = s;
}
__global__ void k_dmma(double* out, int iters) {
    double a = threadIdx.x * 1e-3, b = 1.0000001;
    double c[8][2] = {};
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[u][0]), "+d"(c[u][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int u = 0; u < 8; u++) s += c[u][0] + c[u][1];
    if (s == 1234.5) out[0] = s;
}
int main() {
    int* d; cudaMalloc(&d, 64); double* dd; cudaMalloc(&dd, 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int warps = 4; warps <= 16; warps *= 2) {
        const int iters = 1 << 14, grid = sms * 2, block = warps * 32;
        float ms;
        for (int rep = 0; rep < 3; rep++) { cudaEventRecord(e0); k_imma<<<grid, block>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); }
        double ops = 2.0 * 16 * 8 * 32 * 8.0 * iters * grid * warps;
        printf("IMMA m16n8k32 warps/CTA=%d: %.1f TOPS  (%.2f clk per mma per SMSP @1.965GHz)\n", warps, ops / ms / 1e9,
               ms * 1e-3 * 1.965e9 / (8.0 * iters * (grid * warps / (double)(sms * 4))));
        for (int rep = 0; rep < 3; rep++) { cudaEventRecord(e0); k_dmma<<<grid, block>>>(dd, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); }
        double fl = 2.0 * 8 * 8 * 4 * 8.0 * iters * grid * warps;
        printf("DMMA m8n8k4   warps/CTA=%d: %.2f TFLOPS\n", warps, fl / ms / 1e9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

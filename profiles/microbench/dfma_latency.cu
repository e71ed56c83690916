// dfma_latency.cu — dependent-issue latency and per-SM throughput of DFMA on B200 as a function of ILP and warps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o dfma_latency dfma_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k(double* out, long long* clk, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-3 + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

template <int ILP>
void run(int warps, double* out, long long* clk) {
    const int iters = 2000;
    k<ILP><<<1, 32 * warps>>>(out, clk, iters, 0.999999, 1e-9);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
    const double per = (double)c / (iters * 8.0 * ILP);
    printf("warps/SM %2d  ILP %d : %.2f clk per DFMA per warp, SM rate %.2f warp-DFMA/clk (%.1f lanes/clk)\n", warps, ILP, per,
           warps / per, 32.0 * warps / per);
}

int main() {
    double* out;
    long long* clk;
    cudaMalloc(&out, 8 * 1024 * 64);
    cudaMalloc(&clk, 64);
    for (int w : {1, 4, 8, 16, 32}) {
        run<1>(w, out, clk);
        run<2>(w, out, clk);
        run<4>(w, out, clk);
        run<8>(w, out, clk);
    }
    return 0;
}

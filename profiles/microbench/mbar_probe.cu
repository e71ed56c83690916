// mbar_probe.cu — what does mbarrier.pending_count report for the state returned by mbarrier.arrive?
// (used by batched_umma.cu: the warp whose arrival completes the phase issues the MMAs)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(unsigned* out) {
    __shared__ __align__(8) unsigned long long bar;
    const unsigned b = (unsigned)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(b));
    __syncthreads();
    for (int round = 0; round < 2; round++) {
        const int w = threadIdx.x >> 5;
        for (int turn = 0; turn < 4; turn++) {
            if (w == (turn + round) % 4 && (threadIdx.x & 31) == 0) {
                unsigned long long st;
                unsigned pend;
                asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 %0, [%1];" : "=l"(st) : "r"(b) : "memory");
                asm volatile("mbarrier.pending_count.b64 %0, %1;" : "=r"(pend) : "l"(st));
                unsigned ok = 0;
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(b), "r"((unsigned)round) : "memory");
                out[round * 8 + turn * 2] = pend;
                out[round * 8 + turn * 2 + 1] = ok;
            }
            __syncthreads();
        }
    }
}
int main() {
    unsigned* d;
    cudaMalloc(&d, 64);
    k<<<1, 128>>>(d);
    unsigned h[16];
    cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    for (int r = 0; r < 2; r++) {
        printf("phase %d: ", r);
        for (int t = 0; t < 4; t++) printf("arrival %d -> pending_count %u, phase complete %u; ", t, h[r * 8 + 2 * t], h[r * 8 + 2 * t + 1]);
        printf("\n");
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}

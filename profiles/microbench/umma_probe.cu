// umma_probe.cu — pins the tcgen05.mma kind::i8 operand encodings used by batched_umma.cu on real sm_100a hardware.
//   D[128 x N] (s32, TMEM) = A[128 x 128] (u8, K-major, no swizzle) * B[128 x N] (u8, MN-major or K-major, no swizzle)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu
// Run:   umma_probe <amode> <bmode> <N> [reps]     exit 0 = D matches the CPU product bit for bit
//   amode 0: LBO = 128 (between the two 16-byte K chunks), SBO = 1024 (between 8-row groups); 1: swapped
//         2: A from tensor memory (row r in lane r, 4 K bytes per 32-bit column, written with tcgen05.st)
//   bmode 0: MN-major, LBO = 128 (between 8-deep K groups), SBO = 2048 (between 16-column chunks); 1: swapped
//         2: K-major,  LBO = 128, SBO = 1024; 3: swapped
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);     \
            return 2;                                                                           \
        }                                                                                       \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    return d;                // base offset 0, layout type 0 = no swizzle
}

struct Params {
    int amode, bmode, N, reps;
};

__global__ void __launch_bounds__(128, 1) k_probe(const uint8_t* A, const uint8_t* Bm, int* D, long long* clk, Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;                 // 16 KB
    uint8_t* sB = smem + 16384;         // up to 32 KB
    __shared__ uint32_t tmem_base;
    __shared__ __align__(8) uint64_t mbar;
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int N = p.N;
    // A: row r, k byte kb -> (r/8)*1024 + (kb/16)*128 + (r%8)*16 + kb%16
    for (int i = tid; i < 128 * 128; i += 128) {
        const int r = i >> 7, kb = i & 127;
        sA[(r >> 3) * 1024 + (kb >> 4) * 128 + (r & 7) * 16 + (kb & 15)] = A[r * 128 + kb];
    }
    // B logical [k][n] (k = row of B, n = column)
    for (int i = tid; i < 128 * N; i += 128) {
        const int k = i / N, n = i - k * N;
        int off;
        if (p.bmode < 2)
            off = (n >> 4) * 2048 + k * 16 + (n & 15);  // MN-major: 16 columns contiguous per k
        else
            off = (n >> 3) * 1024 + (k >> 4) * 128 + (n & 7) * 16 + (k & 15);  // K-major
        sB[off] = Bm[k * N + n];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    }
    if (wid == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base;
    const uint32_t ta = tb + 256u;  // A operand: columns 256..287 (allocation below covers 512 columns)
    if (p.amode == 2) {
        uint32_t w[32];
        for (int c = 0; c < 32; c++) {
            const uint8_t* q = A + tid * 128 + 4 * c;
            w[c] = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
        }
        const uint32_t taddr = ta + ((uint32_t)(32 * wid) << 16);
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
            "%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
            ::"r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]), "r"(w[9]),
            "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]), "r"(w[16]), "r"(w[17]), "r"(w[18]), "r"(w[19]),
            "r"(w[20]), "r"(w[21]), "r"(w[22]), "r"(w[23]), "r"(w[24]), "r"(w[25]), "r"(w[26]), "r"(w[27]), "r"(w[28]), "r"(w[29]),
            "r"(w[30]), "r"(w[31]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }

    const uint32_t a_lbo = p.amode == 0 ? 128 : 1024, a_sbo = p.amode == 0 ? 1024 : 128;
    uint32_t b_lbo, b_sbo;
    if (p.bmode == 0) { b_lbo = 128; b_sbo = 2048; }
    else if (p.bmode == 1) { b_lbo = 2048; b_sbo = 128; }
    else if (p.bmode == 2) { b_lbo = 128; b_sbo = 1024; }
    else { b_lbo = 1024; b_sbo = 128; }
    const uint32_t b_major = p.bmode < 2 ? 1u : 0u;
    // instruction descriptor: D = s32, A = u8, B = u8, A K-major, N >> 3, M >> 4
    const uint32_t idesc = (2u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (b_major << 16) | ((uint32_t)(N >> 3) << 17) |
                           ((128u >> 4) << 24);
    long long t0 = 0, t1 = 0;
    uint32_t phase = 0;
    for (int rep = 0; rep < p.reps; rep++) {
        if (tid == 0) {
            if (rep == 1) t0 = clock64();
#pragma unroll
            for (int ks = 0; ks < 4; ks++) {
                const uint64_t da = make_desc(smem_u32(sA) + ks * 256, a_lbo, a_sbo);
                // 32 k rows further: MN-major 32 * 16 bytes; K-major two 16-byte chunks = 2 * 128
                const uint64_t db = make_desc(smem_u32(sB) + (p.bmode < 2 ? ks * 512 : ks * 256), b_lbo, b_sbo);
                const uint32_t accum = ks > 0 ? 1u : 0u;
                if (p.amode == 2) {
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
                        ::"r"(tb), "r"(ta + 8u * ks), "l"(db), "r"(idesc), "r"(accum) : "memory");
                    continue;
                }
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                    ::"r"(tb), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
        }
        // everybody waits for the MMAs
        {
            uint32_t done = 0;
            while (!done) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(done) : "r"(smem_u32(&mbar)), "r"(phase) : "memory");
            }
            phase ^= 1u;
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 0 && rep == p.reps - 1) t1 = clock64();
    }
    // read back: thread (warp w, lane l) owns TMEM lane 32 w + l
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        const uint32_t taddr = tb + ((uint32_t)(32 * wid) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int q = 0; q < 16; q++) D[(32 * wid + lane) * N + c0 + q] = (int)v[q];
    }
    if (tid == 0) {
        clk[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb));
}

int main(int argc, char** argv) {
    Params p;
    p.amode = argc > 1 ? atoi(argv[1]) : 0;
    p.bmode = argc > 2 ? atoi(argv[2]) : 0;
    p.N = argc > 3 ? atoi(argv[3]) : 96;
    p.reps = argc > 4 ? atoi(argv[4]) : 1;
    if (p.N % 16 || p.N < 16 || p.N > 256) return 3;
    const int N = p.N;
    std::vector<uint8_t> A(128 * 128), B(128 * N);
    uint32_t s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return s >> 24; };
    for (auto& x : A) x = (uint8_t)(rnd() & 3);
    for (auto& x : B) x = (uint8_t)rnd();
    std::vector<int> ref(128 * N, 0);
    for (int m = 0; m < 128; m++)
        for (int k = 0; k < 128; k++) {
            const int a = A[m * 128 + k];
            if (a)
                for (int n = 0; n < N; n++) ref[m * N + n] += a * (int)B[k * N + n];
        }
    uint8_t *dA, *dB;
    int* dD;
    long long* dclk;
    CK(cudaMalloc(&dA, A.size()));
    CK(cudaMalloc(&dB, B.size()));
    CK(cudaMalloc(&dD, sizeof(int) * 128 * N));
    CK(cudaMalloc(&dclk, 16));
    CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xff, sizeof(int) * 128 * N));
    const int smem = 16384 + 32768 + 1024;
    CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_probe<<<1, 128, smem>>>(dA, dB, dD, dclk, p);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<int> D(128 * N);
    long long clk = 0;
    CK(cudaMemcpy(D.data(), dD, sizeof(int) * 128 * N, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&clk, dclk, 8, cudaMemcpyDeviceToHost));
    long long bad = 0;
    for (int i = 0; i < 128 * N; i++) bad += D[i] != ref[i];
    printf("amode %d bmode %d N %d reps %d: mismatches %lld / %d", p.amode, p.bmode, N, p.reps, bad, 128 * N);
    if (p.reps > 1) printf("   clocks per 4-MMA group (issue..mbarrier) %.1f", (double)clk / (p.reps - 1));
    printf("\n");
    if (bad) {
        for (int i = 0, shown = 0; i < 128 * N && shown < 6; i++)
            if (D[i] != ref[i]) {
                printf("  D[%d][%d] = %d, expected %d\n", i / N, i % N, D[i], ref[i]);
                shown++;
            }
    }
    return bad ? 1 : 0;
}

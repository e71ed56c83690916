// dist_mat.cu — the materialising consumer of the pair tiles as its own HBM-write-bound kernel.
//
// Reference: distance_matrix / _compute_distance! (src/eval_clustering.jl:164-254):
//     D[i,j] = sum_g w_g * sum_{f in g} |X[f,i] - X[f,j]|,   symmetric, zero diagonal, n x n column-major Float64.
// For few features (C1: F = 19, C2: F = 4) this stage is bound by writing 8 n^2 bytes.  The general tile kernel
// (pair_tiles.cuh, 64 x 64 tiles, 4 x 4 register blocks, tile staged through shared memory for the mirrored store) reached
// 0.36 of the HBM copy peak: its shared-memory pipe was 74 % busy.  This kernel keeps the arithmetic per pair IDENTICAL
// (same order of additions, same fma at the group ends — the on-the-fly DBSCAN, which uses pair_tiles.cuh, must see
// bit-identical distances) and changes everything around it:
//   * persistent CTAs (one per SM) walk the 128 x 64 tiles of the upper triangle;
//   * feature slabs (8 features x 128 rows, 8 x 64 columns) arrive by TMA (cp.async.bulk.tensor.2d, tensor maps over
//     X [F][n]; rows / features past the end are zero-filled by the hardware) into a 3-stage ring, completion on mbarriers;
//   * every thread owns an 8 x 4 block of pairs (64 FP64 accumulators): 6 128-bit shared-memory loads feed 64 FP64 adds;
//   * the block is stored straight from registers: 8 consecutive rows of a column are 64 contiguous bytes of D (direct
//     tile), 4 consecutive columns of a row are 32 contiguous bytes of the mirrored tile — streaming 16-byte stores, no
//     staging, no barrier; the stores of one tile drain while the next tile is being accumulated.
#include <cuda.h>

#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace mcbb {

constexpr int DM_R = 128;     // tile rows
constexpr int DM_C = 64;      // tile columns
constexpr int DM_FC = 8;      // features per slab
constexpr int DM_STAGES = 3;  // slabs in flight
constexpr int DM_TX = 16;      // thread grid 16 x 16: 8 rows x 4 columns per thread
constexpr unsigned DM_SLAB_BYTES = (DM_R + DM_C) * DM_FC * sizeof(double);

struct DistMatArgs {
    long long n;
    int F, n_slabs;
    int nbr, nbc;             // row blocks of 128, column blocks of 64
    long long n_tiles;        // tiles (br, bc) with bc >= 2 br
    const double* fw;         // [F] weight of the feature's group
    const unsigned short* fmask;  // [ceil(F/16)] bit f%16: feature f closes its group
    double* D;
};

__device__ __forceinline__ unsigned dm_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dm_mbar_wait(const unsigned mbar, const unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(mbar), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void dm_tma_load_2d(const unsigned dst, const CUtensorMap* tmap, const int x, const int y,
                                               const unsigned mbar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(tmap), "r"(x), "r"(y), "r"(mbar)
        : "memory");
}
__device__ __forceinline__ void dm_st2(double* p, const double a, const double b) {
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}

// tile index t -> (br, bc): rows br = 0, 1, ... hold nbc - 2 br tiles each (bc = 2 br .. nbc - 1)
__device__ __forceinline__ void dm_tile_of(const long long t, const int nbc, int* br, int* bc) {
    // cumulative count C(b) = b nbc - b (b - 1); largest b with C(b) <= t
    const double B = (double)nbc + 1.;
    int b = (int)((B - sqrt(fmax(B * B - 4. * (double)t, 0.))) * 0.5);
    b = max(b, 0);
    while ((long long)b * nbc - (long long)b * (b - 1) > t) b--;
    while ((long long)(b + 1) * nbc - (long long)(b + 1) * b <= t) b++;
    *br = b;
    *bc = 2 * b + (int)(t - ((long long)b * nbc - (long long)b * (b - 1)));
}

__global__ void __launch_bounds__(256, 1) k_dist_mat_tma(const DistMatArgs a, const __grid_constant__ CUtensorMap tm_rows,
                                                         const __grid_constant__ CUtensorMap tm_cols) {
    __shared__ __align__(128) double s_slab[DM_STAGES][(DM_R + DM_C) * DM_FC];  // [stage][8 x 128 rows | 8 x 64 columns]
    __shared__ __align__(8) unsigned long long s_full[DM_STAGES];
    const int tid = threadIdx.x;
    const int tx = tid & (DM_TX - 1), ty = tid >> 4;
    if (tid == 0) {
        for (int s = 0; s < DM_STAGES; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dm_smem_u32(&s_full[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long my_tiles = (a.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;  // tiles t = blockIdx.x + k gridDim.x
    const long long total_slabs = my_tiles * a.n_slabs;

    // producer (thread 0): slab number q of this CTA's stream -> (tile, feature offset)
    auto issue = [&](const long long q) {
        const long long k = q / a.n_slabs;
        const int sl = (int)(q - k * a.n_slabs);
        int br, bc;
        dm_tile_of(blockIdx.x + k * gridDim.x, a.nbc, &br, &bc);
        const int stage = (int)(q % DM_STAGES);
        const unsigned mbar = dm_smem_u32(&s_full[stage]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(DM_SLAB_BYTES) : "memory");
        dm_tma_load_2d(dm_smem_u32(&s_slab[stage][0]), &tm_rows, br * DM_R, sl * DM_FC, mbar);
        dm_tma_load_2d(dm_smem_u32(&s_slab[stage][DM_R * DM_FC]), &tm_cols, bc * DM_C, sl * DM_FC, mbar);
    };
    if (tid == 0)
        for (long long q = 0; q < DM_STAGES - 1 && q < total_slabs; q++) issue(q);

    long long q = 0;
    for (long long k = 0; k < my_tiles; k++) {
        int br, bc;
        dm_tile_of(blockIdx.x + k * gridDim.x, a.nbc, &br, &bc);
        const long long tr = (long long)br * DM_R, tc = (long long)bc * DM_C;
        double acc[8][4], accg[8][4];
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j] = accg[i][j] = 0.;
        for (int sl = 0; sl < a.n_slabs; sl++, q++) {
            __syncthreads();  // everyone is done with the stage that is refilled next
            if (tid == 0 && q + DM_STAGES - 1 < total_slabs) issue(q + DM_STAGES - 1);
            const int stage = (int)(q % DM_STAGES);
            dm_mbar_wait(dm_smem_u32(&s_full[stage]), (unsigned)((q / DM_STAGES) & 1));
            const double* s_r = &s_slab[stage][0];
            const double* s_c = &s_slab[stage][DM_R * DM_FC];
            const int f0 = sl * DM_FC;
            const unsigned emask = (unsigned)a.fmask[f0 >> 4] >> (f0 & 15);
            const int fc = min(DM_FC, a.F - f0);
#pragma unroll 2
            for (int f = 0; f < fc; f++) {
                double xr[8], xc[4];
                const double2* pr = reinterpret_cast<const double2*>(s_r + f * DM_R + tx * 8);
                const double2* pc = reinterpret_cast<const double2*>(s_c + f * DM_C + ty * 4);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const double2 t = pr[i];
                    xr[2 * i] = t.x;
                    xr[2 * i + 1] = t.y;
                }
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const double2 t = pc[j];
                    xc[2 * j] = t.x;
                    xc[2 * j + 1] = t.y;
                }
#pragma unroll
                for (int i = 0; i < 8; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) accg[i][j] += fabs(xr[i] - xc[j]);
                if ((emask >> f) & 1u) {  // group closes: mat_elements += weight * distance_func(...)
                    const double w = __ldg(a.fw + f0 + f);
#pragma unroll
                    for (int i = 0; i < 8; i++)
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            acc[i][j] = fma(w, accg[i][j], acc[i][j]);
                            accg[i][j] = 0.;
                        }
                }
            }
        }
        // ---- store the 8 x 4 block: direct D[ri, cj] and mirrored D[cj, ri] ------------------------------------------
        const long long ri0 = tr + tx * 8, cj0 = tc + ty * 4;
        const bool interior = (tr + DM_R <= a.n) && (tc + DM_C <= a.n) && (tc >= tr + DM_R);
        if (interior) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                double* p = a.D + ri0 + a.n * (cj0 + j);
#pragma unroll
                for (int i = 0; i < 8; i += 2) dm_st2(p + i, acc[i][j], acc[i + 1][j]);
            }
#pragma unroll
            for (int i = 0; i < 8; i++) {
                double* p = a.D + cj0 + a.n * (ri0 + i);
                dm_st2(p, acc[i][0], acc[i][1]);
                dm_st2(p + 2, acc[i][2], acc[i][3]);
            }
        } else {  // tiles on the diagonal or at the edge of the matrix: element by element
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const long long ri = ri0 + i, cj = cj0 + j;
                    if (ri >= a.n || cj >= a.n || cj < ri) continue;  // below the diagonal: written by another tile's mirror
                    if (cj == ri) {
                        __stcs(a.D + ri + a.n * cj, 0.);
                    } else {
                        __stcs(a.D + ri + a.n * cj, acc[i][j]);
                        __stcs(a.D + cj + a.n * ri, acc[i][j]);
                    }
                }
        }
    }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled encode_tiled_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// Returns handled = false (and launches nothing) when the shapes do not meet the TMA requirements: the caller then uses
// the general tile kernel.  X [F][n] on the device, fw / fmask from upload_groups.
int distance_matrix_tma(const double* X, long long n, int F, const double* fw, const unsigned short* fmask, double* D,
                        cudaStream_t st, bool* handled) {
    *handled = false;
    // Measured (tools/bench_distmat.py, n = 20 000): 1.93 ms against 1.40 ms of the staged tile kernel at F = 19, 1.16 ms
    // against 0.83 ms at F = 4 — storing straight from registers leaves every store instruction with 32 scattered half
    // sectors, and that costs more than the shared-memory staging it saves.  Kept as an opt-in experiment (test hook
    // "dist_tma" = 1) with its bitwise test; the default path is the staged tile kernel.
    if (!tune_get("dist_tma", 0)) return 0;
    if (n < 2 * DM_R || (n & 1) || n > (1LL << 31) - 1) return 0;  // TMA strides are multiples of 16 bytes: n even
    if ((reinterpret_cast<size_t>(X) & 15) || (reinterpret_cast<size_t>(D) & 15)) return 0;
    PFN_encodeTiled enc = encode_tiled_fn();
    if (!enc) return 0;
    alignas(64) CUtensorMap tm_rows, tm_cols;
    const cuuint64_t gdim[2] = {(cuuint64_t)n, (cuuint64_t)F};
    const cuuint64_t gstr[1] = {(cuuint64_t)n * sizeof(double)};
    const cuuint32_t estr[2] = {1, 1};
    const cuuint32_t box_r[2] = {DM_R, DM_FC}, box_c[2] = {DM_C, DM_FC};
    if (enc(&tm_rows, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)X, gdim, gstr, box_r, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return 0;
    if (enc(&tm_cols, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)X, gdim, gstr, box_c, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return 0;
    DistMatArgs a = {};
    a.n = n;
    a.F = F;
    a.n_slabs = (F + DM_FC - 1) / DM_FC;
    a.nbr = (int)((n + DM_R - 1) / DM_R);
    a.nbc = (int)((n + DM_C - 1) / DM_C);
    long long tiles = 0;
    for (int b = 0; b < a.nbr; b++) tiles += std::max(0, a.nbc - 2 * b);
    a.n_tiles = tiles;
    a.fw = fw;
    a.fmask = fmask;
    a.D = D;
    int dev = 0, sms = 148;
    MCBB_CUDA_OK(cudaGetDevice(&dev));
    MCBB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = (int)std::min<long long>(tiles, (long long)sms);
    k_dist_mat_tma<<<grid, 256, 0, st>>>(a, tm_rows, tm_cols);
    MCBB_LAUNCHED();
    *handled = true;
    return 0;
}

}  // namespace mcbb

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
//     X [F][n]; rows / features past the end are zero-filled by the hardware) into a ring of stages, issued by a dedicated
//     producer warp; full / empty mbarriers per stage, no CTA barrier in the accumulation loop;
//   * every thread owns an 8 x 4 block of pairs (64 FP64 accumulators; rows i*16 + tx, so that the lanes of a warp hold
//     consecutive rows): 8 + 2 shared-memory loads feed 64 FP64 adds;
//   * the finished tile is staged ONCE in shared memory in both orientations — [column][row] for D[tile] and, 128-byte
//     swizzled, [row][column] for the mirrored tile — and leaves through TMA stores (cp.async.bulk.tensor, UTMASTG):
//     whole 128-byte lines, no LSU traffic, clipped at the matrix edge by the hardware; the stores of one tile drain
//     while the next tile is being accumulated (the staging is reused only after cp.async.bulk.wait_group.read).
//     d(i,j) is bitwise symmetric, so a tile that straddles the diagonal simply writes all its elements (some twice, with
//     identical bits); only the diagonal itself is forced to zero.
//
// Measured history at n = 20 000 (profiles/README.md): stores straight from registers, no staging: 1.93 ms at F = 19 (every
// store instruction of the mirrored orientation touched 32 scattered half sectors); staged TMA stores, thread 0 of the
// consumers issuing the loads behind a CTA barrier per slab: 1.04 ms; a dedicated producer warp with full / empty
// barriers (the tile-index arithmetic no longer sits in front of the consumers), operands of feature f + 1 requested
// before the adds of feature f, group weights in shared memory, 64-row tiles with two CTAs per SM: 0.92 ms (0.53 of the
// HBM copy peak; F = 12 → 0.75, F <= 8 → 0.83-0.92).  ncu at F = 64: FP64 pipe 65 % active — the bound above F ~ 10.
#include <cuda.h>

#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace mcbb {

// Tile rows R = 128 (one CTA of 256 threads per SM) or 64 (128 threads, TWO CTAs per SM: one CTA's staging / store epilogue
// overlaps the other's accumulation).
constexpr int DM_C = 64;      // tile columns
constexpr int DM_FC = 8;      // features per slab
constexpr int DM_FMAX = 256;  // largest feature count of this kernel: weights / group masks live in shared memory
// slabs in flight: the slab loads of the NEXT tile must be queued before this tile's TMA stores (the TMA unit works in
// order), or the next tile stalls behind the stores
template <int R> struct DmCfg {
    static constexpr int STAGES = (R == 128) ? 6 : 4;
    static constexpr unsigned SLAB_BYTES = (R + DM_C) * DM_FC * sizeof(double);
    static constexpr int STAGE_A = R * DM_C;   // doubles: [64 columns][R rows]
    static constexpr int STAGE_B = R * DM_C;   // doubles: 4 boxes of [R rows][16 columns], 128-byte swizzle
    static constexpr size_t SMEM = 1024 /* alignment slack */ + sizeof(double) * ((size_t)STAGES * (R + DM_C) * DM_FC + STAGE_A + STAGE_B);
};

struct DistMatArgs {
    long long n;
    int F, n_slabs;
    int nbr, nbc;             // row blocks of R, column blocks of 64
    long long n_tiles;        // tiles (br, bc) with bc >= (R / 64) br
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
__device__ __forceinline__ void dm_tma_store_2d(const CUtensorMap* tmap, const int x, const int y, const unsigned src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tmap), "r"(x), "r"(y), "r"(src)
                 : "memory");
}

// tile index t -> (br, bc): rows br = 0, 1, ... hold nbc - M br tiles each (bc = M br .. nbc - 1), M = R / 64
template <int M>
__device__ __forceinline__ void dm_tile_of(const long long t, const int nbc, int* br, int* bc) {
    // cumulative count C(b) = b nbc - M b (b - 1) / 2; largest b with C(b) <= t
    auto C = [&](const long long b) { return b * nbc - (long long)M * b * (b - 1) / 2; };
    const double B = (double)nbc + 0.5 * M;
    int b = (int)((B - sqrt(fmax(B * B - 2. * M * (double)t, 0.))) / (double)M);
    b = max(b, 0);
    while (C(b) > t) b--;
    while (C(b + 1) <= t) b++;
    *br = b;
    *bc = M * b + (int)(t - C(b));
}

// RPT = rows per thread: 8 -> 16 x 16 = 256 threads (8 x 4 pairs each), 4 -> 32 x 16 = 512 threads (4 x 4 pairs each: twice
// the warps per scheduler to hide the shared-memory and FP64 latencies, at 1.6 x the shared-memory reads per pair)
template <int RPT, int DM_R>
__global__ void __launch_bounds__(DM_R / RPT * 16 + 32, 128 / DM_R) k_dist_mat_tma(const DistMatArgs a, const __grid_constant__ CUtensorMap tm_rows,
                                                         const __grid_constant__ CUtensorMap tm_cols,
                                                         const __grid_constant__ CUtensorMap tm_out,
                                                         const __grid_constant__ CUtensorMap tm_out_t) {
    constexpr int DM_STAGES = DmCfg<DM_R>::STAGES, DM_STAGE_A = DmCfg<DM_R>::STAGE_A, DM_STAGE_B = DmCfg<DM_R>::STAGE_B;
    constexpr unsigned DM_SLAB_BYTES = DmCfg<DM_R>::SLAB_BYTES;
    extern __shared__ unsigned char dm_raw[];
    __shared__ __align__(8) unsigned long long s_full[DM_STAGES], s_empty[DM_STAGES];
    // group weights and group-end masks: read once — a global load per slab in front of the first add stalled every warp
    // on the L2 (ncu: long-scoreboard 0.97 stalls per issue at F = 64)
    __shared__ double s_fw[DM_FMAX];
    __shared__ unsigned short s_fmask[DM_FMAX / 16];
    // 1024-byte aligned carve-up: the swizzled staging boxes need it
    unsigned char* base = dm_raw + ((1024u - (dm_smem_u32(dm_raw) & 1023u)) & 1023u);
    double* s_stB = reinterpret_cast<double*>(base);                      // 4 swizzled boxes
    double* s_stA = s_stB + DM_STAGE_B;
    double (*s_slab)[(DM_R + DM_C) * DM_FC] = reinterpret_cast<double (*)[(DM_R + DM_C) * DM_FC]>(s_stA + DM_STAGE_A);
    constexpr int TXN = DM_R / RPT;   // threads along the rows
    constexpr int NTH = TXN * 16;     // consumer threads; one more warp produces (issues the TMA slab loads)
    constexpr int NCW = NTH / 32;
    const int tid = threadIdx.x;
    const int tx = tid % TXN, ty = tid / TXN;
    for (int f = tid; f < a.F; f += blockDim.x) s_fw[f] = a.fw[f];  // F <= DM_FMAX (checked by the launcher)
    for (int f = tid; f < (a.F + 15) / 16; f += blockDim.x) s_fmask[f] = a.fmask[f];
    if (tid == 0) {
        for (int s = 0; s < DM_STAGES; s++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dm_smem_u32(&s_full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dm_smem_u32(&s_empty[s])), "n"(NCW));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long my_tiles = (a.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;  // tiles t = blockIdx.x + k gridDim.x

    if (tid >= NTH) {
        // ---- producer warp: keeps DM_STAGES slabs in flight, never synchronises with the consumers except through the
        // full / empty barriers of the ring (no CTA barrier in the accumulation loop; the tile-index arithmetic — a square
        // root and 64-bit products — runs beside the consumers instead of in front of them)
        if (tid == NTH) {
            int stage = 0;
            unsigned round = 0;
            for (long long k = 0; k < my_tiles; k++) {
                int br, bc;
                dm_tile_of<DM_R / 64>(blockIdx.x + k * gridDim.x, a.nbc, &br, &bc);
                for (int sl = 0; sl < a.n_slabs; sl++) {
                    if (round > 0) dm_mbar_wait(dm_smem_u32(&s_empty[stage]), (round - 1) & 1u);
                    const unsigned mbar = dm_smem_u32(&s_full[stage]);
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(DM_SLAB_BYTES) : "memory");
                    dm_tma_load_2d(dm_smem_u32(&s_slab[stage][0]), &tm_rows, br * DM_R, sl * DM_FC, mbar);
                    dm_tma_load_2d(dm_smem_u32(&s_slab[stage][DM_R * DM_FC]), &tm_cols, bc * DM_C, sl * DM_FC, mbar);
                    if (++stage == DM_STAGES) {
                        stage = 0;
                        round++;
                    }
                }
            }
        }
        return;
    }

    int stage = 0;
    unsigned phase = 0;
    for (long long k = 0; k < my_tiles; k++) {
        int br, bc;
        dm_tile_of<DM_R / 64>(blockIdx.x + k * gridDim.x, a.nbc, &br, &bc);
        const long long tr = (long long)br * DM_R, tc = (long long)bc * DM_C;
        double acc[RPT][4], accg[RPT][4];
#pragma unroll
        for (int i = 0; i < RPT; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j] = accg[i][j] = 0.;
        for (int sl = 0; sl < a.n_slabs; sl++) {
            dm_mbar_wait(dm_smem_u32(&s_full[stage]), phase);
            const double* s_r = &s_slab[stage][0];
            const double* s_c = &s_slab[stage][DM_R * DM_FC];
            const int f0 = sl * DM_FC;
            const unsigned emask = (unsigned)s_fmask[f0 >> 4] >> (f0 & 15);
            const int fc = min(DM_FC, a.F - f0);
            auto load = [&](const int f, double (&xr)[RPT], double (&xc)[4]) {
                const double* pr = s_r + f * DM_R + tx;  // rows i * TXN + tx: the lanes of a warp hold consecutive rows
                const double2* pc = reinterpret_cast<const double2*>(s_c + f * DM_C + ty * 4);
#pragma unroll
                for (int i = 0; i < RPT; i++) xr[i] = pr[i * TXN];
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const double2 t = pc[j];
                    xc[2 * j] = t.x;
                    xc[2 * j + 1] = t.y;
                }
            };
            auto compute = [&](const int f, const double (&xr)[RPT], const double (&xc)[4]) {
#pragma unroll
                for (int i = 0; i < RPT; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) accg[i][j] += fabs(xr[i] - xc[j]);
                if ((emask >> f) & 1u) {  // group closes: mat_elements += weight * distance_func(...)
                    const double w = s_fw[f0 + f];
#pragma unroll
                    for (int i = 0; i < RPT; i++)
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            acc[i][j] = fma(w, accg[i][j], acc[i][j]);
                            accg[i][j] = 0.;
                        }
                }
            };
            if (fc == DM_FC) {
                // a full slab, unrolled: the operands of feature f + 1 are requested before the 64 adds of feature f
                double xr[2][RPT], xc[2][4];
                load(0, xr[0], xc[0]);
#pragma unroll
                for (int f = 0; f < DM_FC; f++) {
                    if (f + 1 < DM_FC) load(f + 1, xr[(f + 1) & 1], xc[(f + 1) & 1]);
                    compute(f, xr[f & 1], xc[f & 1]);
                }
            } else {
#pragma unroll 2
                for (int f = 0; f < fc; f++) {
                    double xr[RPT], xc[4];
                    load(f, xr, xc);
                    compute(f, xr, xc);
                }
            }
            __syncwarp();  // the whole warp has read the slab: hand the stage back to the producer
            if ((tid & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(dm_smem_u32(&s_empty[stage])) : "memory");
            if (++stage == DM_STAGES) {
                stage = 0;
                phase ^= 1u;
            }
        }
        // ---- stage the 8 x 4 block in both orientations, then TMA stores -----------------------------------------------
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the previous tile has left the staging
        asm volatile("bar.sync 1, %0;" ::"n"(NTH) : "memory");
        const int cj0 = ty * 4;
#pragma unroll
        for (int i = 0; i < RPT; i++) {
            const int r = i * TXN + tx;
            double v[4];
#pragma unroll
            for (int j = 0; j < 4; j++) v[j] = (tr + r == tc + cj0 + j) ? 0. : acc[i][j];  // D[p,p] = 0 also for Inf / NaN features
#pragma unroll
            for (int j = 0; j < 4; j++) s_stA[(cj0 + j) * DM_R + r] = v[j];
            // mirrored: box ty / 4 of [R rows][16 columns]; 16-byte chunk index XOR (row mod 8) = the 128-byte swizzle
            double* rowB = s_stB + (ty >> 2) * (DM_R * 16) + r * 16;
            const int ch = (ty & 3) * 2;
            *reinterpret_cast<double2*>(rowB + ((ch ^ (r & 7)) << 1)) = make_double2(v[0], v[1]);
            *reinterpret_cast<double2*>(rowB + (((ch + 1) ^ (r & 7)) << 1)) = make_double2(v[2], v[3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the TMA engine
        asm volatile("bar.sync 1, %0;" ::"n"(NTH) : "memory");
        if (tid == 0) {
            dm_tma_store_2d(&tm_out, (int)tr, (int)tc, dm_smem_u32(s_stA));  // D[tr.., tc..]: box {R rows, 64 columns}
#pragma unroll
            for (int b = 0; b < 4; b++)                                     // D[tc + 16 b .., tr..]: box {16, R}
                dm_tma_store_2d(&tm_out_t, (int)tc + 16 * b, (int)tr, dm_smem_u32(s_stB + b * (DM_R * 16)));
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all stores complete before the CTA exits
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
    // default for shapes that meet the TMA requirements; test hook "dist_tma" = 0 forces the general tile kernel
    if (!tune_get("dist_tma", 1)) return 0;
    if (n < 256 || (n & 1) || n > (1LL << 31) - 1) return 0;  // TMA strides are multiples of 16 bytes: n even
    if (F > DM_FMAX) return 0;  // weights / group masks live in shared memory; wider feature matrices use the tile kernel
    const int DM_R = (tune_get("dist_tma_r", 64) == 128 || tune_get("dist_tma_rpt", 8) != 8) ? 128 : 64;
    if ((reinterpret_cast<size_t>(X) & 15) || (reinterpret_cast<size_t>(D) & 15)) return 0;
    PFN_encodeTiled enc = encode_tiled_fn();
    if (!enc) return 0;
    alignas(64) CUtensorMap tm_rows, tm_cols;
    const cuuint64_t gdim[2] = {(cuuint64_t)n, (cuuint64_t)F};
    const cuuint64_t gstr[1] = {(cuuint64_t)n * sizeof(double)};
    const cuuint32_t estr[2] = {1, 1};
    const cuuint32_t box_r[2] = {(cuuint32_t)DM_R, DM_FC}, box_c[2] = {DM_C, DM_FC};
    if (enc(&tm_rows, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)X, gdim, gstr, box_r, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return 0;
    if (enc(&tm_cols, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)X, gdim, gstr, box_c, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return 0;
    alignas(64) CUtensorMap tm_out, tm_out_t;
    const cuuint64_t odim[2] = {(cuuint64_t)n, (cuuint64_t)n};
    const cuuint32_t box_o[2] = {(cuuint32_t)DM_R, DM_C}, box_t[2] = {16, (cuuint32_t)DM_R};
    if (enc(&tm_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)D, odim, gstr, box_o, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return 0;
    if (enc(&tm_out_t, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)D, odim, gstr, box_t, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return 0;
    DistMatArgs a = {};
    a.n = n;
    a.F = F;
    a.n_slabs = (F + DM_FC - 1) / DM_FC;
    a.nbr = (int)((n + DM_R - 1) / DM_R);
    a.nbc = (int)((n + DM_C - 1) / DM_C);
    long long tiles = 0;
    for (int b = 0; b < a.nbr; b++) tiles += std::max(0, a.nbc - (DM_R / 64) * b);
    a.n_tiles = tiles;
    a.fw = fw;
    a.fmask = fmask;
    a.D = D;
    int dev = 0, sms = 148;
    MCBB_CUDA_OK(cudaGetDevice(&dev));
    MCBB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // measured at n = 20 000, F = 19: 128-row tiles, one CTA per SM: 1.04 ms (8 rows per thread) / 1.27 ms (4 rows, 512
    // threads); 64-row tiles, two CTAs per SM (default): see profiles/README.md
    if (DM_R == 64) {
        const int grid = (int)std::min<long long>(tiles, 2LL * sms);
        MCBB_CUDA_OK(cudaFuncSetAttribute(k_dist_mat_tma<8, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DmCfg<64>::SMEM));
        k_dist_mat_tma<8, 64><<<grid, 128 + 32, DmCfg<64>::SMEM, st>>>(a, tm_rows, tm_cols, tm_out, tm_out_t);
    } else if (tune_get("dist_tma_rpt", 8) == 8) {
        const int grid = (int)std::min<long long>(tiles, (long long)sms);
        MCBB_CUDA_OK(cudaFuncSetAttribute(k_dist_mat_tma<8, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DmCfg<128>::SMEM));
        k_dist_mat_tma<8, 128><<<grid, 256 + 32, DmCfg<128>::SMEM, st>>>(a, tm_rows, tm_cols, tm_out, tm_out_t);
    } else {
        const int grid = (int)std::min<long long>(tiles, (long long)sms);
        MCBB_CUDA_OK(cudaFuncSetAttribute(k_dist_mat_tma<4, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DmCfg<128>::SMEM));
        k_dist_mat_tma<4, 128><<<grid, 512 + 32, DmCfg<128>::SMEM, st>>>(a, tm_rows, tm_cols, tm_out, tm_out_t);
    }
    MCBB_LAUNCHED();
    *handled = true;
    return 0;
}

}  // namespace mcbb

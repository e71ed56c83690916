// multi.cuh — interface between the per-rank stages (cluster.cu) and the in-process multi-device driver (multi.cu).
#pragma once
#include "common.cuh"

namespace mcbb {

// collectives of one rank over int32 arrays, provided by multi.cu (NCCL); op: 0 = sum, 1 = min
struct MultiComm {
    int rank, world;
    int (*allreduce_i32)(void* ctx, int* buf, long long count, int op, cudaStream_t st);
    int (*allgather_i32)(void* ctx, const int* send, int* recv, long long count_per_rank, cudaStream_t st);
    void* ctx;
};

int dbscan_features_multi_rank(const double* X, long long n, int n_groups, const int32_t* group_len,
                               const double* group_weight, double eps, int minpts, const MultiComm& mc, long long* assignments,
                               long long* seeds, long long* counts, long long* n_clusters_host, cudaStream_t st);

}  // namespace mcbb

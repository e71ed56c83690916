// batched.cu — block-cooperative ensemble integrator for large networks (placeholder until the batched
// kernel lands: reports "not handled" so the one-trajectory-per-lane kernel is used).
#include "common.cuh"

namespace mcbb {
int launch_batched_kuramoto(const SysDev&, const SolverDev&, const FeatDev&, const SolveIO&, cudaStream_t,
                            bool* handled) {
    *handled = false;
    return 0;
}
}  // namespace mcbb

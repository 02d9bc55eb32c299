// f32 instantiation of the persistent solve kernels, zero-padded operator (valid-window masks).
#include "bsgp_solve_kernel.cuh"

namespace bsgp {
template cudaError_t launch_solve<float, true>(const LaunchCfg&, const SolveArgs<float>&, const SmemPlan&, size_t);
template cudaError_t query_solve_clusters<float, true>(const LaunchCfg&, int, int*);
template cudaError_t launch_frame<float, true>(const LaunchCfg&, const SolveArgs<float>&, const SmemPlan&, size_t, double*);
template cudaError_t query_frame_ctas<float, true>(const LaunchCfg&, int*);
}  // namespace bsgp

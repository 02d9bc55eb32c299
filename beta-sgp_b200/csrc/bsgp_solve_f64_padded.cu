// f64 instantiation of the persistent solve kernels, zero-padded operator (valid-window masks).
#include "bsgp_solve_kernel.cuh"

namespace bsgp {
template cudaError_t launch_solve<double, true>(const LaunchCfg&, const SolveArgs<double>&, const SmemPlan&, size_t);
template cudaError_t query_solve_clusters<double, true>(const LaunchCfg&, int, int*);
template cudaError_t launch_frame<double, true>(const LaunchCfg&, const SolveArgs<double>&, const SmemPlan&, size_t, double*);
template cudaError_t query_frame_ctas<double, true>(const LaunchCfg&, int*);
}  // namespace bsgp

// fp64 instantiation of the persistent solve kernel (the reference's precision).
#include "bsgp_solve_kernel.cuh"

namespace bsgp {
template cudaError_t launch_solve<double>(const LaunchCfg&, const SolveArgs<double>&, const SmemPlan&, size_t);
template cudaError_t query_solve_clusters<double>(const LaunchCfg&, int, int*);
template cudaError_t launch_frame<double>(const LaunchCfg&, const SolveArgs<double>&, const SmemPlan&, size_t, double*);
template cudaError_t query_frame_ctas<double>(const LaunchCfg&, int*);
}  // namespace bsgp

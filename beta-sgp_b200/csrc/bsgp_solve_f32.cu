// fp32 instantiation of the persistent solve kernel (optional reduced-precision mode).
#include "bsgp_solve_kernel.cuh"

namespace bsgp {
template cudaError_t launch_solve<float>(const LaunchCfg&, const SolveArgs<float>&, const SmemPlan&, size_t);
template cudaError_t query_solve_clusters<float>(const LaunchCfg&, int, int*);
template cudaError_t launch_frame<float>(const LaunchCfg&, const SolveArgs<float>&, const SmemPlan&, size_t, double*);
template cudaError_t query_frame_ctas<float>(const LaunchCfg&, int*);
}  // namespace bsgp

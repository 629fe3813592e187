// rollout engine kernels: format fp16, with the fused actor
#include "tc_rollout_launch.cuh"
namespace bd {
namespace tc {
int launch_rollout_f0a(int act, bool prof, unsigned grid, const RolloutArgs& ra, cudaStream_t s) {
  return launch_rollout_a<0, true>(act, prof, grid, ra, s);
}
}  // namespace tc
}  // namespace bd

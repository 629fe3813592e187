// Launch code of the rollout engine kernel, instantiated per (format, actor) in tc_rollout_f*.cu
// so the template variants compile in parallel.
#pragma once
#include "api_internal.h"
#include "tc_engine.cuh"

namespace bd {
namespace tc {

template <int FMT, int ACT, bool WITH_ACTOR>
static int launch_rollout_t(bool prof, unsigned grid, const RolloutArgs& ra, cudaStream_t s) {
  // `grid` counts row tiles (clusters); column-split mode launches nranks CTAs per tile
  const unsigned R = ra.nranks > 1 ? (unsigned)ra.nranks : 1u;
  auto go = [&](auto kernel) -> int {
    set_smem_attr(kernel, ra.sm.total);
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(kThreads2);
    cfg.dynamicSmemBytes = ra.sm.total;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = R; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = R > 1 ? 1 : 0;
    unsigned clusters = grid;
    if (R > 1) {
      // co-resident clusters of this (kernel, cluster size, shared memory): queried once per variant
      static thread_local int cached_maxc[5] = {0, 0, 0, 0, 0};
      static thread_local uint32_t cached_smem[5] = {0, 0, 0, 0, 0};
      int maxc = cached_maxc[R];
      if (maxc == 0 || cached_smem[R] != ra.sm.total) {
        cfg.gridDim = dim3(R);
        if (cudaOccupancyMaxActiveClusters(&maxc, kernel, &cfg) != cudaSuccess || maxc < 1) {
          cudaGetLastError();
          BD_FAIL(BD_ERR_CUDA, "tensor-core rollout: a cluster of %u CTAs cannot be scheduled", R);
        }
        cached_maxc[R] = maxc; cached_smem[R] = ra.sm.total;
      }
      if (clusters > (unsigned)maxc) clusters = (unsigned)maxc;
    }
    cfg.gridDim = dim3(clusters * R);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, ra);
    if (e != cudaSuccess) BD_FAIL(BD_ERR_CUDA, "tensor-core rollout launch: %s", cudaGetErrorString(e));
    return BD_OK;
  };
  // (debug counters are compiled for fp16 ELU / ReLU only)
  constexpr bool kProfBuilt = FMT == 0 && (ACT == BD_ACT_ELU || ACT == BD_ACT_RELU);
  const bool p = prof && kProfBuilt;
  if (R > 1) {
    if (p) BD_TRY(go(rollout_fwd_kernel<FMT, ACT, WITH_ACTOR, kProfBuilt, true>));
    else BD_TRY(go(rollout_fwd_kernel<FMT, ACT, WITH_ACTOR, false, true>));
  } else {
    if (p) BD_TRY(go(rollout_fwd_kernel<FMT, ACT, WITH_ACTOR, kProfBuilt, false>));
    else BD_TRY(go(rollout_fwd_kernel<FMT, ACT, WITH_ACTOR, false, false>));
  }
  BD_CUDA_LAUNCH_CHECK();
  return BD_OK;
}
template <int FMT, bool WITH_ACTOR>
static int launch_rollout_a(int act, bool prof, unsigned grid, const RolloutArgs& ra, cudaStream_t s) {
  switch (act) {
    case BD_ACT_ELU: return launch_rollout_t<FMT, BD_ACT_ELU, WITH_ACTOR>(prof, grid, ra, s);
    case BD_ACT_RELU: return launch_rollout_t<FMT, BD_ACT_RELU, WITH_ACTOR>(prof, grid, ra, s);
    case BD_ACT_TANH: return launch_rollout_t<FMT, BD_ACT_TANH, WITH_ACTOR>(prof, grid, ra, s);
    default: return launch_rollout_t<FMT, BD_ACT_IDENTITY, WITH_ACTOR>(prof, grid, ra, s);
  }
}

}  // namespace tc
}  // namespace bd

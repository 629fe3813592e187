// Launch code of the rollout engine kernel, instantiated per (format, actor) in tc_rollout_f*.cu
// so the template variants compile in parallel.
#pragma once
#include "api_internal.h"
#include "tc_engine.cuh"

namespace bd {
namespace tc {

// Weight-share cluster size (R == 1 only): BD_TC_WS=2|4 makes clusters of 2 or 4 CTAs with their own row
// tiles read every weight stage from L2 once (producer_role).  OFF by default: measured on B200 at 148 and
// 296 row tiles, ws = 2 changes nothing (1003.5 vs 1008.5 us, 1666.7 vs 1659.0 us; outputs bit-identical)
// and ws = 4 is 40-50 % slower (fewer co-resident clusters) -- the issuer's weight waits are the ring's
// depth against the L2 latency, not L2 bandwidth.  Needs tile counts that divide evenly, so that the CTAs
// of a cluster run the same number of tiles (they refill each other's rings: lockstep).
static inline unsigned pick_weight_share(long long ntiles, unsigned grid) {
  unsigned want = 1u;
  if (const char* e = getenv("BD_TC_WS")) {
    const int v = atoi(e);
    if (v == 1 || v == 2 || v == 4) want = (unsigned)v;
  }
  while (want > 1 && (ntiles % want != 0 || grid < want || grid % want != 0)) want >>= 1;
  return want;
}

template <int FMT, int ACT, bool WITH_ACTOR>
static int launch_rollout_t(bool prof, unsigned grid, const RolloutArgs& ra_in, cudaStream_t s) {
  // `grid` counts row tiles (clusters); column-split mode launches nranks CTAs per tile
  RolloutArgs ra = ra_in;
  const unsigned R = ra.nranks > 1 ? (unsigned)ra.nranks : 1u;
  const long long ntiles = (ra.N + kTileRows - 1) / kTileRows;
  const unsigned WS = 1u;      // (pick_weight_share: compiled out of the kernels, see rollout_fwd_kernel)
  (void)ntiles;
  const unsigned CS = R > 1 ? R : WS;             // CTAs per cluster
  ra.ws = (int)WS;
  auto go = [&](auto kernel) -> int {
    set_smem_attr(kernel, ra.sm.total);
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(kThreads2);
    cfg.dynamicSmemBytes = ra.sm.total;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = CS > 1 ? 1 : 0;
    unsigned clusters = R > 1 ? grid : grid / WS;
    if (CS > 1) {
      // co-resident clusters of this (kernel, cluster size, shared memory): queried once per variant
      static thread_local int cached_maxc[5] = {0, 0, 0, 0, 0};
      static thread_local uint32_t cached_smem[5] = {0, 0, 0, 0, 0};
      int maxc = cached_maxc[CS];
      if (maxc == 0 || cached_smem[CS] != ra.sm.total) {
        cfg.gridDim = dim3(CS);
        if (cudaOccupancyMaxActiveClusters(&maxc, kernel, &cfg) != cudaSuccess || maxc < 1) {
          cudaGetLastError();
          BD_FAIL(BD_ERR_CUDA, "tensor-core rollout: a cluster of %u CTAs cannot be scheduled", CS);
        }
        cached_maxc[CS] = maxc; cached_smem[CS] = ra.sm.total;
      }
      if (clusters > (unsigned)maxc) clusters = (unsigned)maxc;
    }
    cfg.gridDim = dim3(clusters * CS);
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

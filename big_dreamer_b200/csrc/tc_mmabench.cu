// Micro-benchmark (debug entry, not part of the product path): cycles per tcgen05.mma for a chain of
// MMAs accumulating into one TMEM tile, for the no-swizzle KM8 operand layout and for SWIZZLE_128B.
// Operand contents are irrelevant (smem is zero-filled); only fetch/issue rates are measured.
#include "api_internal.h"
#include "tc_common.cuh"

namespace bd {
namespace tc {

__global__ void __launch_bounds__(128, 1) mma_bench_kernel(int N, int nmma, int layout, int dep, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t holder;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 + 256) * 256 * 2 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc<512>(&holder);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = holder;
  if (warp == 0) {
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 128 * 256 * 2);
    const uint32_t idesc = make_idesc_f16(0, 128, N);
    uint64_t ad, bdsc;
    uint32_t a_step, b_step;   // descriptor increment per K=16 step (in 16-byte units)
    if (layout == 0) {         // KM8 no-swizzle: LBO = rows*16, SBO = 128; K step = 2 groups
      ad = make_smem_desc(a_addr, 128 * 16, 128);
      bdsc = make_smem_desc(b_addr, N * 16, 128);
      a_step = (2 * 128 * 16) >> 4; b_step = (2 * N * 16) >> 4;
    } else {                   // SWIZZLE_128B K-major: rows of 128 B, 8-row groups 1024 B apart
      ad = make_smem_desc(a_addr, 16, 1024) | ((uint64_t)2 << 61);
      bdsc = make_smem_desc(b_addr, 16, 1024) | ((uint64_t)2 << 61);
      a_step = 32 >> 4; b_step = 32 >> 4;      // 16 elements = 32 bytes inside the 128-byte row
    }
    const long long t0 = clock64();
    if (elect_one()) {
      for (int i = 0; i < nmma; ++i) {
        const int ks = i & 3;                  // walk 4 K-steps (one 64-element swizzle atom)
        const uint32_t d = dep ? tmem : tmem + (uint32_t)((i & 1) * 256);
        umma_f16(d, ad + (uint64_t)(ks * a_step), bdsc + (uint64_t)(ks * b_step), idesc, i > 1 ? 1u : 0u);
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (tid == 0) out[0] = t1 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

}  // namespace tc
}  // namespace bd

extern "C" int bd_tc_mmabench(int N, int nmma, int layout, int dep, long long* out_cycles, bd_stream_t stream) {
  using namespace bd;
  BD_CHECK_ARG(N >= 16 && N <= 256 && (N % 16) == 0 && nmma > 0 && out_cycles, "bd_tc_mmabench: bad args");
  cudaFuncSetAttribute(tc::mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  tc::mma_bench_kernel<<<1, 128, (128 + 256) * 256 * 2, static_cast<cudaStream_t>(stream)>>>(N, nmma, layout, dep, out_cycles);
  BD_CUDA_LAUNCH_CHECK();
  return BD_OK;
}

// ---------------------------------------------------------------------------------------------
// Micro-benchmark: cost of exchanging an operand-tile slice between the CTAs of a cluster through
// distributed shared memory.  mode 0: per-thread st.shared::cluster.v4 + per-warp remote arrive;
// mode 1: one bulk async copy (smem -> peer smem) per peer, completing on the peer's mbarrier.
namespace bd {
namespace tc {
__global__ void __launch_bounds__(256, 1) dsmem_bench_kernel(int R, int bytes, int mode, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];
  const int tid = threadIdx.x, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  uint8_t* src = smem;                  // my slice
  uint8_t* dst = smem + 64 * 1024;      // where peers deposit theirs: [peer][bytes]
  for (int i = tid; i < bytes / 16; i += 256) reinterpret_cast<uint4*>(src)[i] = make_uint4(rank, i, 0, 0);
  if (tid == 0) {
    mbar_init(&bar[0], mode == 0 ? 8 * (R - 1) : 1);
    fence_barrier_init();
  }
  fence_proxy_async_smem();
  __syncthreads();
  cluster_sync_all();
  long long total = 0;
  for (int it = 0; it < iters; ++it) {
    const long long t0 = clock64();
    if (mode == 0) {
      for (int i = tid; i < bytes / 16; i += 256) {
        const uint4 u = reinterpret_cast<const uint4*>(src)[i];
        const uint32_t la = smem_u32(dst + (size_t)rank * bytes + (size_t)i * 16);
        for (uint32_t k = 0; k < (uint32_t)R; ++k)
          if (k != rank) st_cluster_v4(mapa_u32(la, k), u);
      }
      fence_proxy_async_all();
      __syncwarp();
      if (lane < R && (uint32_t)lane != rank) mbar_arrive_cluster(mapa_u32(smem_u32(&bar[0]), (uint32_t)lane));
    } else {
      if (tid == 0) {
        mbar_expect_tx(&bar[0], (uint32_t)(bytes * (R - 1)));
        for (uint32_t k = 0; k < (uint32_t)R; ++k) {
          if (k == rank) continue;
          const uint32_t rd = mapa_u32(smem_u32(dst + (size_t)rank * bytes), k);
          const uint32_t rb = mapa_u32(smem_u32(&bar[0]), k);
          asm volatile(
              "cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(rd),
              "r"(smem_u32(src)), "r"((uint32_t)bytes), "r"(rb)
              : "memory");
        }
      }
    }
    mbar_wait_cluster(&bar[0], it & 1);
    total += clock64() - t0;
    cluster_sync_all();
  }
  if (tid == 0 && blockIdx.x == 0) out[0] = total / iters;
  // consume so nothing is optimised away
  if (tid == 0 && reinterpret_cast<uint4*>(dst)[0].x == 0xdeadbeefu) out[1] = 1;
}
}  // namespace tc
}  // namespace bd

extern "C" int bd_tc_dsmembench(int R, int bytes, int mode, int iters, long long* out_cycles, bd_stream_t stream) {
  using namespace bd;
  BD_CHECK_ARG((R == 2 || R == 4) && bytes > 0 && bytes <= 32768 && (bytes % 16) == 0 && out_cycles, "bd_tc_dsmembench: bad args");
  const size_t sm = 64 * 1024 + (size_t)4 * 32768;
  cudaFuncSetAttribute(tc::dsmem_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(R); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = sm;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = R; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, tc::dsmem_bench_kernel, R, bytes, mode, iters, out_cycles);
  if (e != cudaSuccess) BD_FAIL(BD_ERR_CUDA, "dsmembench launch: %s", cudaGetErrorString(e));
  return BD_OK;
}

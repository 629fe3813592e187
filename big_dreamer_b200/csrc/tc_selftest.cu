// Self-test of the tcgen05 building block: one 128-row tile, Y = [X | 1] * [W | b]^T computed with
// tcgen05.mma (operands in shared memory in the KM8 layout, weights delivered by a bulk TMA copy
// of the packed image, accumulator in TMEM, read back with tcgen05.ld).  Exposed through the C ABI
// as bd_tc_selftest so the GPU test-suite can validate descriptors/layouts in isolation.
#include "api_internal.h"
#include "../../include/bd_b200_test.h"
#include "tc_common.cuh"
#include "tc_pack.cuh"

namespace bd {
namespace tc {

template <int FMT>
__global__ void __launch_bounds__(128, 1)
selftest_kernel(const float* __restrict__ x, int K, int Kp, int Np, const uint16_t* __restrict__ wp,
                float* __restrict__ y, int N, int swap_lbo_sbo) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* a_tile = smem;                                   // 128 x Kp
  uint8_t* b_tile = smem + 128 * Kp * 2;                    // Np x Kp
  __shared__ uint64_t bar_w, bar_mma;
  __shared__ uint32_t tmem_base_holder;
  const int tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) {
    mbar_init(&bar_w, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(&tmem_base_holder);
  // A tile: row = tid; fp32 -> 16-bit, constant 1 in column K (bias column), zero padding after
  for (int k0 = 0; k0 < Kp; k0 += 8) {
    uint32_t pk[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int k = k0 + 2 * j;
      float v0 = k < K ? x[tid * K + k] : (k == K ? 1.f : 0.f);
      float v1 = (k + 1) < K ? x[tid * K + k + 1] : ((k + 1) == K ? 1.f : 0.f);
      pk[j] = Half16<FMT>::pack2(v0, v1);
    }
    *reinterpret_cast<uint4*>(a_tile + km8_offset(128, tid, k0)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const int d_col = swap_lbo_sbo >> 8;   // test hook: accumulator column offset
  const bool a_in_tmem = (swap_lbo_sbo & 2) != 0;   // test hook: A operand from tensor memory (TS mode)
  const bool warp_issue = (swap_lbo_sbo & 4) != 0;  // test hook: whole-warp issue, elect-predicated instructions
  swap_lbo_sbo &= 1;
  const uint32_t tmem_base = tmem_base_holder + d_col;
  const uint32_t a_tmem = tmem_base_holder + 384;   // A image: Kp/2 columns (<= 128)
  if (a_in_tmem) {
    // row = tid = TMEM lane; one 32-bit column holds two K-adjacent 16-bit elements
    for (int k0 = 0; k0 < Kp; k0 += 16) {
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int k = k0 + 2 * j;
        float v0 = k < K ? x[tid * K + k] : (k == K ? 1.f : 0.f);
        float v1 = (k + 1) < K ? x[tid * K + k + 1] : ((k + 1) == K ? 1.f : 0.f);
        pk[j] = Half16<FMT>::pack2(v0, v1);
      }
      tmem_st8(a_tmem + ((uint32_t)(warp * 32) << 16) + (k0 >> 1), pk);
    }
    tmem_st_wait();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
  }

  if (warp_issue) {
    if (warp == 0) {     // the engine's issuer structure (tc_engine.cuh issuer_role)
      const uint32_t bytes = (uint32_t)Np * Kp * 2;
      if (tid == 0) {
        mbar_expect_tx(&bar_w, bytes);
        tma_bulk_g2s(b_tile, wp, bytes, &bar_w);
      }
      mbar_wait(&bar_w, 0);
      __syncwarp();
      tc_fence_after_sync();
      const uint32_t idesc = make_idesc_f16(FMT, 128, Np);
      const uint32_t tm_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t a0 = __shfl_sync(0xffffffffu, smem_u32(a_tile), 0), b0 = __shfl_sync(0xffffffffu, smem_u32(b_tile), 0);
      const uint32_t bar_u = __shfl_sync(0xffffffffu, smem_u32(&bar_mma), 0);
      const uint32_t hi = (uint32_t)(make_smem_desc(0, 0, 128) >> 32);
      uint32_t a_lo = (a0 >> 4) | ((128u * 16u >> 4) << 16), b_lo = (b0 >> 4) | (((uint32_t)Np * 16u >> 4) << 16);
      uint32_t acc = 0;
      for (int ks = 0; ks < Kp / 16; ++ks) {
        umma_f16_elect(tm_u, a_lo, hi, b_lo, hi, idesc, acc);
        acc = 1;
        a_lo += 2 * (128 * 16 >> 4);
        b_lo += 2 * ((uint32_t)Np * 16 >> 4);
      }
      umma_commit_elect(bar_u);
    }
  } else if (tid == 0) {
    const uint32_t bytes = (uint32_t)Np * Kp * 2;
    mbar_expect_tx(&bar_w, bytes);
    tma_bulk_g2s(b_tile, wp, bytes, &bar_w);
    mbar_wait(&bar_w, 0);
    tc_fence_after_sync();
    const uint32_t idesc = make_idesc_f16(FMT, 128, Np);
    uint32_t lbo_a = 128 * 16, lbo_b = Np * 16, sbo = 128;
    for (int ks = 0; ks < Kp / 16; ++ks) {
      uint32_t a_addr = smem_u32(a_tile) + ks * 2 * lbo_a;
      uint32_t b_addr = smem_u32(b_tile) + ks * 2 * lbo_b;
      uint64_t ad = swap_lbo_sbo ? make_smem_desc(a_addr, sbo, lbo_a) : make_smem_desc(a_addr, lbo_a, sbo);
      uint64_t bd_ = swap_lbo_sbo ? make_smem_desc(b_addr, sbo, lbo_b) : make_smem_desc(b_addr, lbo_b, sbo);
      if (a_in_tmem) umma_f16_ts(tmem_base, a_tmem + ks * 8, bd_, idesc, ks > 0 ? 1u : 0u);
      else umma_f16(tmem_base, ad, bd_, idesc, ks > 0 ? 1u : 0u);
    }
    umma_commit(&bar_mma);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0);
  tc_fence_after_sync();
  // epilogue: warp w owns TMEM lanes [32w, 32w+32) = rows
  for (int c = 0; c < Np; c += 16) {
    float v[16];
    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (c + j < N) y[(size_t)tid * N + c + j] = v[j];
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base - d_col);
}

}  // namespace tc
}  // namespace bd

using namespace bd;

extern "C" int bd_tc_selftest(const float* x, const float* w, const float* b, int K, int N, int fmt,
                              int swap_lbo_sbo, void* ws, size_t ws_bytes, float* y,
                              bd_stream_t stream) {
  BD_CHECK_ARG(x && w && b && y && ws, "bd_tc_selftest: null pointer");
  BD_CHECK_ARG(K >= 1 && N >= 1 && N <= 256 && (fmt == 0 || fmt == 1), "bd_tc_selftest: bad K/N/fmt");
  const int Kp = (K + 1 + 15) / 16 * 16, Np = (N + 15) / 16 * 16;
  BD_CHECK_ARG((size_t)Np * Kp * 2 <= ws_bytes, "bd_tc_selftest: workspace too small");
  const size_t smem = (size_t)(128 + Np) * Kp * 2;
  BD_CHECK_ARG(smem <= 220 * 1024, "bd_tc_selftest: tile does not fit shared memory");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  tc::PackTable tab{};
  tab.njobs = 1;
  tc::PackJob& j = tab.job[0];
  j.w = w; j.bias = b; j.dst_off = 0; j.ld = K; j.row0 = 0; j.N = N; j.Np = Np; j.Kp = Kp;
  j.bias_k = K; j.nseg = 1; j.seg[0] = {0, 0, K}; j.transpose = 0;
  dim3 pgrid((unsigned)((Np * Kp + 255) / 256), 1);
  if (fmt == 0) {
    tc::pack_weights_kernel<0><<<pgrid, 256, 0, s>>>(tab, static_cast<uint16_t*>(ws));
    BD_CUDA_LAUNCH_CHECK();
    cudaFuncSetAttribute(tc::selftest_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    tc::selftest_kernel<0><<<1, 128, smem, s>>>(x, K, Kp, Np, static_cast<const uint16_t*>(ws), y, N, swap_lbo_sbo);
  } else {
    tc::pack_weights_kernel<1><<<pgrid, 256, 0, s>>>(tab, static_cast<uint16_t*>(ws));
    BD_CUDA_LAUNCH_CHECK();
    cudaFuncSetAttribute(tc::selftest_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    tc::selftest_kernel<1><<<1, 128, smem, s>>>(x, K, Kp, Np, static_cast<const uint16_t*>(ws), y, N, swap_lbo_sbo);
  }
  BD_CUDA_LAUNCH_CHECK();
  return BD_OK;
}

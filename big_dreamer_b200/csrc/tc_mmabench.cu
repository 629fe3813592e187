// Micro-benchmark (debug entry, not part of the product path): cycles per tcgen05.mma for a chain of
// MMAs accumulating into one TMEM tile, for the no-swizzle KM8 operand layout and for SWIZZLE_128B.
// Operand contents are irrelevant (smem is zero-filled); only fetch/issue rates are measured.
#include "api_internal.h"
#include "../../include/bd_b200_test.h"
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
// Micro-benchmark: the same MMA chain as the engine issues for one layer (KM8 tiles, `ksteps`
// distinct K=16 steps of A and B), optionally with a producer warp streaming bulk TMA copies into
// a 4-slot ring at the same time (tma != 0) and epilogue-like warps hammering shared memory
// with 16-byte stores (stw != 0).  Reports cycles for nrep passes over the K range.
namespace bd {
namespace tc {
__global__ void __launch_bounds__(320, 1) mma_bench2_kernel(int N, int ksteps, int nrep, int tma, int stw,
                                                            const uint8_t* gsrc, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, rbar[4];
  __shared__ uint32_t holder;
  __shared__ volatile int done;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* A = smem;                       // 128 x 208 x 2 = 53248
  uint8_t* B = smem + 53248;               // up to 256 x 208 x 2 = 106496 -> cap N*ksteps*32
  uint8_t* ring = smem + 53248 + 90112;    // 4 x 13312
  uint8_t* scratch = ring + 4 * 13312;     // 16 KB for the store warps
  if (tid == 0) {
    mbar_init(&bar, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&rbar[i], 1);
    fence_barrier_init();
    done = 0;
  }
  if (warp == 1) tmem_alloc<512>(&holder);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = holder;
  if (warp == 0) {
    const uint32_t idesc = make_idesc_f16(0, 128, N);
    const uint64_t ad = make_smem_desc(smem_u32(A), 128 * 16, 128);
    const uint64_t bdsc = make_smem_desc(smem_u32(B), N * 16, 128);
    const uint32_t a_step = (2 * 128 * 16) >> 4, b_step = (2 * N * 16) >> 4;
    const long long t0 = clock64();
    for (int r = 0; r < nrep; ++r) {
      // stw bits: 1 = store warps, 2 = commit to a dummy barrier every 2 MMAs, 4 = tcgen05 fence
      // every 2 MMAs, 8 = elect + syncwarp every 2 MMAs (the engine's per-stage structure)
      if (stw & 256) {          // CUTLASS style: warp-uniform descriptors, only the MMA itself elected
        // values made provably warp-uniform (shfl from lane 0) so they live in uniform registers
        const uint32_t tm_u = __shfl_sync(0xffffffffu, tmem, 0);
        const uint32_t alo = __shfl_sync(0xffffffffu, (uint32_t)ad, 0), ahi = __shfl_sync(0xffffffffu, (uint32_t)(ad >> 32), 0);
        const uint32_t blo = __shfl_sync(0xffffffffu, (uint32_t)bdsc, 0), bhi = __shfl_sync(0xffffffffu, (uint32_t)(bdsc >> 32), 0);
        const uint32_t id_u = __shfl_sync(0xffffffffu, idesc, 0);
        const uint32_t bar_u = __shfl_sync(0xffffffffu, smem_u32(&rbar[3]), 0);
        for (int ks = 0; ks < ksteps; ++ks) {
          umma_f16_elect(tm_u + (uint32_t)(((stw & 512) ? 0 : (r & 1)) * 256), alo + ks * a_step, ahi, blo + ks * b_step, bhi, id_u, ks > 0 ? 1u : 0u);
          if ((stw & 2) && (ks & 1)) umma_commit_elect(bar_u);
        }
        __syncwarp();
      } else if (stw & 0xE) {
        const int grp = (stw >> 4) ? (stw >> 4) : 2;      // MMAs per elected block
        for (int k0 = 0; k0 < ksteps; k0 += grp) {
          if (stw & 4) tc_fence_after_sync();
          if (elect_one()) {
            for (int ks = k0; ks < min(k0 + grp, ksteps); ++ks)
              umma_f16(tmem + (uint32_t)((r & 1) * 256), ad + (uint64_t)(ks * a_step), bdsc + (uint64_t)(ks * b_step), idesc, ks > 0 ? 1u : 0u);
            if (stw & 2) umma_commit(&rbar[3]);
          }
          __syncwarp();
        }
      } else {
        if (elect_one()) {
          for (int ks = 0; ks < ksteps; ++ks)
            umma_f16(tmem + (uint32_t)((r & 1) * 256), ad + (uint64_t)(ks * a_step), bdsc + (uint64_t)(ks * b_step), idesc, ks > 0 ? 1u : 0u);
        }
        __syncwarp();
      }
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (tid == 0) { out[0] = t1 - t0; done = 1; }
  } else if (warp == 2 && tma) {
    uint32_t ph[4] = {0, 0, 0, 0};
    int n = 0;
    for (int it = 0; !done && it < 100000; ++it) {
      const int s = it & 3;
      if (it >= 4) { mbar_wait(&rbar[s], ph[s]); ph[s] ^= 1; }
      if (elect_one()) {
        mbar_expect_tx(&rbar[s], 13312);
        tma_bulk_g2s(ring + s * 13312, gsrc + (size_t)(it % 64) * 13312, 13312, &rbar[s]);
      }
      __syncwarp();
      ++n;
    }
    // drain
    for (int k = 0; k < 4 && k < n; ++k) { const int s = (n - 1 - k) & 3; (void)s; }
    if ((tid & 31) == 0) out[2] = n;
    __nanosleep(20000);
  } else if (warp >= 2 && (stw & 512)) {
    // epilogue-like TMEM reads (of the accumulator half the MMAs are not writing) + ELU-like math
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256;
    float acc_ = 0.f;
    long long cnt = 0;
    for (int it = 0; !done && it < 1000000; ++it) {
      float v[32];
      tmem_ld32(trow + ((it & 3) * 32), v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc_ += v[j];
      ++cnt;
    }
    if (acc_ == 1234.5f) out[1] = 1;
    if (tid == 96) out[3] = cnt;
  } else if (warp >= 2 && (stw & 1) && !(warp == 2 && tma)) {
    long long cnt = 0;
    const int et = tid - 64;
    for (int it = 0; !done && it < 1000000; ++it) {
      *reinterpret_cast<uint4*>(scratch + ((et * 16 + it * 4096) & 16383)) = make_uint4(it, et, 0, 0);
      ++cnt;
    }
    if (tid == 96) out[3] = cnt;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}
}  // namespace tc
}  // namespace bd

extern "C" int bd_tc_mmabench2(int N, int ksteps, int nrep, int tma, int stw, const void* gsrc, long long* out, bd_stream_t stream) {
  using namespace bd;
  BD_CHECK_ARG(N >= 16 && N <= 256 && (N % 16) == 0 && ksteps >= 1 && ksteps <= 13 && (size_t)N * ksteps * 32 <= 90112 && out,
               "bd_tc_mmabench2: bad args");
  const int sm = 53248 + 90112 + 4 * 13312 + 16384;
  cudaFuncSetAttribute(tc::mma_bench2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
  tc::mma_bench2_kernel<<<1, 320, sm, static_cast<cudaStream_t>(stream)>>>(N, ksteps, nrep, tma, stw, static_cast<const uint8_t*>(gsrc), out);
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

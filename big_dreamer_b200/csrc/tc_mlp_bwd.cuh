// Tensor-core backward of a build_mlp chain (DenseModel; the reward / value heads and the actor
// inside Dreamer's behaviour-learning block): per 128-row tile, recompute the hidden activations,
// run the dgrad chain dY_{l-1} = (dY_l W_l) . act'(h_{l-1}) on tcgen05, and -- when parameter
// gradients are wanted -- leave 16-bit KM8 images of every layer input X_l and every dY_l in a
// scratch area.  A second, streaming kernel (wgrad_kernel) then contracts those images over the
// row dimension with MN-major UMMA descriptors (the same tile viewed transposed), accumulating
// dW for one layer per CTA in TMEM across all of its tiles and flushing once.  The constant-1
// column of X_l makes the bias gradient fall out of the same contraction.
#pragma once
#include "tc_engine.cuh"

namespace bd {
namespace tc {

enum MlpBwdEpi : uint8_t {
  EPI_B_ACT_SAVE = 1,  // forward recompute: h_l = act(D) -> H tile + scratch image
  EPI_B_LOAD_DY = 2,   // (no MMA) dy -> G tile
  EPI_B_DACT = 3,      // dY_{l-1} = D . act'(h_{l-1}) -> G tile
  EPI_B_DX = 4         // d[x1 | x2] = D -> global
};

struct MlpBwdArgs {
  Program prog;
  SmemPlan sm;
  const uint16_t* wpack;
  long long N;
  int T;  // = 1
  long long* prof;
  int k1, k2, out, n_layers, act;
  int Kp_b, Ks, Kp_h, Kp_g;
  int want_images;              // 1: also dump X / dY images for wgrad
  int need_x;                   // 0: inputs are not needed (saved hidden images, no wgrad): skip tile init
  const float *x1, *x2, *dy;
  // optional second input segment: rows >= split come from x1b / x2b at (row - split).  The actor's
  // inputs of imagine_ahead are (prev_belief, prev_state) for step 0 and (beliefs, states)[t-1] after.
  const float *x1b, *x2b;
  long long split;              // >= N when unused
  // Segmented tiling (the actor's backward over the T steps of a rollout whose forward saved its hidden images per
  // (t, tile)): tile g = tile_base + tile covers rows [t N' + 128 j, ...) of step t = g / seg_tiles, j = g % seg_tiles,
  // with N' = seg_rows rows per step; all row pointers are then those of row 0.  seg_tiles = 0: plain tiling of N rows.
  long long seg_rows, tile_base, ntiles;
  int seg_tiles;
  // Paired tiles (saved hidden images, more tiles than SMs): a CTA keeps TWO row tiles in flight -- the program
  // interleaves their dgrad chains (phase.pad = sub-tile, each phase depending on the epilogue two phases back), so
  // the MMAs of one tile run under the epilogue of the other.  Each sub-tile owns one G tile (rewritten in place:
  // the MMAs that read it are complete when its epilogue runs) and one 256-column TMEM half.  nloop = tile pairs.
  int pair;                       // 1: two row tiles (above); 2: two HEADS on the same row tile (heads_pair_backward): the
                                  // sub-chains differ in weights, upstream gradient (dy / dy_b) and hidden images
                                  // (xs / xs_b); their last GEMMs accumulate into ONE dX accumulator
  const float* dy_b;              // pair == 2: upstream gradient of the second head
  uint16_t* xs_b[BD_MAX_LAYERS];  // pair == 2: hidden images of the second head
  long long nloop;
  int c2pair;                     // 1: CTA-pair launch (clusters of 2, cta_group::2 MMAs, half weight stages per CTA; see
                                  // tc_engine.cuh issuer_pair_role); ntiles is even, nloop = tile pairs
  int dbg;                        // BD_BWD_DBG builds only (BD_BWD_DBGV): 1 no hidden-image loads, 2 no image stores, 4 no epilogue math
  float *dx1, *dx2;
  uint16_t* xs[BD_MAX_LAYERS];  // xs[l]: images of hidden h_l (cols kp_xs[l]), l = 0..L-2
  uint16_t* ds[BD_MAX_LAYERS];  // ds[l]: images of dY_l (cols kp_ds[l]), l = 0..L-1
  uint16_t *x0b, *x0s;          // images of [x1 | 1] and x2
  int kp_xs[BD_MAX_LAYERS], kp_ds[BD_MAX_LAYERS];
  const unsigned int* amax_bits;  // device: bit pattern of max|dy| (see grad_scale)
  PrefetchPlan pf;                // next tile's inputs, pulled into L2 by the producer warp
};

// Gradients reach this kernel scaled by 1/(T*N) and can sit far below fp16's normal range, so the
// 16-bit operands carry dy * 2^k with k chosen (on the device, from max|dy|) so that the largest
// magnitude lands in [128, 256); results are multiplied by 2^-k on the way out.  Powers of two
// make the scaling exact.
// `shift` lowers the target by 2^shift (headroom when the kernel itself amplifies the upstream
// gradients, e.g. the fused heads: a lambda-return adjoint sums up to 1 / (1 - disc lam) terms).
__device__ __forceinline__ float grad_scale(const unsigned int* amax_bits, float* inv, int shift = 0) {
  const float amax = __uint_as_float(*amax_bits);
  int e = 0;
  float sc = 1.f;
  if (amax > 0.f && amax < 3.0e38f) {
    frexpf(amax, &e);
    int k = 8 - shift - e;
    k = max(-100, min(100, k));
    sc = ldexpf(1.f, k);
  }
  *inv = 1.f / sc;
  return sc;
}
static __global__ void absmax_kernel(const float* __restrict__ x, long long n, unsigned int* out) {
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float v = fabsf(x[i]);
    if (v < 3.0e38f) m = fmaxf(m, v);      // ignore inf / nan
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}

// several tensors in one launch (blockIdx.y selects the tensor): the BPTT's upstream gradients
struct AbsmaxJobs { const float* x[8]; long long n[8]; };
static __global__ void absmax_multi_kernel(const AbsmaxJobs jobs, unsigned int* out) {
  const float* __restrict__ x = jobs.x[blockIdx.y];
  const long long n = jobs.n[blockIdx.y];
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float v = fabsf(x[i]);
    if (v < 3.0e38f) m = fmaxf(m, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}

template <int FMT>
__device__ __forceinline__ void unpack8(const uint4& u, float* v) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (FMT == 0) {
      const __half2 h = *reinterpret_cast<const __half2*>(&w[j]);
      const float2 f = __half22float2(h);
      v[2 * j] = f.x; v[2 * j + 1] = f.y;
    } else {
      const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
      const float2 f = __bfloat1622float2(h);
      v[2 * j] = f.x; v[2 * j + 1] = f.y;
    }
  }
}
template <int FMT>
__device__ __forceinline__ uint4 pack8(const float* v) {
  return make_uint4(Half16<FMT>::pack2(v[0], v[1]), Half16<FMT>::pack2(v[2], v[3]),
                    Half16<FMT>::pack2(v[4], v[5]), Half16<FMT>::pack2(v[6], v[7]));
}

// 12 epilogue warps (3 column parts x 4 TMEM quadrants, 14 warps = 128 registers per thread), 16-column chunks
constexpr int kBwdParts = 3;
constexpr int kBwdEpiThreads = kBwdParts * 128;
constexpr int kBwdThreads = 64 + kBwdEpiThreads;
template <int FMT, int ACT, bool C2 = false>
__global__ void __launch_bounds__(kBwdThreads, 1) mlp_bwd_kernel(const __grid_constant__ MlpBwdArgs A_) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const MlpBwdArgs& a = A_;
  uint8_t* smem = smem_raw;
  __shared__ EngineShared sh;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform: the role code stays on the uniform datapath
  __shared__ Program sprog;
  stage_program(sprog, a.prog);
  constexpr bool c2 = C2;                                   // CTA-pair variant (launched in clusters of 2)
  const uint32_t crank = c2 ? (blockIdx.x & 1u) : 0u;       // = %cluster_ctarank for (2,1,1) clusters
  const uint32_t tmem_base = engine_setup<C2>(sh, a.sm.nstage, 1, kBwdEpiThreads, 1);
  const long long ntiles = a.ntiles;
  const long long nloop = (a.pair == 1 || c2) ? a.nloop : a.ntiles;     // loop items per CTA walk: tiles, or tile pairs
  const Program& P = sprog;

  if (warp == 0) {
    producer_role(P, a.sm, a.wpack, nloop, 1, smem, sh, &a.pf, c2 ? 2u : 1u, 1, false, c2 ? (int)crank : -1);
  } else if (warp == 1) {
    if constexpr (!C2) {
      issuer_role<FMT, false>(P, a.sm, nloop, 1, smem, sh, tmem_base, nullptr);
    } else {
      if (crank == 0) issuer_pair_role<FMT>(P, a.sm, nloop, 1, smem, sh, tmem_base);
      else relay_role(P, a.sm, nloop, 1, sh);
    }
  } else {
    const int q = warp & 3, part = (warp - 2) >> 2;
    const uint32_t lead_epi_done = c2 ? mapa_u32(smem_u32(&sh.epi_done[0]), 0) : 0u;
    auto epi_arrive = [&](uint32_t ge) {     // one arrival per warp (CTA pair: on the leader's barrier)
      if (!c2) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.epi_done[ge & 7]);
      } else {
        fence_proxy_async_all();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(lead_epi_done + (ge & 7) * 8);
      }
    };
    const int row = q * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const int etid = tid - 64;
    const uint32_t rowoff = (row >> 3) * 128 + (row & 7) * 16;   // byte offset of this row in a KM8 group
    uint8_t* B0 = smem + a.sm.off_tile[0];
    uint8_t* SA = smem + a.sm.off_tile[2];
    uint8_t* H = smem + a.sm.off_tile[3];
    uint32_t Ge = 0, Gm = 0;
    float inv_scale;
    const float scale = grad_scale(a.amax_bits, &inv_scale);
    const int nsub = a.pair ? 2 : 1;
    const long long it0 = c2 ? blockIdx.x / 2 : blockIdx.x, itstep = c2 ? gridDim.x / 2 : gridDim.x;
    for (long long it = it0; it < nloop; it += itstep) {
      // the one or two row tiles of this walk: index, first row, valid rows (0 for the missing half of an odd pair,
      // which runs through the phases on zeros and touches no global memory)
      long long tl[2] = {0, 0}, tr0[2] = {0, 0};
      int tv[2] = {0, 0};
      bool tok[2] = {false, false};
#pragma unroll
      for (int sb = 0; sb < 2; ++sb) {
        const long long tile_s = c2 ? it * 2 + crank : (a.pair == 1 ? it * 2 + sb : it);
        if (sb >= nsub || tile_s >= ntiles) continue;
        tok[sb] = true; tl[sb] = tile_s;
        if (a.seg_tiles > 0) {
          const long long g = a.tile_base + tile_s, tt = g / a.seg_tiles;
          const long long j0 = (g - tt * a.seg_tiles) * kTileRows;
          tr0[sb] = tt * a.seg_rows + j0;
          tv[sb] = (int)(a.seg_rows - j0 < kTileRows ? a.seg_rows - j0 : kTileRows);
        } else {
          tr0[sb] = tile_s * kTileRows;
          tv[sb] = (int)(a.N - tr0[sb] < kTileRows ? a.N - tr0[sb] : kTileRows);
        }
      }
      // ---------------- init: B0 <- [x1 | 1], SA <- x2 (and their images for wgrad); paired tiles (no forward
      // recompute) only leave the images
      if (a.need_x) {
#pragma unroll
       for (int sb = 0; sb < 2; ++sb) {
        if (!tok[sb]) continue;
        const long long tile = tl[sb], trow0 = tr0[sb];
        const int tvalid = tv[sb];
        const int gb = a.Kp_b >> 3;
        if ((a.k1 & 3) == 0 && ((reinterpret_cast<uintptr_t>(a.x1) | reinterpret_cast<uintptr_t>(a.x1b)) & 15) == 0) {
          // 16-byte loads, four items (8 columns of one row) requested per thread before the first is used (one
          // item at a time with scalar loads: ~9 dependent memory round trips per tile in this prologue)
          for (int i0 = etid; i0 < kTileRows * gb; i0 += 4 * kBwdEpiThreads) {
            float4 lo[4], hi[4];
#pragma unroll
            for (int u4 = 0; u4 < 4; ++u4) {
              const int i = i0 + u4 * kBwdEpiThreads;
              const int kg = i / kTileRows, r = i - kg * kTileRows;
              const long long gr = trow0 + r;
              lo[u4] = hi[u4] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (i < kTileRows * gb && r < tvalid) {
                const float* src = (gr < a.split ? a.x1 + gr * a.k1 : a.x1b + (gr - a.split) * a.k1) + kg * 8;
                if (kg * 8 + 4 <= a.k1) lo[u4] = *reinterpret_cast<const float4*>(src);
                if (kg * 8 + 8 <= a.k1) hi[u4] = *reinterpret_cast<const float4*>(src + 4);
              }
            }
#pragma unroll
            for (int u4 = 0; u4 < 4; ++u4) {
              const int i = i0 + u4 * kBwdEpiThreads;
              if (i >= kTileRows * gb) break;
              const int kg = i / kTileRows, r = i - kg * kTileRows;
              float v[8] = {lo[u4].x, lo[u4].y, lo[u4].z, lo[u4].w, hi[u4].x, hi[u4].y, hi[u4].z, hi[u4].w};
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (kg * 8 + j == a.k1 && r < tvalid) v[j] = 1.f;       // (k1 is a multiple of 4: starts a 4-group)
              const uint4 u = pack8<FMT>(v);
              if (!a.pair) *reinterpret_cast<uint4*>(B0 + kg * kLboA + r * 16) = u;
              if (a.want_images)
                *reinterpret_cast<uint4*>(a.x0b + (size_t)tile * kTileRows * a.Kp_b + (size_t)kg * kTileRows * 8 + r * 8) = u;
            }
          }
        } else {
        for (int i = etid; i < kTileRows * gb; i += kBwdEpiThreads) {
          const int kg = i / kTileRows, r = i - kg * kTileRows;
          const long long gr = trow0 + r;
          const bool rv = r < tvalid;
          const float* x1r = gr < a.split ? a.x1 + gr * a.k1 : a.x1b + (gr - a.split) * a.k1;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int k = kg * 8 + j;
            v[j] = (k < a.k1) ? (rv ? x1r[k] : 0.f) : ((k == a.k1 && rv) ? 1.f : 0.f);
          }
          const uint4 u = pack8<FMT>(v);
          if (!a.pair) *reinterpret_cast<uint4*>(B0 + kg * kLboA + r * 16) = u;
          if (a.want_images)
            *reinterpret_cast<uint4*>(a.x0b + (size_t)tile * kTileRows * a.Kp_b + (size_t)kg * kTileRows * 8 + r * 8) = u;
        }
        }
        const int gs = a.Ks >> 3;
        for (int i = etid; i < kTileRows * gs; i += kBwdEpiThreads) {
          const int kg = i / kTileRows, r = i - kg * kTileRows;
          const long long gr = trow0 + r;
          const bool rv = r < tvalid;
          const float* x2r = gr < a.split ? a.x2 + gr * a.k2 : a.x2b + (gr - a.split) * a.k2;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int k = kg * 8 + j;
            v[j] = (k < a.k2 && rv) ? x2r[k] : 0.f;
          }
          const uint4 u = pack8<FMT>(v);
          if (!a.pair) *reinterpret_cast<uint4*>(SA + kg * kLboA + r * 16) = u;
          if (a.want_images)
            *reinterpret_cast<uint4*>(a.x0s + (size_t)tile * kTileRows * a.Ks + (size_t)kg * kTileRows * 8 + r * 8) = u;
        }
       }
      }
      epi_arrive(Ge);
      ++Ge;
      for (int pi = 0; pi < P.n_phases; ++pi) {
        const Phase ph = P.p[pi];
        const uint32_t tacc = trow + ph.d_col;
        const int sb = a.pair ? (int)ph.pad : 0;              // sub-tile of this phase
        const long long tile = sb ? tl[1] : tl[0], trow0 = sb ? tr0[1] : tr0[0];
        const int tvalid = sb ? tv[1] : tv[0];
        const bool tile_ok = sb ? tok[1] : tok[0];
        const long long grow = trow0 + row;
        const bool rvalid = row < tvalid;
        uint8_t* Gt = smem + a.sm.off_tile[ph.out_tile];
#define BD_WAIT_ACC()                                        \
  do {                                                       \
    mbar_wait(&sh.acc_full[Gm & 3], (Gm >> 2) & 1);          \
    tc_fence_after_sync();                                   \
  } while (0)
        switch (ph.epi) {
          case EPI_B_ACT_SAVE: {
            BD_WAIT_ACC();
            const int l = ph.aux0, nv = ph.n_valid;
            uint16_t* img = a.xs[l] + (size_t)tile * kTileRows * ph.Kp_out;
            for (int c = part * 16; c < ph.Kp_out; c += kBwdParts * 16) {
              float v[16];
              if (c < ph.Np) tmem_ld16(tacc + c, v);
              else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = 0.f;
              }
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = tc_act_t<ACT>(v[j]);
              if (c + 16 > nv) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const int col = c + j;
                  if (col >= nv) v[j] = (col == nv && rvalid) ? 1.f : 0.f;
                }
              }
              if (!rvalid) {   // padded rows carry exact zeros so they add nothing to dW / db
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = 0.f;
              }
              uint8_t* p = H + (c >> 3) * kLboA + rowoff;
              uint16_t* gi = img + (size_t)(c >> 3) * kTileRows * 8 + row * 8;
#pragma unroll
              for (int g8 = 0; g8 < 2; ++g8) {
                const uint4 u = pack8<FMT>(v + 8 * g8);
                *reinterpret_cast<uint4*>(p + g8 * kLboA) = u;
                *reinterpret_cast<uint4*>(gi + (size_t)g8 * kTileRows * 8) = u;
              }
            }
          } break;
          case EPI_B_LOAD_DY: {
            BD_WAIT_ACC();
            const int l = a.n_layers - 1, kp = a.kp_ds[l];
            uint16_t* img = a.ds[l] + (size_t)tile * kTileRows * kp;
            for (int c = part * 8; c < kp; c += kBwdParts * 8) {
              float v[8];
              const float* dyp = (a.pair == 2 && sb) ? a.dy_b : a.dy;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                v[j] = (rvalid && c + j < a.out) ? dyp[grow * a.out + c + j] * scale : 0.f;
              const uint4 u = pack8<FMT>(v);
              *reinterpret_cast<uint4*>(Gt + (c >> 3) * kLboA + rowoff) = u;
              if (a.want_images && tile_ok) *reinterpret_cast<uint4*>(img + (size_t)(c >> 3) * kTileRows * 8 + row * 8) = u;
            }
          } break;
          case EPI_B_DACT: {
            const int l = ph.aux0;          // produces dY_l from D and h_l
            const int nv = ph.n_valid, kp = a.kp_ds[l];
            const uint16_t* himg = ((a.pair == 2 && sb) ? a.xs_b[l] : a.xs[l]) + (size_t)tile * kTileRows * a.kp_xs[l];
            uint16_t* img = a.ds[l] + (size_t)tile * kTileRows * kp;
            // 16-column chunks c = 16 (part + kBwdParts k); the hidden-image pieces of three chunks are in flight
            // (rotating registers), the first ones requested before the accumulator wait
            const uint16_t* hrow = himg + row * 8;
            const int cstep = kBwdParts * 16;
#ifdef BD_BWD_DBG
            const bool dbg_nold = a.dbg & 1, dbg_nost = a.dbg & 2;
#else
            constexpr bool dbg_nold = false, dbg_nost = false;
#endif
            auto ld0 = [&](int c) { return (c < kp && !dbg_nold) ? *reinterpret_cast<const uint4*>(hrow + (size_t)(c >> 3) * kTileRows * 8) : make_uint4(0, 0, 0, 0); };
            auto ld1 = [&](int c) { return (c < kp && !dbg_nold) ? *reinterpret_cast<const uint4*>(hrow + (size_t)((c >> 3) + 1) * kTileRows * 8) : make_uint4(0, 0, 0, 0); };
            const int cfirst = part * 16;
            uint4 a0 = ld0(cfirst), a1 = ld1(cfirst);
            uint4 b0 = ld0(cfirst + cstep), b1 = ld1(cfirst + cstep);
            uint4 d0 = ld0(cfirst + 2 * cstep), d1 = ld1(cfirst + 2 * cstep);
            BD_WAIT_ACC();
            for (int c = cfirst; c < kp; c += cstep) {
              float v[16];
              if (c < ph.Np) tmem_ld16(tacc + c, v);
              else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = 0.f;
              }
              const uint4 hu[2] = {a0, a1};
              a0 = b0; a1 = b1; b0 = d0; b1 = d1;
              d0 = ld0(c + 3 * cstep); d1 = ld1(c + 3 * cstep);
              tmem_ld_wait();
              // (padded rows: the packed result is masked to exact zeros -- they add nothing to dW / db; padded
              // columns, only in the chunk that crosses n_valid: zeroed before the product)
              const uint32_t rmask = rvalid ? 0xFFFFFFFFu : 0u;
              if (c + 16 > nv) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (c + j >= nv) v[j] = 0.f;
              }
#pragma unroll
              for (int g8 = 0; g8 < 2; ++g8) {
                float h[8];
                unpack8<FMT>(hu[g8], h);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[g8 * 8 + j] *= tc_dact_from_out<ACT>(h[j]);
                uint4 u = pack8<FMT>(v + 8 * g8);
                u.x &= rmask; u.y &= rmask; u.z &= rmask; u.w &= rmask;
                *reinterpret_cast<uint4*>(Gt + ((c >> 3) + g8) * kLboA + rowoff) = u;
                if (a.want_images && tile_ok && !dbg_nost)
                  *reinterpret_cast<uint4*>(img + (size_t)((c >> 3) + g8) * kTileRows * 8 + row * 8) = u;
              }
            }
          } break;
          case EPI_B_DX: {
            BD_WAIT_ACC();
            const int nin = a.k1 + a.k2;
            // dX rows of a tile are one contiguous block of dx1 (and of dx2): stage the tile row-major
            // in shared memory (every operand tile is dead once this phase's MMAs are done) and copy
            // it out coalesced, instead of one 16-byte store per row per warp instruction.
            const int st1 = a.k1 + ((12 - a.k1 % 8) % 8);          // row stride (floats): = 4 mod 8 -> conflict-free float4
            const int st2 = a.k2 | 1;
            float* stage1 = reinterpret_cast<float*>(smem + a.sm.off_tile[0]);
            float* stage2 = stage1 + (size_t)kTileRows * st1;
            const bool staged = a.pair != 1 && (size_t)kTileRows * (st1 + st2) * 4 <= (size_t)(a.sm.off_ring - a.sm.off_tile[0]);   // (paired tiles: the other tile is live)
            const bool v4 = ((a.k1 & 3) == 0);
            for (int c = part * 16; c < ph.Np; c += kBwdParts * 16) {
              float v[16];
              tmem_ld16(tacc + c, v);
              tmem_ld_wait();
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                const int col = c + j4 * 4;
                const float4 o = make_float4(v[j4 * 4] * inv_scale, v[j4 * 4 + 1] * inv_scale,
                                             v[j4 * 4 + 2] * inv_scale, v[j4 * 4 + 3] * inv_scale);
                const float ov[4] = {o.x, o.y, o.z, o.w};
                if (staged) {
                  if (col + 3 < a.k1) {
                    *reinterpret_cast<float4*>(stage1 + (size_t)row * st1 + col) = o;
                  } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                      const int cc = col + j;
                      if (cc < a.k1) stage1[(size_t)row * st1 + cc] = ov[j];
                      else if (cc < nin) stage2[(size_t)row * st2 + (cc - a.k1)] = ov[j];
                    }
                  }
                } else if (rvalid) {
                  if (v4 && col + 3 < a.k1) {        // whole float4 inside dx1 (rows are 16-byte aligned)
                    if (a.dx1) *reinterpret_cast<float4*>(a.dx1 + grow * a.k1 + col) = o;
                  } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                      const int cc = col + j;
                      if (cc < a.k1) { if (a.dx1) a.dx1[grow * a.k1 + cc] = ov[j]; }
                      else if (cc < nin) { if (a.dx2) a.dx2[grow * a.k2 + (cc - a.k1)] = ov[j]; }
                    }
                  }
                }
              }
            }
            if (staged) {      // uniform per CTA
              asm volatile("bar.sync 1, %0;" ::"n"(kBwdEpiThreads) : "memory");
              const long long row0 = trow0;
              const int nrows = tvalid;
              const int et = tid - 64;
              if (a.dx1) {
                float* dst = a.dx1 + row0 * a.k1;
                if (v4 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                  const int q4 = a.k1 >> 2, n4 = nrows * q4;
                  for (int i = et; i < n4; i += kBwdEpiThreads) {
                    const int r = i / q4, c4 = i - r * q4;
                    reinterpret_cast<float4*>(dst)[i] = *reinterpret_cast<const float4*>(stage1 + (size_t)r * st1 + c4 * 4);
                  }
                } else {
                  const int n = nrows * a.k1;
                  for (int i = et; i < n; i += kBwdEpiThreads) { const int r = i / a.k1; dst[i] = stage1[(size_t)r * st1 + (i - r * a.k1)]; }
                }
              }
              if (a.dx2 && a.k2 > 0) {
                float* dst = a.dx2 + row0 * a.k2;
                const int n = nrows * a.k2;
                for (int i = et; i < n; i += kBwdEpiThreads) { const int r = i / a.k2; dst[i] = stage2[(size_t)r * st2 + (i - r * a.k2)]; }
              }
            }
          } break;
          default: BD_WAIT_ACC(); break;
        }
#undef BD_WAIT_ACC
        tc_fence_before_sync();
        epi_arrive(Ge);
        ++Ge;
        ++Gm;
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if constexpr (C2) {
    cluster_sync_all();        // neither CTA retires while the pair's MMAs / commits may still touch it
    if (warp == 1) tmem_dealloc_pair<512>(tmem_base);
  } else {
    if (warp == 1) tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// Streaming wgrad:  dWp[m, k] += sum_rows dY[row, m] * X[row, k]   per (job = layer part).
// A = dY image and B = X image are the [128 rows x cols] KM8 tiles viewed as MN-major operands
// (M/N = feature index, K = row): 8 features x 16 B is again a canonical core matrix, with
// SBO = 2048 B between feature groups and LBO = 128 B between 8-row groups.
// ---------------------------------------------------------------------------------------------
struct WgradJob {
  const uint16_t* dyimg;   // per tile: 128 * kp_dy elements
  const uint16_t* ximg;    // per tile: 128 * kp_x elements
  float* dwp;              // (256, kp_x) fp32, zero-initialised, accumulated with atomics
  int kp_dy, kp_x, m_valid;
};
struct WgradArgs {
  WgradJob job[BD_MAX_LAYERS + 1];
  long long ntiles;
  uint32_t stage_bytes, nstage;
  uint32_t x_off;          // byte offset of the X image inside a stage (= the widest dY image, 1024-aligned)
  const unsigned int* amax_bits;
};

__host__ __device__ constexpr uint32_t make_idesc_f16_mn(int ab_format, int M, int N) {
  return make_idesc_f16(ab_format, M, N) | (1u << 15) | (1u << 16);   // A and B MN-major
}

template <int FMT>
__global__ void __launch_bounds__(128, 1) wgrad_kernel(const __grid_constant__ WgradArgs A_) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const WgradArgs& a = A_;
  const WgradJob j = a.job[blockIdx.y];
  __shared__ uint64_t full[2], empty[2], done;
  __shared__ uint32_t tmem_holder;
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(&tmem_holder);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_holder;
  const int nmt = j.m_valid > 128 ? 2 : 1;
  const uint32_t bytes_dy = 128u * j.kp_dy * 2, bytes_x = 128u * j.kp_x * 2;
  const long long first = blockIdx.x;
  const bool any = first < a.ntiles;

  if (warp == 0) {
    uint32_t st = 0, ph = 0;
    for (long long tile = first; tile < a.ntiles; tile += gridDim.x) {
      mbar_wait(&empty[st], ph ^ 1);
      if (elect_one()) {
        uint8_t* dst = smem + st * a.stage_bytes;
        mbar_expect_tx(&full[st], bytes_dy + bytes_x);
        tma_bulk_g2s(dst, j.dyimg + (size_t)tile * 128 * j.kp_dy, bytes_dy, &full[st]);
        tma_bulk_g2s(dst + a.x_off, j.ximg + (size_t)tile * 128 * j.kp_x, bytes_x, &full[st]);
      }
      __syncwarp();
      if (++st == a.nstage) { st = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    uint32_t st = 0, ph = 0;
    const uint32_t idesc = make_idesc_f16_mn(FMT, 128, j.kp_x);
    uint32_t acc = 0;
    for (long long tile = first; tile < a.ntiles; tile += gridDim.x) {
      mbar_wait(&full[st], ph);
      tc_fence_after_sync();
      const uint32_t base = smem_u32(smem + st * a.stage_bytes);
      if (elect_one()) {
        for (int mt = 0; mt < nmt; ++mt) {
          // dY image has kp_dy columns; the second M tile may run past them into whatever follows in shared
          // memory (the X image, the next stage: finite or not): those accumulator rows (>= m_valid) are
          // never read back.  The launcher keeps 128 x 256 x 2 bytes readable behind every stage's start.
          const uint64_t ad0 = make_smem_desc(base + mt * 16 * kLboA, 128, kLboA);
          const uint64_t bd0 = make_smem_desc(base + a.x_off, 128, kLboA);
          for (int ks = 0; ks < 8; ++ks)   // K = 128 rows, 16 per MMA = two 8-row groups of 128 B
            umma_f16(tmem_base + mt * 256, ad0 + (uint64_t)(ks * 16), bd0 + (uint64_t)(ks * 16), idesc,
                     (acc | (uint32_t)ks) ? 1u : 0u);
        }
        umma_commit(&empty[st]);
      }
      __syncwarp();
      acc = 1;
      if (++st == a.nstage) { st = 0; ph ^= 1; }
    }
    if (elect_one()) umma_commit(&done);
    __syncwarp();
  }
  // flush: every warp reads its 32 accumulator lanes (= output features)
  mbar_wait(&done, 0);
  tc_fence_after_sync();
  if (any) {
    float inv_scale;
    grad_scale(a.amax_bits, &inv_scale);
    for (int mt = 0; mt < nmt; ++mt) {
      const int m = mt * 128 + tid;
      for (int c = 0; c < j.kp_x; c += 16) {
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + mt * 256 + c, v);
        tmem_ld_wait();
        if (m < j.m_valid) {
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) atomicAdd(j.dwp + (size_t)m * j.kp_x + c + jj, v[jj] * inv_scale);
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// dw / db (+=) from the packed fp32 gradient image, using the forward pack job as the map
struct UnpackJob {
  const float* dwp;
  float *dw, *db;
  int kp, n, ld, nseg, bias_k;
  PackSeg seg[3];
};
struct UnpackTable { int njobs; UnpackJob job[BD_MAX_LAYERS + 1]; };

static __global__ void unpack_dw_kernel(const __grid_constant__ UnpackTable tab) {
  const UnpackJob& j = tab.job[blockIdx.y];
  for (int s = 0; s < j.nseg; ++s) {
    const int total = j.n * j.seg[s].len;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
      const int n = i / j.seg[s].len, c = i - n * j.seg[s].len;
      if (j.dw) j.dw[(size_t)n * j.ld + j.seg[s].src_c0 + c] += j.dwp[(size_t)n * j.kp + j.seg[s].dst_k0 + c];
    }
  }
  if (j.db && j.bias_k >= 0)
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < j.n; n += gridDim.x * blockDim.x)
      j.db[n] += j.dwp[(size_t)n * j.kp + j.bias_k];
}

}  // namespace tc
}  // namespace bd

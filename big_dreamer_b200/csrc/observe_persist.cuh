// Persistent cluster kernels for the RSSM observe pass (TransitionModel.forward with observations,
// src/models.py:239-271) on small batches (BASELINE configs[3]: 49 steps x 50 rows).
//
// The per-step path launches ~10 (forward) / ~20 (backward) 50-row kernels per time step: 1 400
// dependent launches, each an L2 round trip.  Here ONE thread-block cluster of 16 CTAs walks all L
// steps of a 64-row chunk (one cluster per chunk of rows; rows are independent):
//   * every CTA owns a slice of each layer's output columns, for all 64 rows;
//   * layer inputs travel between CTAs as TRANSPOSED [K][64] fp32 blocks in an L2-resident scratch
//     (one contiguous block per layer, streamed into shared memory with 16-byte cp.async.cg in a
//     4-stage ring of 64 k-rows), the weights as per-CTA pre-packed [K][16 slots][4] images
//     (pack_ops_kernel; constant over the steps, so their first ring stages are requested BEFORE
//     the cluster barrier that waits for the other CTAs' activations);
//   * phases are separated by the hardware cluster barrier (arrive.release / wait.acquire).
// Thread tile: 4 consecutive rows x 4 "slots" of one slot group (one float4 of A, one float4 of W
// per k: 16 FMA per 2 LDS.128).  A slot group is either the four pre-activations of one GRU unit
// (r, z, gi_n, gh_n -- the x and h projections are ONE contraction over [x ; h]) or four adjacent
// output columns of a plain layer.  Layers with few slot groups per CTA split K over the idle
// threads (up to 4 ways) and reduce through shared memory.
//
// What is hoisted out of the recurrence (and runs as ordinary batched GEMMs over all L*B rows, see
// api_fp32.cu): the embedding half of the posterior's first layer (K = 1024 of 1224), the whole
// prior branch (its sample is not fed back in observe mode), every weight / bias gradient, the
// recompute of the activations the backward needs, and d embeddings.
//
// Forward phases per step:  P1 x = act(W_sa [s*nt ; a] + b)           (K = S+A, built in place)
//                           P2 GRU: [x ; h] -> b_new                   (K = 2 Be, slots r,z,gi_n,gh_n)
//                           P3 hq = act(b_new W_q1[:, :Be]^T + PE_t)   (K = Be)
//                           P4 (mu, raw) = hq W_q2^T + b -> sample     (K = Hi, slots mu_j, raw_j)
// Backward phases per step: Q1 d_preq (in place, K = 2S) -> d_hq = (d_preq W_q2) . act'(hq)
//                           Q2 G = Gtot_t + carry_b + d_hq W_q1[:, :Be]; GRU gate backward -> planes
//                           Q3 [d_r d_z d_n d_n.r] -> dx = (.. W_ih) . act'(x) | carry_b = G.z + (.. W_hh)
//                           Q4 dsa = dx W_sa -> carry_s = dsa[:, :S] * nt, d_actions
#pragma once
#include "common.cuh"

namespace bd {
namespace f32 {
namespace obs {

constexpr int kC = 16;          // CTAs per cluster (non-portable size; every B200 GPC has >= 16 SMs)
constexpr int kR = 64;          // rows per cluster
constexpr int kKC = 64;         // k rows per ring stage
constexpr int kNS = 4;          // ring stages
constexpr int kThreads = 256;
constexpr int kRowFloats = 64;  // floats per k row of an A block (64 rows) and of a W image (16 x 4)
constexpr int kStageFloats = kKC * kRowFloats;
constexpr int kSmemBytes = (2 * kNS * kStageFloats + kThreads * 16) * (int)sizeof(float);   // 147 456
constexpr int kMaxSmallK = kNS * kKC;   // in-place operands (embed input, d_preq) fit the A ring

enum OpId { OP_EMB = 0, OP_GRU, OP_Q1F, OP_Q2F, OP_B1, OP_B2, OP_B3, OP_B4, OP_COUNT };

struct OpDesc {
  const float* w;   // packed images: CTA c's image at w + c * K * 64
  int K, NJ, Wc, WP, KS;
};
inline OpDesc make_op(int K, int NJ) {
  OpDesc d{};
  d.K = K; d.NJ = NJ;
  d.Wc = (NJ + kC - 1) / kC;
  d.WP = 1;
  while (d.WP < d.Wc) d.WP <<= 1;
  d.KS = 16 / d.WP < 4 ? 16 / d.WP : 4;
  return d;
}
inline size_t op_floats(const OpDesc& d) { return (size_t)kC * d.K * kRowFloats; }

struct PackArgs {
  float* dst[OP_COUNT];
  int K[OP_COUNT], NJ[OP_COUNT], Wc[OP_COUNT];
  int first, count;   // ops [first, first + count)
  int Be, Bep, Hi, S, A, E;
  const float *w_sa, *w_ih, *w_hh, *w_q1, *w_q2;
};

__device__ __forceinline__ float pack_src(const PackArgs& p, int op, int k, int j, int g) {
  const int Be = p.Be, Hi = p.Hi, S = p.S, SA = p.S + p.A, ldq1 = p.Be + p.E;
  const int col = 4 * j + g;
  switch (op) {
    case OP_EMB: return col < Be ? p.w_sa[(size_t)col * SA + k] : 0.f;
    case OP_GRU:
      if (j >= Be) return 0.f;
      if (k < Be) return g < 3 ? p.w_ih[((size_t)g * Be + j) * Be + k] : 0.f;
      if (g == 2) return 0.f;
      return p.w_hh[((size_t)(g == 3 ? 2 : g) * Be + j) * Be + (k - Be)];
    case OP_Q1F: return col < Hi ? p.w_q1[(size_t)col * ldq1 + k] : 0.f;
    case OP_Q2F:
      if (j >= S || g > 1) return 0.f;
      return p.w_q2[((size_t)g * S + j) * Hi + k];
    case OP_B1: return col < Hi ? p.w_q2[(size_t)k * Hi + col] : 0.f;
    case OP_B2: return col < Be ? p.w_q1[(size_t)k * ldq1 + col] : 0.f;
    case OP_B3: {
      const int pl = k / Be, kk = k - pl * Be;
      if (col < p.Bep) {
        if (col >= Be || pl == 3) return 0.f;
        return p.w_ih[((size_t)pl * Be + kk) * Be + col];
      }
      const int c = col - p.Bep;
      if (c >= Be || pl == 2) return 0.f;
      return p.w_hh[((size_t)(pl == 3 ? 2 : pl) * Be + kk) * Be + c];
    }
    case OP_B4: return col < SA ? p.w_sa[(size_t)k * SA + col] : 0.f;
  }
  return 0.f;
}

// one thread per packed float: dst[((c * K + k) * 16 + js) * 4 + g]
static __global__ void pack_ops_kernel(PackArgs p) {
  const int op = p.first + blockIdx.y;
  const int K = p.K[op], NJ = p.NJ[op], Wc = p.Wc[op];
  const long long total = (long long)kC * K * kRowFloats;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i & 3), js = (int)((i >> 2) & 15);
    const long long ck = i >> 6;
    const int k = (int)(ck % K), c = (int)(ck / K);
    const int j = c * Wc + js;
    p.dst[op][i] = (js < Wc && j < NJ) ? pack_src(p, op, k, j, g) : 0.f;
  }
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// __syncthreads first: the .aligned barrier needs converged warps (epilogues before it diverge)
__device__ __forceinline__ void cluster_arrive() {
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

struct Segs { const float* a0; int K0; const float* a1; };   // transposed [K][64] blocks, a1 after K0 rows

struct Ctx {
  float *As, *Ws, *red;     // shared memory: A ring, W ring, split-K partials
  int tid, rg, jl, crank;
};

__device__ __forceinline__ void issue_w(const Ctx& c, const OpDesc& d, int chunk) {
  const int k0 = chunk * kKC, nrows = min(kKC, d.K - k0);
  const float* src = d.w + ((size_t)c.crank * d.K + k0) * kRowFloats;
  float* dst = c.Ws + (chunk % kNS) * kStageFloats;
  for (int p = c.tid; p < nrows * 16; p += kThreads) cp_async16(dst + p * 4, src + p * 4);
}
__device__ __forceinline__ void issue_a(const Ctx& c, const OpDesc& d, const Segs& s, int chunk) {
  const int k0 = chunk * kKC, nrows = min(kKC, d.K - k0);
  float* dst = c.As + (chunk % kNS) * kStageFloats;
  for (int p = c.tid; p < nrows * 16; p += kThreads) {
    const int k = k0 + (p >> 4);
    const float* src = (k < s.K0 ? s.a0 + (size_t)k * kRowFloats : s.a1 + (size_t)(k - s.K0) * kRowFloats) + (p & 15) * 4;
    cp_async16(dst + p * 4, src);
  }
}
// weights of the first ring stages: independent of the other CTAs, requested before the barrier wait
__device__ __forceinline__ void prefetch_w(const Ctx& c, const OpDesc& d) {
  const int nch = (d.K + kKC - 1) / kKC;
  for (int s = 0; s < kNS - 1 && s < nch; ++s) issue_w(c, d, s);
}
__device__ __forceinline__ void fma_rows(const float* __restrict__ As, const float* __restrict__ Ws,
                                         int nrows, int ks, int KS, int js, int rg, float (&acc)[4][4]) {
#pragma unroll 4
  for (int k = ks; k < nrows; k += KS) {
    const float4 a = *reinterpret_cast<const float4*>(As + k * kRowFloats + rg * 4);
    const float4 b = *reinterpret_cast<const float4*>(Ws + k * kRowFloats + js * 4);
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int g = 0; g < 4; ++g) acc[r][g] = fmaf(av[r], bv[g], acc[r][g]);
  }
}
// split-K partials -> owner threads (ks == 0); returns true for threads that own a result
__device__ __forceinline__ bool reduce_splitk(const Ctx& c, const OpDesc& d, int js, int ks, float (&acc)[4][4]) {
  if (d.KS > 1) {
    float4* mine = reinterpret_cast<float4*>(c.red + c.tid * 16);
#pragma unroll
    for (int r = 0; r < 4; ++r) mine[r] = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    __syncthreads();
    if (ks == 0) {
      for (int q = 1; q < d.KS; ++q) {
        const float4* o = reinterpret_cast<const float4*>(c.red + ((q * d.WP + js) * 16 + c.rg) * 16);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float4 v = o[r];
          acc[r][0] += v.x; acc[r][1] += v.y; acc[r][2] += v.z; acc[r][3] += v.w;
        }
      }
    }
  }
  __syncthreads();    // ring stages and `red` are free for the next op
  return ks == 0 && js < d.Wc && (c.crank * d.Wc + js) < d.NJ;
}
// streamed op: A from the transposed scratch blocks.  prefetch_w(d) must have been called.
__device__ __forceinline__ bool op_run(const Ctx& c, const OpDesc& d, const Segs& s, float (&acc)[4][4], int* jglob) {
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int g = 0; g < 4; ++g) acc[r][g] = 0.f;
  const int js = c.jl & (d.WP - 1), ks = c.jl / d.WP;
  const int nch = (d.K + kKC - 1) / kKC;
  for (int st = 0; st < kNS - 1; ++st) {
    if (st < nch) issue_a(c, d, s, st);
    cp_async_commit();
  }
  for (int ch = 0; ch < nch; ++ch) {
    cp_async_wait<kNS - 2>();
    __syncthreads();
    const int nxt = ch + kNS - 1;
    if (nxt < nch) { issue_w(c, d, nxt); issue_a(c, d, s, nxt); }
    cp_async_commit();
    if (ks < d.KS)
      fma_rows(c.As + (ch % kNS) * kStageFloats, c.Ws + (ch % kNS) * kStageFloats, min(kKC, d.K - ch * kKC),
               ks, d.KS, js, c.rg, acc);
  }
  cp_async_wait<0>();
  *jglob = c.crank * d.Wc + js;
  return reduce_splitk(c, d, js, ks, acc);
}
// in-place op: the caller has written A as [K][64] at the start of the A ring (K <= kMaxSmallK);
// the whole weight image is loaded behind it.  Includes the barrier that publishes the caller's A.
__device__ __forceinline__ bool op_small(const Ctx& c, const OpDesc& d, float (&acc)[4][4], int* jglob) {
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int g = 0; g < 4; ++g) acc[r][g] = 0.f;
  const int js = c.jl & (d.WP - 1), ks = c.jl / d.WP;
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  if (ks < d.KS) fma_rows(c.As, c.Ws, d.K, ks, d.KS, js, c.rg, acc);
  *jglob = c.crank * d.Wc + js;
  return reduce_splitk(c, d, js, ks, acc);
}
__device__ __forceinline__ void load_w_small(const Ctx& c, const OpDesc& d) {
  const float* src = d.w + (size_t)c.crank * d.K * kRowFloats;
  for (int p = c.tid; p < d.K * 16; p += kThreads) cp_async16(c.Ws + p * 4, src + p * 4);
}
__device__ __forceinline__ void st4(float* p, float a, float b, float c_, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c_, d);
}

// ------------------------------------------------------------------------------------------------
struct FwdArgs {
  OpDesc emb, gru, q1, q2;
  int L; long long B;
  int Be, Hi, S, A, act;
  float min_std;
  const float *init_state, *init_belief, *actions, *nonterm, *eps_post, *PE;
  const float *b_sa, *b_ih, *b_hh, *b_q2;
  float *beliefs, *post_s, *post_m, *post_sd;
  float* scratch;          // per cluster: xT [Be][64] | hT [2][Be][64] | hqT [Hi][64]
};
__host__ __device__ inline size_t fwd_scratch_floats(int Be, int Hi) { return (size_t)(3 * Be + Hi) * kRowFloats; }

__global__ void __launch_bounds__(kThreads, 1) observe_fwd_kernel(const __grid_constant__ FwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  Ctx c;
  c.As = smem; c.Ws = smem + kNS * kStageFloats; c.red = smem + 2 * kNS * kStageFloats;
  c.tid = threadIdx.x; c.rg = c.tid & 15; c.jl = c.tid >> 4; c.crank = (int)cluster_ctarank();
  const int Be = a.Be, Hi = a.Hi, S = a.S, Ad = a.A, SA = a.S + a.A;
  const long long B = a.B, row0 = (long long)(blockIdx.x / kC) * kR;
  const int nvalid = (int)min((long long)kR, B - row0);
  float* xT = a.scratch + (size_t)(blockIdx.x / kC) * fwd_scratch_floats(Be, Hi);
  float* hT = xT + (size_t)Be * kRowFloats;
  float* hqT = hT + (size_t)2 * Be * kRowFloats;
  const int r0 = c.rg * 4;
  // h_0^T = init_belief^T (columns interleaved over the CTAs)
  for (int i = c.crank * kThreads + c.tid; i < Be * kR; i += kC * kThreads) {
    const int j = i >> 6, r = i & 63;
    hT[i] = r < nvalid ? a.init_belief[(row0 + r) * Be + j] : 0.f;
  }
  cluster_arrive();
  float acc[4][4];
  int jg;
  for (int t = 0; t < a.L; ++t) {
    const int par = t & 1;
    const long long trow = (long long)t * B + row0;
    // ---------------------------------------------------------------- P1 embed
    load_w_small(c, a.emb);
    cluster_wait();
    {
      const float* sp = t == 0 ? a.init_state + row0 * S : a.post_s + ((long long)(t - 1) * B + row0) * S;
      for (int i = c.tid; i < SA * kR; i += kThreads) {
        const int k = i >> 6, r = i & 63;
        float v = 0.f;
        if (r < nvalid) {
          if (k < S) {
            v = __ldcg(sp + (long long)r * S + k);
            if (a.nonterm) v *= a.nonterm[trow + r];
          } else v = a.actions[(trow + r) * Ad + (k - S)];
        }
        c.As[i] = v;
      }
    }
    if (op_small(c, a.emb, acc, &jg)) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int col = 4 * jg + g;
        if (col < Be) {
          const float b = a.b_sa[col];
          st4(xT + (size_t)col * kRowFloats + r0, act_fwd(a.act, acc[0][g] + b), act_fwd(a.act, acc[1][g] + b),
              act_fwd(a.act, acc[2][g] + b), act_fwd(a.act, acc[3][g] + b));
        }
      }
    }
    cluster_arrive();
    // ---------------------------------------------------------------- P2 GRU
    prefetch_w(c, a.gru);
    cluster_wait();
    const float* hprevT = hT + (size_t)par * Be * kRowFloats;
    float* hnewT = hT + (size_t)(par ^ 1) * Be * kRowFloats;
    if (op_run(c, a.gru, Segs{xT, Be, hprevT}, acc, &jg)) {
      const int j = jg;
      const float4 h4 = ldcg4(hprevT + (size_t)j * kRowFloats + r0);
      const float hp[4] = {h4.x, h4.y, h4.z, h4.w};
      const float br = a.b_ih[j] + a.b_hh[j], bz = a.b_ih[Be + j] + a.b_hh[Be + j];
      const float bin = a.b_ih[2 * Be + j], bhn = a.b_hh[2 * Be + j];
      float hn[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float rr = sigmoidf_(acc[r][0] + br), z = sigmoidf_(acc[r][1] + bz);
        const float n = tanhf(acc[r][2] + bin + rr * (acc[r][3] + bhn));
        hn[r] = (1.f - z) * n + z * hp[r];
        if (r0 + r < nvalid) a.beliefs[(trow + r0 + r) * Be + j] = hn[r];
      }
      st4(hnewT + (size_t)j * kRowFloats + r0, hn[0], hn[1], hn[2], hn[3]);
    }
    cluster_arrive();
    // ---------------------------------------------------------------- P3 posterior hidden
    prefetch_w(c, a.q1);
    cluster_wait();
    if (op_run(c, a.q1, Segs{hnewT, Be, nullptr}, acc, &jg)) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int col = 4 * jg + g;
        if (col < Hi) {
          float v[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int rl = min(r0 + r, nvalid - 1);
            v[r] = act_fwd(a.act, acc[r][g] + a.PE[(trow + rl) * Hi + col]);
          }
          st4(hqT + (size_t)col * kRowFloats + r0, v[0], v[1], v[2], v[3]);
        }
      }
    }
    cluster_arrive();
    // ---------------------------------------------------------------- P4 posterior output + sample
    prefetch_w(c, a.q2);
    cluster_wait();
    if (op_run(c, a.q2, Segs{hqT, Hi, nullptr}, acc, &jg)) {
      const int j = jg;
      const float bm = a.b_q2[j], bs = a.b_q2[S + j];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (r0 + r < nvalid) {
          const long long o = (trow + r0 + r) * S + j;
          const float m = acc[r][0] + bm, sd = softplusf_(acc[r][1] + bs) + a.min_std;
          a.post_m[o] = m;
          a.post_sd[o] = sd;
          a.post_s[o] = m + sd * a.eps_post[o];
        }
      }
    }
    cluster_arrive();
  }
  cluster_wait();
}

// ------------------------------------------------------------------------------------------------
struct BwdArgs {
  OpDesc b1, b2, b3, b4;
  int L; long long B;
  int Be, Bep, Hi, S, A, act;
  const float *nonterm, *eps_post, *g_post_s, *g_post_m, *g_post_sd;
  const float *preq, *hq, *x, *gi, *gh, *init_belief, *beliefs, *Gtot;    // (L*B, .) row-major
  float *tdpreq, *tdhq, *tdgi, *tdgh, *tdx;                                 // (L*B, .) row-major, written
  float *d_actions, *d_init_state, *d_init_belief;
  float* scratch;   // per cluster: dhqT [Hi][64] | planesT [4 Be][64] | dxT [Be][64] | czT [Be][64] | cbT [Be][64] | csT [S][64]
};
__host__ __device__ inline size_t bwd_scratch_floats(int Be, int Hi, int S) { return (size_t)(Hi + 7 * Be + S) * kRowFloats; }

__global__ void __launch_bounds__(kThreads, 1) observe_bwd_kernel(const __grid_constant__ BwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  Ctx c;
  c.As = smem; c.Ws = smem + kNS * kStageFloats; c.red = smem + 2 * kNS * kStageFloats;
  c.tid = threadIdx.x; c.rg = c.tid & 15; c.jl = c.tid >> 4; c.crank = (int)cluster_ctarank();
  const int Be = a.Be, Bep = a.Bep, Hi = a.Hi, S = a.S, Ad = a.A;
  const long long B = a.B, row0 = (long long)(blockIdx.x / kC) * kR;
  const int nvalid = (int)min((long long)kR, B - row0);
  float* dhqT = a.scratch + (size_t)(blockIdx.x / kC) * bwd_scratch_floats(Be, Hi, S);
  float* planesT = dhqT + (size_t)Hi * kRowFloats;
  float* dxT = planesT + (size_t)4 * Be * kRowFloats;
  float* czT = dxT + (size_t)Be * kRowFloats;
  float* cbT = czT + (size_t)Be * kRowFloats;
  float* csT = cbT + (size_t)Be * kRowFloats;
  const int r0 = c.rg * 4;
  cluster_arrive();
  float acc[4][4];
  int jg;
  for (int i = 0; i < a.L; ++i) {
    const int t = a.L - 1 - i;
    const bool first = i == 0;
    const long long trow = (long long)t * B + row0;
    // ---------------------------------------------------------------- Q1 posterior output backward
    load_w_small(c, a.b1);
    cluster_wait();
    for (int e = c.tid; e < S * kR; e += kThreads) {
      const int j = e >> 6, r = e & 63;
      const int rl = min(r, nvalid - 1);
      const long long o = (trow + rl) * S + j;
      float gs = a.g_post_s ? a.g_post_s[o] : 0.f;
      if (!first) gs += __ldcg(csT + e);
      const float dm = gs + (a.g_post_m ? a.g_post_m[o] : 0.f);
      const float dsd = gs * a.eps_post[o] + (a.g_post_sd ? a.g_post_sd[o] : 0.f);
      const float draw = dsd * softplus_gradf_(a.preq[(trow + rl) * 2 * S + S + j]);
      c.As[e] = dm;
      c.As[S * kR + e] = draw;
      if (c.crank == 0 && r < nvalid) {
        a.tdpreq[(trow + r) * 2 * S + j] = dm;
        a.tdpreq[(trow + r) * 2 * S + S + j] = draw;
      }
    }
    if (op_small(c, a.b1, acc, &jg)) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int col = 4 * jg + g;
        if (col < Hi) {
          float v[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int rl = min(r0 + r, nvalid - 1);
            v[r] = acc[r][g] * act_bwd_from_out(a.act, a.hq[(trow + rl) * Hi + col]);
            if (r0 + r < nvalid) a.tdhq[(trow + r0 + r) * Hi + col] = v[r];
          }
          st4(dhqT + (size_t)col * kRowFloats + r0, v[0], v[1], v[2], v[3]);
        }
      }
    }
    cluster_arrive();
    // ---------------------------------------------------------------- Q2 belief gradient + GRU gates
    prefetch_w(c, a.b2);
    cluster_wait();
    if (op_run(c, a.b2, Segs{dhqT, Hi, nullptr}, acc, &jg)) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int col = 4 * jg + g;
        if (col < Be) {
          float cb[4] = {0.f, 0.f, 0.f, 0.f};
          if (!first) {
            const float4 q = ldcg4(cbT + (size_t)col * kRowFloats + r0);
            cb[0] = q.x; cb[1] = q.y; cb[2] = q.z; cb[3] = q.w;
          }
          float dr[4], dz_[4], dn_[4], dnr[4], cz[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int rl = min(r0 + r, nvalid - 1);
            const long long row = trow + rl;
            const long long o3 = row * 3 * Be;
            const float G = a.Gtot[row * Be + col] + cb[r] + acc[r][g];
            const float ghn = a.gh[o3 + 2 * Be + col];
            const float rr = sigmoidf_(a.gi[o3 + col] + a.gh[o3 + col]);
            const float z = sigmoidf_(a.gi[o3 + Be + col] + a.gh[o3 + Be + col]);
            const float n = tanhf(a.gi[o3 + 2 * Be + col] + rr * ghn);
            const float h = t == 0 ? a.init_belief[(row0 + rl) * Be + col]
                                   : a.beliefs[((long long)(t - 1) * B + row0 + rl) * Be + col];
            const float dn = G * (1.f - z), dzz = G * (h - n);
            const float dpn = dn * (1.f - n * n);
            const float dpr = dpn * ghn * rr * (1.f - rr);
            const float dpz = dzz * z * (1.f - z);
            dr[r] = dpr; dz_[r] = dpz; dn_[r] = dpn; dnr[r] = dpn * rr; cz[r] = G * z;
            if (r0 + r < nvalid) {
              float* gi_o = a.tdgi + o3;
              float* gh_o = a.tdgh + o3;
              gi_o[col] = dpr; gh_o[col] = dpr;
              gi_o[Be + col] = dpz; gh_o[Be + col] = dpz;
              gi_o[2 * Be + col] = dpn; gh_o[2 * Be + col] = dpn * rr;
            }
          }
          st4(planesT + (size_t)col * kRowFloats + r0, dr[0], dr[1], dr[2], dr[3]);
          st4(planesT + (size_t)(Be + col) * kRowFloats + r0, dz_[0], dz_[1], dz_[2], dz_[3]);
          st4(planesT + (size_t)(2 * Be + col) * kRowFloats + r0, dn_[0], dn_[1], dn_[2], dn_[3]);
          st4(planesT + (size_t)(3 * Be + col) * kRowFloats + r0, dnr[0], dnr[1], dnr[2], dnr[3]);
          st4(czT + (size_t)col * kRowFloats + r0, cz[0], cz[1], cz[2], cz[3]);
        }
      }
    }
    cluster_arrive();
    // ---------------------------------------------------------------- Q3 GRU input / hidden dgrad
    prefetch_w(c, a.b3);
    cluster_wait();
    if (op_run(c, a.b3, Segs{planesT, 4 * Be, nullptr}, acc, &jg)) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int col = 4 * jg + g;
        if (col < Bep) {
          if (col < Be) {
            float v[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const int rl = min(r0 + r, nvalid - 1);
              v[r] = acc[r][g] * act_bwd_from_out(a.act, a.x[(trow + rl) * Be + col]);
              if (r0 + r < nvalid) a.tdx[(trow + r0 + r) * Be + col] = v[r];
            }
            st4(dxT + (size_t)col * kRowFloats + r0, v[0], v[1], v[2], v[3]);
          }
        } else {
          const int cc = col - Bep;
          if (cc < Be) {
            const float4 q = ldcg4(czT + (size_t)cc * kRowFloats + r0);
            const float v[4] = {acc[0][g] + q.x, acc[1][g] + q.y, acc[2][g] + q.z, acc[3][g] + q.w};
            st4(cbT + (size_t)cc * kRowFloats + r0, v[0], v[1], v[2], v[3]);
            if (t == 0 && a.d_init_belief) {
#pragma unroll
              for (int r = 0; r < 4; ++r)
                if (r0 + r < nvalid) a.d_init_belief[(row0 + r0 + r) * Be + cc] = v[r];
            }
          }
        }
      }
    }
    cluster_arrive();
    // ---------------------------------------------------------------- Q4 embed dgrad
    prefetch_w(c, a.b4);
    cluster_wait();
    if (op_run(c, a.b4, Segs{dxT, Be, nullptr}, acc, &jg)) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int col = 4 * jg + g;
        if (col < S) {
          float v[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int rl = min(r0 + r, nvalid - 1);
            v[r] = acc[r][g] * (a.nonterm ? a.nonterm[trow + rl] : 1.f);
            if (t == 0 && a.d_init_state && r0 + r < nvalid) a.d_init_state[(row0 + r0 + r) * S + col] = v[r];
          }
          st4(csT + (size_t)col * kRowFloats + r0, v[0], v[1], v[2], v[3]);
        } else if (col < S + Ad && a.d_actions) {
#pragma unroll
          for (int r = 0; r < 4; ++r)
            if (r0 + r < nvalid) a.d_actions[(trow + r0 + r) * Ad + (col - S)] = acc[r][g];
        }
      }
    }
    cluster_arrive();
  }
  cluster_wait();
}

}  // namespace obs
}  // namespace f32
}  // namespace bd

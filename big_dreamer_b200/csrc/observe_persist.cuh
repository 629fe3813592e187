// Persistent cluster kernels for the RSSM observe pass (TransitionModel.forward with observations,
// src/models.py:239-271) on small batches (BASELINE configs[3]: 49 steps x 50 rows).
//
// The per-step path launches ~10 (forward) / ~20 (backward) 50-row kernels per time step: 1 400
// dependent launches, each an L2 round trip.  Here ONE thread-block cluster of 16 CTAs walks all L
// steps of a 64-row chunk (one cluster per chunk of rows; rows are independent):
//   * every CTA owns a slice of each layer's output columns, for all 64 rows;
//   * layer inputs travel between CTAs as TRANSPOSED [K][64] fp32 blocks in an L2-resident scratch
//     (one contiguous block per layer, streamed into shared memory with 16-byte cp.async.cg through a
//     double-buffered ring of 104 k-rows per stage -- K = 200 layers take two hand-offs; four stages of
//     64 rows measured 6 % slower per step), the weights as per-CTA pre-packed [K][16 slots][4] images
//     (pack_ops_kernel; constant over the steps, so their first ring stages are requested BEFORE
//     the cluster barrier that waits for the other CTAs' activations);
//   * phases are separated by the hardware cluster barrier (arrive.release / wait.acquire).
// Outputs are organised in "slot groups" of four: the four pre-activations of one GRU unit (r, z,
// gi_n, gh_n -- the x and h projections are ONE contraction over [x ; h]), (mu_j, raw_j) of the
// posterior, or four adjacent output columns of a plain layer.  Compute thread tile: 8 rows x 8 slots
// (4 LDS.128 per 32 packed FFMA2 = 1 B of shared-memory traffic per FMA, the balance point of
// 128 B/clk against 128 FMA/clk; a 4 x 4 tile measured 90 cycles per k-row, LDS-bound).  A 64 x 64 output
// tile only needs 64 such threads, so the k rows of every ring stage are dealt round-robin to 4-32
// groups of threads ("k-groups") whose partial sums meet in shared memory; the epilogue then runs
// one (row, slot group) item per thread over all 256 threads, with its constant inputs requested
// before the cluster barrier.  In the 16-bit precision modes the two big contractions (P2, Q3) run on
// TF32 mma.sync m16n8k8 over the same shared-memory blocks (op_compute_tc: 4x the FFMA2 rate measured
// on B200; everything else, including all state, stays fp32).
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
constexpr int kKC = 104;        // k rows per ring stage
constexpr int kNS = 2;          // ring stages
constexpr int kThreads = 256;
constexpr int kRowFloats = 64;  // floats per k row of an A block (64 rows) and of a W image (16 x 4)
constexpr int kLds = 72;        // floats between k rows in SHARED memory: 64 + 8 keeps the mma.sync fragment loads
                                // (4 k rows x 8 columns per instruction) on 32 different banks
constexpr int kStageFloats = kKC * kLds;
constexpr int kRedFloats = 64 * 64 * 4;   // k-group partial sums: (KG * WP) x 64 rows x 4 slots, KG * WP <= 64
constexpr int kSmemBytes = (2 * kNS * kStageFloats + kRedFloats) * (int)sizeof(float);   // 212 992
constexpr int kMaxSmallK = kNS * kKC;   // in-place operands (embed input, d_preq) fit the A ring

enum OpId { OP_EMB = 0, OP_GRU, OP_Q1F, OP_Q2F, OP_B1, OP_B2, OP_B3, OP_B4, OP_COUNT };

struct OpDesc {
  const float* w;   // packed images: CTA c's image at w + c * K * 64
  int K, NJ;        // contraction length, slot groups over the whole cluster
  int Wc, WP;       // slot groups per CTA, rounded up to a power of two
  int TS, KG, tgs;  // slots per compute thread (4 | 8), k-groups, log2(threads per k-group)
  int tc;           // 1: TF32 mma.sync contraction (16-bit precision modes, WP = 8 | 16, K % 8 == 0; KG = 4)
};
inline OpDesc make_op(int K, int NJ) {
  OpDesc d{};
  d.K = K; d.NJ = NJ;
  d.Wc = (NJ + kC - 1) / kC;
  d.WP = 1;
  while (d.WP < d.Wc) d.WP <<= 1;
  d.TS = d.WP >= 2 ? 8 : 4;
  const int tg = 8 * (d.WP * 4 / d.TS);      // 8 row octets x slot tiles
  d.KG = kThreads / tg;
  d.tgs = 0;
  while ((1 << d.tgs) < tg) ++d.tgs;
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
  float *As, *Ws, *red;     // shared memory: A ring, W ring, k-group partials
  int tid, crank, nvalid;
};

__device__ __forceinline__ void issue_w(const Ctx& c, const OpDesc& d, int chunk) {
  const int k0 = chunk * kKC, nrows = min(kKC, d.K - k0);
  const float* src = d.w + ((size_t)c.crank * d.K + k0) * kRowFloats;
  float* dst = c.Ws + (chunk % kNS) * kStageFloats;
  for (int p = c.tid; p < nrows * 16; p += kThreads) cp_async16(dst + (p >> 4) * kLds + (p & 15) * 4, src + p * 4);
}
__device__ __forceinline__ void issue_a(const Ctx& c, const OpDesc& d, const Segs& s, int chunk) {
  const int k0 = chunk * kKC, nrows = min(kKC, d.K - k0);
  float* dst = c.As + (chunk % kNS) * kStageFloats;
  for (int p = c.tid; p < nrows * 16; p += kThreads) {
    const int k = k0 + (p >> 4);
    const float* src = (k < s.K0 ? s.a0 + (size_t)k * kRowFloats : s.a1 + (size_t)(k - s.K0) * kRowFloats) + (p & 15) * 4;
    cp_async16(dst + (p >> 4) * kLds + (p & 15) * 4, src);
  }
}
// weights of the first ring stages: independent of the other CTAs, requested before the barrier wait
__device__ __forceinline__ void prefetch_w(const Ctx& c, const OpDesc& d) {
  const int nch = (d.K + kKC - 1) / kKC;
  for (int s = 0; s < kNS - 1 && s < nch; ++s) issue_w(c, d, s);
}
__device__ __forceinline__ void load_w_small(const Ctx& c, const OpDesc& d) {
  const float* src = d.w + (size_t)c.crank * d.K * kRowFloats;
  for (int p = c.tid; p < d.K * 16; p += kThreads) cp_async16(c.Ws + (p >> 4) * kLds + (p & 15) * 4, src + p * 4);
}

// Software-pipelined: the operands of k-row k + KG are requested before the FMAs of k-row k (the
// plain loop left the 4 LDS.128 -> 64 FFMA dependency exposed every iteration: with two warps per
// scheduler and the LDS pipe as busy as the FMA pipe it ran at half the issue rate).
template <int TS>
__device__ __forceinline__ void fma_rows(const float* __restrict__ ap, const float* __restrict__ bp,
                                         int nrows, int kg, int KG, float (&acc)[8][8]) {
  int k = kg;
  if (k >= nrows) return;
  float4 a0 = *reinterpret_cast<const float4*>(ap + k * kLds);
  float4 a1 = *reinterpret_cast<const float4*>(ap + k * kLds + 32);
  float4 b0 = *reinterpret_cast<const float4*>(bp + k * kLds);
  float4 b1 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (TS == 8) b1 = *reinterpret_cast<const float4*>(bp + k * kLds + 4);
  for (;;) {
    const int kn = k + KG;
    const bool more = kn < nrows;
    const int kl = more ? kn : k;                 // clamped: the loads are always issued
    const float4 na0 = *reinterpret_cast<const float4*>(ap + kl * kLds);
    const float4 na1 = *reinterpret_cast<const float4*>(ap + kl * kLds + 32);
    const float4 nb0 = *reinterpret_cast<const float4*>(bp + kl * kLds);
    float4 nb1 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (TS == 8) nb1 = *reinterpret_cast<const float4*>(bp + kl * kLds + 4);
    const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int g = 0; g < TS; g += 2) ffma2(acc[r][g], acc[r][g + 1], av[r], bv[g], bv[g + 1]);
    if (!more) break;
    a0 = na0; a1 = na1; b0 = nb0; b1 = nb1;
    k = kn;
  }
}

// ---- TF32 tensor-core variant of the contraction (precision modes other than fp32) ----------------
// Legacy mma.sync m16n8k8 issues every 8 cycles per scheduler on sm_100a = 512 MAC/clk/SM, 4x the packed
// FFMA2 rate (scripts/mma_tf32_bench.cu); the operands stay the fp32 blocks of the fp32 path (rounded to
// TF32 with cvt.rna at fragment load), so nothing else changes.  Warp (kg, half): k-group kg takes every
// fourth k8 step, `half` selects the lower / upper half of the CTA's slots; 4 m tiles x NT n tiles.
template <int NT>
__device__ __forceinline__ void mma_rows(const float* __restrict__ As, const float* __restrict__ Ws, int nrows,
                                         int kg, int n0, int lane, int nvalid, float (&acc)[4][NT][4]) {
  const int g = lane >> 2, t = lane & 3;
  for (int k0 = kg * 8; k0 < nrows; k0 += 32) {
    uint32_t af[4][4], bf[NT][2];
    const float* ab = As + (k0 + t) * kLds + g;
    const float* bb = Ws + (k0 + t) * kLds + n0 + g;
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      af[mt][0] = f32_to_tf32(ab[16 * mt]);
      af[mt][1] = f32_to_tf32(ab[16 * mt + 8]);
      af[mt][2] = f32_to_tf32(ab[4 * kLds + 16 * mt]);
      af[mt][3] = f32_to_tf32(ab[4 * kLds + 16 * mt + 8]);
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      bf[nt][0] = f32_to_tf32(bb[8 * nt]);
      bf[nt][1] = f32_to_tf32(bb[4 * kLds + 8 * nt]);
    }
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      if (16 * mt < nvalid) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
          asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                       : "+f"(acc[mt][nt][0]), "+f"(acc[mt][nt][1]), "+f"(acc[mt][nt][2]), "+f"(acc[mt][nt][3])
                       : "r"(af[mt][0]), "r"(af[mt][1]), "r"(af[mt][2]), "r"(af[mt][3]), "r"(bf[nt][0]), "r"(bf[nt][1]));
      }
    }
  }
}
// streamed contraction on mma.sync; same contract as op_compute<false> (partials in c.red, KG = 4)
template <int NT>
__device__ __forceinline__ void op_compute_tc(const Ctx& c, const OpDesc& d, const Segs& s) {
  float acc[4][NT][4];
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
  const int warp = c.tid >> 5, lane = c.tid & 31;
  const int kg = warp >> 1, n0 = (warp & 1) * 8 * NT;
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
    mma_rows<NT>(c.As + (ch % kNS) * kStageFloats, c.Ws + (ch % kNS) * kStageFloats, min(kKC, d.K - ch * kKC), kg,
                 n0, lane, c.nvalid, acc);
  }
  cp_async_wait<0>();
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int slot = n0 + 8 * nt + 2 * t;
      float* r0p = c.red + ((size_t)((kg * d.WP + (slot >> 2)) * kR + 16 * mt + g)) * 4 + (slot & 3);
      *reinterpret_cast<float2*>(r0p) = make_float2(acc[mt][nt][0], acc[mt][nt][1]);
      *reinterpret_cast<float2*>(r0p + 8 * 4) = make_float2(acc[mt][nt][2], acc[mt][nt][3]);
    }
  __syncthreads();
}

// One contraction of the CTA's column slice: leaves the k-group partial sums in c.red
// ([(kg * WP + js) * 64 + row] float4 = the four slots of slot group js) and ends with a
// __syncthreads.  SMALL: the caller has written A as [K][64] at the start of the A ring
// (K <= kMaxSmallK) and requested the whole weight image with load_w_small; otherwise A streams
// from the transposed scratch blocks and prefetch_w(d) has been called.
template <bool SMALL>
__device__ __forceinline__ void op_compute(const Ctx& c, const OpDesc& d, const Segs& s) {
  float acc[8][8];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int g = 0; g < 8; ++g) acc[r][g] = 0.f;
  const int kg = c.tid >> d.tgs, u = c.tid & ((1 << d.tgs) - 1);
  const int rr = u & 7, sg = u >> 3;
  // rows of this thread: 4 rr .. 4 rr + 3 and 32 + 4 rr .. 32 + 4 rr + 3, so that the eight lanes of a
  // quarter warp read 128 contiguous bytes per LDS.128 (8 consecutive rows per thread would put
  // them 32 bytes apart: a 2-way bank conflict on every A load)
  const bool active = rr * 4 < c.nvalid;
  const int aoff = rr * 4, boff = sg * d.TS;
  if (SMALL) {
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    if (active) {
      if (d.TS == 8) fma_rows<8>(c.As + aoff, c.Ws + boff, d.K, kg, d.KG, acc);
      else fma_rows<4>(c.As + aoff, c.Ws + boff, d.K, kg, d.KG, acc);
    }
  } else {
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
      if (active) {
        const float* ap = c.As + (ch % kNS) * kStageFloats + aoff;
        const float* bp = c.Ws + (ch % kNS) * kStageFloats + boff;
        const int nrows = min(kKC, d.K - ch * kKC);
        if (d.TS == 8) fma_rows<8>(ap, bp, nrows, kg, d.KG, acc);
        else fma_rows<4>(ap, bp, nrows, kg, d.KG, acc);
      }
    }
    cp_async_wait<0>();
  }
  if (active) {
    float4* red4 = reinterpret_cast<float4*>(c.red);
    const int js0 = sg * (d.TS >> 2);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int row = aoff + (r < 4 ? r : 28 + r);
      red4[(kg * d.WP + js0) * kR + row] = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
      if (d.TS == 8)
        red4[(kg * d.WP + js0 + 1) * kR + row] = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
    }
  }
  __syncthreads();
}
// the four slots of slot group js at one row, summed over the k-groups
__device__ __forceinline__ float4 red_sum(const Ctx& c, const OpDesc& d, int js, int row) {
  const float4* red4 = reinterpret_cast<const float4*>(c.red);
  float4 v = red4[js * kR + row];
  for (int kg = 1; kg < d.KG; ++kg) {
    const float4 w = red4[(kg * d.WP + js) * kR + row];
    v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
  }
  return v;
}
// epilogue item of a thread: (row, slot group); `it` < 64 * WP
__device__ __forceinline__ bool item_of(const Ctx& c, const OpDesc& d, int it, int* row, int* js, int* jg) {
  *row = it & 63; *js = it >> 6;
  *jg = c.crank * d.Wc + *js;
  return it < kR * d.WP && *row < c.nvalid && *js < d.Wc && *jg < d.NJ;
}
// four consecutive columns of a row-major row (vec: 16-byte aligned and ncols % 4 == 0)
__device__ __forceinline__ float4 ldrow4(const float* rowp, int col, int ncols, bool vec) {
  if (vec) return *reinterpret_cast<const float4*>(rowp + col);
  float4 v;
  v.x = rowp[min(col, ncols - 1)]; v.y = rowp[min(col + 1, ncols - 1)];
  v.z = rowp[min(col + 2, ncols - 1)]; v.w = rowp[min(col + 3, ncols - 1)];
  return v;
}
__device__ __forceinline__ void strow4(float* rowp, int col, int ncols, bool vec, float4 v) {
  if (vec) { *reinterpret_cast<float4*>(rowp + col) = v; return; }
  if (col < ncols) rowp[col] = v.x;
  if (col + 1 < ncols) rowp[col + 1] = v.y;
  if (col + 2 < ncols) rowp[col + 2] = v.z;
  if (col + 3 < ncols) rowp[col + 3] = v.w;
}

// ------------------------------------------------------------------------------------------------
#define OBS_T(i) do { if (a.prof) tm[i] = clock64(); } while (0)
#define OBS_ACC(ph) do { if (a.prof) { pw[ph] += tm[1] - tm[0]; po[ph] += tm[2] - tm[1]; pe[ph] += tm[3] - tm[2]; } } while (0)
#define OBS_REPORT(name) do { if (a.prof && c.tid == 0 && blockIdx.x < kC && (c.crank == 0 || c.crank == kC - 1)) \
    for (int ph = 0; ph < 4; ++ph) printf("%s rank %d phase %d: wait %lld op %lld epi %lld cycles/step\n", name, c.crank, ph, \
           pw[ph] / a.L, po[ph] / a.L, pe[ph] / a.L); } while (0)

struct FwdArgs {
  OpDesc emb, gru, q1, q2;
  int L; long long B;
  int Be, Hi, S, A, act;
  float min_std;
  const float *init_state, *init_belief, *actions, *nonterm, *eps_post, *PE;
  const float *b_sa, *b_ih, *b_hh, *b_q2;
  float *beliefs, *post_s, *post_m, *post_sd;
  float* scratch;          // per cluster: xT [Be][64] | hT [2][Be][64] | hqT [Hi][64]
  int prof;                // debug: print per-phase cycle counts of cluster ranks 0 and 15 (BD_OBS_PROF=1)
};
__host__ __device__ inline size_t fwd_scratch_floats(int Be, int Hi) { return (size_t)(3 * Be + Hi) * kRowFloats; }

__global__ void __launch_bounds__(kThreads, 1) observe_fwd_kernel(const __grid_constant__ FwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  Ctx c;
  c.As = smem; c.Ws = smem + kNS * kStageFloats; c.red = smem + 2 * kNS * kStageFloats;
  c.tid = threadIdx.x; c.crank = (int)cluster_ctarank();
  const int Be = a.Be, Hi = a.Hi, S = a.S, Ad = a.A, SA = a.S + a.A;
  const long long B = a.B, row0 = (long long)(blockIdx.x / kC) * kR;
  c.nvalid = (int)min((long long)kR, B - row0);
  const int nvalid = c.nvalid;
  float* xT = a.scratch + (size_t)(blockIdx.x / kC) * fwd_scratch_floats(Be, Hi);
  float* hT = xT + (size_t)Be * kRowFloats;
  float* hqT = hT + (size_t)2 * Be * kRowFloats;
  const bool vecH = (Hi & 3) == 0;
  // h_0^T = init_belief^T (columns interleaved over the CTAs)
  for (int i = c.crank * kThreads + c.tid; i < Be * kR; i += kC * kThreads) {
    const int j = i >> 6, r = i & 63;
    hT[i] = r < nvalid ? a.init_belief[(row0 + r) * Be + j] : 0.f;
  }
  cluster_arrive();
  long long tm[4] = {0, 0, 0, 0}, pw[4] = {0, 0, 0, 0}, po[4] = {0, 0, 0, 0}, pe[4] = {0, 0, 0, 0};
  int row, js, jg;
  // GRU epilogue items of this thread (up to 4 units j for its row): the gate biases never change
  float gbias[4][4];
  bool gown[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    gown[q] = item_of(c, a.gru, c.tid + q * kThreads, &row, &js, &jg);
    const int j = gown[q] ? jg : 0;
    gbias[q][0] = a.b_ih[j] + a.b_hh[j];
    gbias[q][1] = a.b_ih[Be + j] + a.b_hh[Be + j];
    gbias[q][2] = a.b_ih[2 * Be + j];
    gbias[q][3] = a.b_hh[2 * Be + j];
  }
  for (int t = 0; t < a.L; ++t) {
    const int par = t & 1;
    const long long trow = (long long)t * B + row0;
    // ---------------------------------------------------------------- P1 embed
    OBS_T(0);
    load_w_small(c, a.emb);
    // in-place operand [s * nt ; a]^T: the action rows and the mask do not depend on the other CTAs
    const int rme = c.tid & 63;
    float ntr = 1.f;
    if (a.nonterm && rme < nvalid) ntr = a.nonterm[trow + rme];
    for (int i = S * kR + c.tid; i < SA * kR; i += kThreads)
      c.As[(i >> 6) * kLds + rme] = rme < nvalid ? a.actions[(trow + rme) * Ad + ((i >> 6) - S)] : 0.f;
    cluster_wait();
    OBS_T(1);
    {
      const float* sp = t == 0 ? a.init_state + row0 * S : a.post_s + ((long long)(t - 1) * B + row0) * S;
      const int rl = min(rme, nvalid - 1);
      for (int i0 = c.tid; i0 < S * kR; i0 += 8 * kThreads) {     // eight independent loads in flight
        float sv[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) sv[q] = __ldcg(sp + (long long)rl * S + min((i0 >> 6) + 4 * q, S - 1));
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int k = (i0 >> 6) + 4 * q;
          if (k < S) c.As[k * kLds + rme] = rme < nvalid ? sv[q] * ntr : 0.f;
        }
      }
    }
    op_compute<true>(c, a.emb, Segs{});
    OBS_T(2);
    if (item_of(c, a.emb, c.tid, &row, &js, &jg)) {
      const float4 v = red_sum(c, a.emb, js, row);
      const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int col = 4 * jg + g;
        if (col < Be) xT[(size_t)col * kRowFloats + row] = act_fwd(a.act, vv[g] + a.b_sa[col]);
      }
    }
    cluster_arrive();
    OBS_T(3);
    OBS_ACC(0);
    // ---------------------------------------------------------------- P2 GRU
    OBS_T(0);
    prefetch_w(c, a.gru);
    const float* hprevT = hT + (size_t)par * Be * kRowFloats;
    float* hnewT = hT + (size_t)(par ^ 1) * Be * kRowFloats;
    // h_{t-1} of this thread's items: written two phases ago (or by the prologue), so it can be
    // requested before the barrier
    float hprev[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      hprev[q] = 0.f;
      if (gown[q]) {
        item_of(c, a.gru, c.tid + q * kThreads, &row, &js, &jg);
        hprev[q] = __ldcg(hprevT + (size_t)jg * kRowFloats + row);
      }
    }
    cluster_wait();
    OBS_T(1);
    if (a.gru.tc) {
      if (a.gru.WP == 16) op_compute_tc<4>(c, a.gru, Segs{xT, Be, hprevT});
      else op_compute_tc<2>(c, a.gru, Segs{xT, Be, hprevT});
    } else {
      op_compute<false>(c, a.gru, Segs{xT, Be, hprevT});
    }
    OBS_T(2);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (gown[q]) {
        item_of(c, a.gru, c.tid + q * kThreads, &row, &js, &jg);
        const int j = jg;
        const float4 v = red_sum(c, a.gru, js, row);
        const float rr = sigmoidf_(v.x + gbias[q][0]);
        const float z = sigmoidf_(v.y + gbias[q][1]);
        const float n = tanhf(v.z + gbias[q][2] + rr * (v.w + gbias[q][3]));
        const float hn = (1.f - z) * n + z * hprev[q];
        a.beliefs[(trow + row) * Be + j] = hn;
        hnewT[(size_t)j * kRowFloats + row] = hn;
      }
    }
    cluster_arrive();
    OBS_T(3);
    OBS_ACC(1);
    // ---------------------------------------------------------------- P3 posterior hidden
    OBS_T(0);
    prefetch_w(c, a.q1);
    const bool own3 = item_of(c, a.q1, c.tid, &row, &js, &jg);
    float4 pe4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (own3) pe4 = ldrow4(a.PE + (trow + row) * Hi, 4 * jg, Hi, vecH);
    cluster_wait();
    OBS_T(1);
    op_compute<false>(c, a.q1, Segs{hnewT, Be, nullptr});
    OBS_T(2);
    if (own3) {
      const float4 v = red_sum(c, a.q1, js, row);
      const float vv[4] = {v.x + pe4.x, v.y + pe4.y, v.z + pe4.z, v.w + pe4.w};
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int col = 4 * jg + g;
        if (col < Hi) hqT[(size_t)col * kRowFloats + row] = act_fwd(a.act, vv[g]);
      }
    }
    cluster_arrive();
    OBS_T(3);
    OBS_ACC(2);
    // ---------------------------------------------------------------- P4 posterior output + sample
    OBS_T(0);
    prefetch_w(c, a.q2);
    cluster_wait();
    OBS_T(1);
    op_compute<false>(c, a.q2, Segs{hqT, Hi, nullptr});
    OBS_T(2);
    for (int it = c.tid; it < kR * a.q2.WP; it += kThreads) {
      if (item_of(c, a.q2, it, &row, &js, &jg)) {
        const int j = jg;
        const float4 v = red_sum(c, a.q2, js, row);
        const long long o = (trow + row) * S + j;
        const float m = v.x + a.b_q2[j], sd = softplusf_(v.y + a.b_q2[S + j]) + a.min_std;
        a.post_m[o] = m;
        a.post_sd[o] = sd;
        a.post_s[o] = m + sd * a.eps_post[o];
      }
    }
    cluster_arrive();
    OBS_T(3);
    OBS_ACC(3);
  }
  OBS_REPORT("fwd");
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
  int vec_h;        // init_belief / beliefs are 16-byte aligned (float4 loads of their rows)
  int prof;
};
__host__ __device__ inline size_t bwd_scratch_floats(int Be, int Hi, int S) { return (size_t)(Hi + 7 * Be + S) * kRowFloats; }

__global__ void __launch_bounds__(kThreads, 1) observe_bwd_kernel(const __grid_constant__ BwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  Ctx c;
  c.As = smem; c.Ws = smem + kNS * kStageFloats; c.red = smem + 2 * kNS * kStageFloats;
  c.tid = threadIdx.x; c.crank = (int)cluster_ctarank();
  const int Be = a.Be, Bep = a.Bep, Hi = a.Hi, S = a.S, Ad = a.A;
  const long long B = a.B, row0 = (long long)(blockIdx.x / kC) * kR;
  c.nvalid = (int)min((long long)kR, B - row0);
  const int nvalid = c.nvalid;
  float* dhqT = a.scratch + (size_t)(blockIdx.x / kC) * bwd_scratch_floats(Be, Hi, S);
  float* planesT = dhqT + (size_t)Hi * kRowFloats;
  float* dxT = planesT + (size_t)4 * Be * kRowFloats;
  float* czT = dxT + (size_t)Be * kRowFloats;
  float* cbT = czT + (size_t)Be * kRowFloats;
  float* csT = cbT + (size_t)Be * kRowFloats;
  const bool vecH = (Hi & 3) == 0, vecB = (Be & 3) == 0;
  cluster_arrive();
  long long tm[4] = {0, 0, 0, 0}, pw[4] = {0, 0, 0, 0}, po[4] = {0, 0, 0, 0}, pe[4] = {0, 0, 0, 0};
  int row, js, jg;
  for (int i = 0; i < a.L; ++i) {
    const int t = a.L - 1 - i;
    const bool first = i == 0;
    const long long trow = (long long)t * B + row0;
    // ---------------------------------------------------------------- Q1 posterior output backward
    OBS_T(0);
    load_w_small(c, a.b1);
    const bool own1 = item_of(c, a.b1, c.tid, &row, &js, &jg);
    float4 hq4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (own1) hq4 = ldrow4(a.hq + (trow + row) * Hi, 4 * jg, Hi, vecH);
    // in-place operand d_preq^T: everything but the carried state gradient is known before the
    // barrier.  dm = P0 + cs, draw = P1 + cs * Q with P0, P1 parked in the A ring and Q in `red`
    // (both free here; every thread revisits only its own elements)
    // (loads of eight elements are issued together: one loop iteration per element left an L2 round
    // trip per iteration exposed -- 7.7 K cycles for 7.5 elements per thread)
    const int nel = S * kR;
    for (int e0 = c.tid; e0 < nel; e0 += 8 * kThreads) {
      float gs_[8], gm_[8], gd_[8], ep_[8], pq_[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int e = min(e0 + q * kThreads, nel - 1);
        const int j = e >> 6, rl = min(e & 63, nvalid - 1);
        const long long o = (trow + rl) * S + j;
        gs_[q] = a.g_post_s ? a.g_post_s[o] : 0.f;
        gm_[q] = a.g_post_m ? a.g_post_m[o] : 0.f;
        gd_[q] = a.g_post_sd ? a.g_post_sd[o] : 0.f;
        ep_[q] = a.eps_post[o];
        pq_[q] = a.preq[(trow + rl) * 2 * S + S + j];
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int e = e0 + q * kThreads;
        if (e < nel) {
          const int j = e >> 6, r = e & 63;
          const float spg = softplus_gradf_(pq_[q]);
          c.As[j * kLds + r] = gs_[q] + gm_[q];
          c.As[(S + j) * kLds + r] = (gs_[q] * ep_[q] + gd_[q]) * spg;
          c.red[e] = ep_[q] * spg;
        }
      }
    }
    cluster_wait();
    OBS_T(1);
    for (int e0 = c.tid; e0 < nel; e0 += 8 * kThreads) {
      float cs_[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) cs_[q] = first ? 0.f : __ldcg(csT + min(e0 + q * kThreads, nel - 1));
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int e = e0 + q * kThreads;
        if (e < nel) {
          const int j = e >> 6, r = e & 63;
          const float dm = r < nvalid ? c.As[j * kLds + r] + cs_[q] : 0.f;
          const float draw = r < nvalid ? c.As[(S + j) * kLds + r] + cs_[q] * c.red[e] : 0.f;
          c.As[j * kLds + r] = dm;
          c.As[(S + j) * kLds + r] = draw;
          if (c.crank == 0 && r < nvalid) {
            a.tdpreq[(trow + r) * 2 * S + j] = dm;
            a.tdpreq[(trow + r) * 2 * S + S + j] = draw;
          }
        }
      }
    }
    op_compute<true>(c, a.b1, Segs{});
    OBS_T(2);
    if (own1) {
      const float4 v = red_sum(c, a.b1, js, row);
      float4 o;
      o.x = v.x * act_bwd_from_out(a.act, hq4.x); o.y = v.y * act_bwd_from_out(a.act, hq4.y);
      o.z = v.z * act_bwd_from_out(a.act, hq4.z); o.w = v.w * act_bwd_from_out(a.act, hq4.w);
      const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
      for (int g = 0; g < 4; ++g)
        if (4 * jg + g < Hi) dhqT[(size_t)(4 * jg + g) * kRowFloats + row] = ov[g];
      strow4(a.tdhq + (trow + row) * Hi, 4 * jg, Hi, vecH, o);
    }
    cluster_arrive();
    OBS_T(3);
    OBS_ACC(0);
    // ---------------------------------------------------------------- Q2 belief gradient + GRU gates
    OBS_T(0);
    prefetch_w(c, a.b2);
    const bool own2 = item_of(c, a.b2, c.tid, &row, &js, &jg);
    float4 G4, gir, giz, gin, ghr, ghz, ghn4, h4;
    float cbv[4] = {0.f, 0.f, 0.f, 0.f};
    G4 = gir = giz = gin = ghr = ghz = ghn4 = h4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (own2) {
      const int col = 4 * jg;
      const long long o3 = (trow + row) * 3 * Be;
      G4 = ldrow4(a.Gtot + (trow + row) * Be, col, Be, vecB);
      gir = ldrow4(a.gi + o3, col, Be, vecB); giz = ldrow4(a.gi + o3 + Be, col, Be, vecB);
      gin = ldrow4(a.gi + o3 + 2 * Be, col, Be, vecB);
      ghr = ldrow4(a.gh + o3, col, Be, vecB); ghz = ldrow4(a.gh + o3 + Be, col, Be, vecB);
      ghn4 = ldrow4(a.gh + o3 + 2 * Be, col, Be, vecB);
      h4 = t == 0 ? ldrow4(a.init_belief + (row0 + row) * Be, col, Be, vecB && a.vec_h)
                  : ldrow4(a.beliefs + ((long long)(t - 1) * B + row0 + row) * Be, col, Be, vecB && a.vec_h);
      if (!first) {     // carry_b of step t + 1: written in its Q3, two barriers ago
#pragma unroll
        for (int g = 0; g < 4; ++g)
          if (col + g < Be) cbv[g] = __ldcg(cbT + (size_t)(col + g) * kRowFloats + row);
      }
    }
    cluster_wait();
    OBS_T(1);
    op_compute<false>(c, a.b2, Segs{dhqT, Hi, nullptr});
    OBS_T(2);
    if (own2) {
      const float4 v = red_sum(c, a.b2, js, row);
      const float Gq[4] = {v.x, v.y, v.z, v.w}, Gt[4] = {G4.x, G4.y, G4.z, G4.w};
      const float ir[4] = {gir.x, gir.y, gir.z, gir.w}, iz[4] = {giz.x, giz.y, giz.z, giz.w};
      const float in_[4] = {gin.x, gin.y, gin.z, gin.w}, hr[4] = {ghr.x, ghr.y, ghr.z, ghr.w};
      const float hz[4] = {ghz.x, ghz.y, ghz.z, ghz.w}, hn[4] = {ghn4.x, ghn4.y, ghn4.z, ghn4.w};
      const float hp[4] = {h4.x, h4.y, h4.z, h4.w};
      float dr[4], dz_[4], dn_[4], dnr[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int col = 4 * jg + g;
        dr[g] = dz_[g] = dn_[g] = dnr[g] = 0.f;
        if (col < Be) {
          const float G = Gt[g] + cbv[g] + Gq[g];
          const float rr = sigmoidf_(ir[g] + hr[g]);
          const float z = sigmoidf_(iz[g] + hz[g]);
          const float n = tanhf(in_[g] + rr * hn[g]);
          const float dn = G * (1.f - z), dzz = G * (hp[g] - n);
          const float dpn = dn * (1.f - n * n);
          dr[g] = dpn * hn[g] * rr * (1.f - rr);
          dz_[g] = dzz * z * (1.f - z);
          dn_[g] = dpn;
          dnr[g] = dpn * rr;
          planesT[(size_t)col * kRowFloats + row] = dr[g];
          planesT[(size_t)(Be + col) * kRowFloats + row] = dz_[g];
          planesT[(size_t)(2 * Be + col) * kRowFloats + row] = dn_[g];
          planesT[(size_t)(3 * Be + col) * kRowFloats + row] = dnr[g];
          czT[(size_t)col * kRowFloats + row] = G * z;
        }
      }
      const long long o3 = (trow + row) * 3 * Be;
      const float4 r4 = make_float4(dr[0], dr[1], dr[2], dr[3]), z4 = make_float4(dz_[0], dz_[1], dz_[2], dz_[3]);
      strow4(a.tdgi + o3, 4 * jg, Be, vecB, r4);
      strow4(a.tdgh + o3, 4 * jg, Be, vecB, r4);
      strow4(a.tdgi + o3 + Be, 4 * jg, Be, vecB, z4);
      strow4(a.tdgh + o3 + Be, 4 * jg, Be, vecB, z4);
      strow4(a.tdgi + o3 + 2 * Be, 4 * jg, Be, vecB, make_float4(dn_[0], dn_[1], dn_[2], dn_[3]));
      strow4(a.tdgh + o3 + 2 * Be, 4 * jg, Be, vecB, make_float4(dnr[0], dnr[1], dnr[2], dnr[3]));
    }
    cluster_arrive();
    OBS_T(3);
    OBS_ACC(1);
    // ---------------------------------------------------------------- Q3 GRU input / hidden dgrad
    OBS_T(0);
    prefetch_w(c, a.b3);
    cluster_wait();
    OBS_T(1);
    if (a.b3.tc) {
      if (a.b3.WP == 16) op_compute_tc<4>(c, a.b3, Segs{planesT, 4 * Be, nullptr});
      else op_compute_tc<2>(c, a.b3, Segs{planesT, 4 * Be, nullptr});
    } else {
      op_compute<false>(c, a.b3, Segs{planesT, 4 * Be, nullptr});
    }
    OBS_T(2);
    for (int it = c.tid; it < kR * a.b3.WP; it += kThreads) {
      if (item_of(c, a.b3, it, &row, &js, &jg)) {
        const float4 v = red_sum(c, a.b3, js, row);
        const float vv[4] = {v.x, v.y, v.z, v.w};
        const int col0 = 4 * jg;
        if (col0 < Bep) {
          const float4 x4 = ldrow4(a.x + (trow + row) * Be, col0, Be, vecB);
          const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
          float ov[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            ov[g] = vv[g] * act_bwd_from_out(a.act, xv[g]);
            if (col0 + g < Be) dxT[(size_t)(col0 + g) * kRowFloats + row] = ov[g];
          }
          strow4(a.tdx + (trow + row) * Be, col0, Be, vecB, make_float4(ov[0], ov[1], ov[2], ov[3]));
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int cc = col0 - Bep + g;
            if (cc < Be) {
              const float o = vv[g] + __ldcg(czT + (size_t)cc * kRowFloats + row);
              cbT[(size_t)cc * kRowFloats + row] = o;
              if (t == 0 && a.d_init_belief) a.d_init_belief[(row0 + row) * Be + cc] = o;
            }
          }
        }
      }
    }
    cluster_arrive();
    OBS_T(3);
    OBS_ACC(2);
    // ---------------------------------------------------------------- Q4 embed dgrad
    OBS_T(0);
    prefetch_w(c, a.b4);
    cluster_wait();
    OBS_T(1);
    op_compute<false>(c, a.b4, Segs{dxT, Be, nullptr});
    OBS_T(2);
    if (item_of(c, a.b4, c.tid, &row, &js, &jg)) {
      const float4 v = red_sum(c, a.b4, js, row);
      const float vv[4] = {v.x, v.y, v.z, v.w};
      const float nt = a.nonterm ? a.nonterm[trow + row] : 1.f;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int col = 4 * jg + g;
        if (col < S) {
          const float o = vv[g] * nt;
          csT[(size_t)col * kRowFloats + row] = o;
          if (t == 0 && a.d_init_state) a.d_init_state[(row0 + row) * S + col] = o;
        } else if (col < S + Ad && a.d_actions) {
          a.d_actions[(trow + row) * Ad + (col - S)] = vv[g];
        }
      }
    }
    cluster_arrive();
    OBS_T(3);
    OBS_ACC(3);
  }
  OBS_REPORT("bwd");
  cluster_wait();
}

}  // namespace obs
}  // namespace f32
}  // namespace bd

// fp32 "check mode" kernels: plain FFMA contractions and the fused elementwise
// stages of the RSSM step.  Every kernel is hand-written for sm_100a; no library
// GEMM is called.  Arithmetic follows SURVEY.md Appendix A (reference lines cited
// at each kernel).
#pragma once
#include "common.cuh"

namespace bd {
namespace f32 {

// =====================================================================================
// Tiled SGEMM   C[m,n] = epi( sum_k A(m,k) * B(k,n) )
//   A_TRANS=false: A(m,k) = A[m*lda + k]   (optionally two K segments: [A1 | A2] = torch.cat)
//   A_TRANS=true : A(m,k) = A[k*lda + m]
//   B_TRANS=true : B(k,n) = B[n*ldb + k]   (nn.Linear weight (out,in): forward)
//   B_TRANS=false: B(k,n) = B[k*ldb + n]   (dgrad: dX = dY * W;  wgrad: X)
// =====================================================================================
enum Epi {
  EPI_BIAS_ACT = 0,  // C = act(acc + bias[n])            (+ beta*C)
  EPI_MUL_DACT = 1,  // C = acc * act'(aux[m,n])          (+ beta*C)   backward through an activation
  EPI_ATOMIC = 2     // atomicAdd(C, acc)                 split-K wgrad
};

struct GemmArgs {
  const float* A1 = nullptr; long long lda1 = 0; int K1 = 0;
  const float* A2 = nullptr; long long lda2 = 0; int K2 = 0;
  const float* rowscale1 = nullptr;  // optional per-row multiplier applied to segment 1
  const float* B = nullptr; long long ldb = 0;
  float* C = nullptr; long long ldc = 0;
  const float* bias = nullptr;
  const float* aux = nullptr; long long ldaux = 0;
  int M = 0, N = 0;
  int act = BD_ACT_IDENTITY;
  int beta = 0;     // 1: accumulate into C
  int ksplit = 1;   // EPI_ATOMIC: K range split over gridDim.z
};

// 16-bit precision modes: the time-batched GEMMs of the observe pass (all L*B rows at once) run the
// same tiling on TF32 mma.sync (set around those calls by Tf32Scope, api_fp32.cu); check mode never sets it.
inline bool& gemm_tf32_flag() { static thread_local bool f = false; return f; }
struct Tf32Scope {
  bool prev;
  explicit Tf32Scope(bool on) : prev(gemm_tf32_flag()) { gemm_tf32_flag() = on; }
  ~Tf32Scope() { gemm_tf32_flag() = prev; }
};

constexpr int BM = 64, BN = 64, BK = 16, GEMM_THREADS = 256;

template <bool A_TRANS, bool B_TRANS, int EPI, bool TF32 = false>
__global__ void __launch_bounds__(GEMM_THREADS) sgemm_kernel(GemmArgs g) {
  // row stride 72: float4 rows stay 16-byte aligned and the mma.sync fragment loads of the TF32 variant
  // (4 k rows x 8 columns per instruction) fall on 32 different banks
  __shared__ __align__(16) float As[2][BK][BM + 8];
  __shared__ __align__(16) float Bs[2][BK][BN + 8];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int K = g.K1 + g.K2;
  int kbeg = 0, kend = K;
  if (EPI == EPI_ATOMIC) {
    int per = (K + g.ksplit - 1) / g.ksplit;
    per = (per + BK - 1) / BK * BK;
    kbeg = blockIdx.z * per;
    kend = min(K, kbeg + per);
    if (kbeg >= kend) return;
  }
  float ra[4], rb[4];
  auto load_tile = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int m, k;
      if (A_TRANS) { m = tid % BM; k = tid / BM + i * (GEMM_THREADS / BM); }
      else         { k = tid % BK; m = tid / BK + i * (GEMM_THREADS / BK); }
      int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < g.M && gk < kend) {
        if (A_TRANS) v = g.A1[(long long)gk * g.lda1 + gm];
        else if (gk < g.K1) {
          v = g.A1[(long long)gm * g.lda1 + gk];
          if (g.rowscale1) v *= g.rowscale1[gm];
        } else v = g.A2[(long long)gm * g.lda2 + (gk - g.K1)];
      }
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int n, k;
      if (B_TRANS) { k = tid % BK; n = tid / BK + i * (GEMM_THREADS / BK); }
      else         { n = tid % BN; k = tid / BN + i * (GEMM_THREADS / BN); }
      int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < g.N && gk < kend)
        v = B_TRANS ? g.B[(long long)gn * g.ldb + gk] : g.B[(long long)gk * g.ldb + gn];
      rb[i] = v;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int m, k;
      if (A_TRANS) { m = tid % BM; k = tid / BM + i * (GEMM_THREADS / BM); }
      else         { k = tid % BK; m = tid / BK + i * (GEMM_THREADS / BK); }
      As[buf][k][m] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int n, k;
      if (B_TRANS) { k = tid % BK; n = tid / BK + i * (GEMM_THREADS / BK); }
      else         { n = tid % BN; k = tid / BN + i * (GEMM_THREADS / BN); }
      Bs[buf][k][n] = rb[i];
    }
  };

  const int tx = tid % 16, ty = tid / 16;
  // TF32 variant: warp w owns m tile (w >> 1) and n tiles 4 (w & 1) .. + 3 of the 64 x 64 block
  const int lane = tid & 31, wrp = tid >> 5, fg = lane >> 2, ft = lane & 3;
  const int mt0 = (wrp >> 1) * 16, nt0 = (wrp & 1) * 32;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  load_tile(kbeg);
  store_tile(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    const bool more = k0 + BK < kend;
    if (more) load_tile(k0 + BK);
    if (TF32) {
#pragma unroll
      for (int ks = 0; ks < BK; ks += 8) {
        uint32_t af[4], bf[4][2];
        af[0] = f32_to_tf32(As[buf][ks + ft][mt0 + fg]);
        af[1] = f32_to_tf32(As[buf][ks + ft][mt0 + fg + 8]);
        af[2] = f32_to_tf32(As[buf][ks + ft + 4][mt0 + fg]);
        af[3] = f32_to_tf32(As[buf][ks + ft + 4][mt0 + fg + 8]);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          bf[nt][0] = f32_to_tf32(Bs[buf][ks + ft][nt0 + 8 * nt + fg]);
          bf[nt][1] = f32_to_tf32(Bs[buf][ks + ft + 4][nt0 + 8 * nt + fg]);
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
          asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                       : "+f"(acc[nt][0]), "+f"(acc[nt][1]), "+f"(acc[nt][2]), "+f"(acc[nt][3])
                       : "r"(af[0]), "r"(af[1]), "r"(af[2]), "r"(af[3]), "r"(bf[nt][0]), "r"(bf[nt][1]));
      }
    } else {
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float4 a4 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
        float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
        float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {     // packed FFMA2: twice the issue rate of scalar FFMA on sm_100
          ffma2(acc[i][0], acc[i][1], a[i], b[0], b[1]);
          ffma2(acc[i][2], acc[i][3], a[i], b[2], b[3]);
        }
      }
    }
    if (more) {
      store_tile(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // FFMA: thread (ty, tx) owns rows 4 ty + i, columns 4 tx + j; TF32: acc[i][j] is element j of the
      // C fragment of n tile i (rows fg / fg + 8, columns 2 ft / 2 ft + 1)
      const int gm = TF32 ? m0 + mt0 + fg + (j >= 2 ? 8 : 0) : m0 + ty * 4 + i;
      const int gn = TF32 ? n0 + nt0 + 8 * i + 2 * ft + (j & 1) : n0 + tx * 4 + j;
      if (gm >= g.M || gn >= g.N) continue;
      float v = acc[i][j];
      float* c = g.C + (long long)gm * g.ldc + gn;
      if (EPI == EPI_BIAS_ACT) {
        if (g.bias) v += g.bias[gn];
        v = act_fwd(g.act, v);
        if (g.beta) v += *c;
        *c = v;
      } else if (EPI == EPI_MUL_DACT) {
        if (g.aux) v *= act_bwd_from_out(g.act, g.aux[(long long)gm * g.ldaux + gn]);
        if (g.beta) v += *c;
        *c = v;
      } else {
        atomicAdd(c, v);
      }
    }
  }
}

// Skinny variant for M <= 64 rows (one 50-row batch of the observe pass, src/models.py:239-271):
// the 64x64 tiling above would leave 4-10 CTAs walking K in 16-column steps, one L2 round trip
// each (25-32 us per GEMM).  Here a CTA owns all rows x 16 output columns (N/16 CTAs) and moves
// K in 128-column chunks with the next chunk's loads in flight during the FMAs.
constexpr int SK_BN = 16, SK_KC = 128;
template <bool B_TRANS, int EPI>
__global__ void __launch_bounds__(256) sgemm_skinny_kernel(GemmArgs g) {
  __shared__ float As[SK_KC][65];
  __shared__ float Bs[SK_KC][SK_BN];
  const int tid = threadIdx.x, n0 = blockIdx.x * SK_BN;
  const int K = g.K1 + g.K2;
  const int m = tid & 63, ng = tid >> 6;      // this thread's outputs: row m, columns n0 + 4 ng .. +3
  const int lk = tid & 127, lh = tid >> 7;    // loader mapping: K offset inside the chunk, row parity
  float ra[32], rb[8];
  // Branch-free loads (clamped addresses, value selected afterwards) so that all 40 loads of a
  // chunk are in flight together; nested `if`s around each load serialise them (22 us per GEMM).
  auto load_chunk = [&](int k0) {
    const int gk = k0 + lk;
    const bool kok = gk < K;
    const bool seg1 = gk < g.K1;
    const int gkc = kok ? gk : K - 1;
    const float* base = (seg1 || g.K2 == 0) ? g.A1 + min(gkc, g.K1 - 1) : g.A2 + (gkc - g.K1);
    const long long ld = (seg1 || g.K2 == 0) ? g.lda1 : g.lda2;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int gm = lh + 2 * i;
      const float v = base[(long long)min(gm, g.M - 1) * ld];
      ra[i] = (kok && gm < g.M) ? v : 0.f;
    }
    if (g.rowscale1 && seg1) {
#pragma unroll
      for (int i = 0; i < 32; ++i) ra[i] *= g.rowscale1[min(lh + 2 * i, g.M - 1)];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int n, k;
      if (B_TRANS) { k = lk; n = lh + 2 * i; }
      else         { n = tid & 15; k = (tid >> 4) + 16 * i; }
      const int gn = n0 + n, gkb = k0 + k;
      const int gnc = min(gn, g.N - 1), gkc2 = min(gkb, K - 1);
      const float v = B_TRANS ? g.B[(long long)gnc * g.ldb + gkc2] : g.B[(long long)gkc2 * g.ldb + gnc];
      rb[i] = (gn < g.N && gkb < K) ? v : 0.f;
    }
  };
  auto store_chunk = [&]() {
#pragma unroll
    for (int i = 0; i < 32; ++i) As[lk][lh + 2 * i] = ra[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (B_TRANS) Bs[lk][lh + 2 * i] = rb[i];
      else Bs[(tid >> 4) + 16 * i][tid & 15] = rb[i];
    }
  };
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  load_chunk(0);
  for (int k0 = 0; k0 < K; k0 += SK_KC) {
    store_chunk();
    __syncthreads();
    if (k0 + SK_KC < K) load_chunk(k0 + SK_KC);
    const int kc = min(SK_KC, K - k0);
#pragma unroll 8
    for (int k = 0; k < kc; ++k) {
      const float a = As[k][m];
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][ng * 4]);
      ffma2(acc[0], acc[1], a, b4.x, b4.y);
      ffma2(acc[2], acc[3], a, b4.z, b4.w);
    }
    __syncthreads();
  }
  if (m >= g.M) return;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int gn = n0 + ng * 4 + j;
    if (gn >= g.N) continue;
    float v = acc[j];
    float* c = g.C + (long long)m * g.ldc + gn;
    if (EPI == EPI_BIAS_ACT) {
      if (g.bias) v += g.bias[gn];
      v = act_fwd(g.act, v);
      if (g.beta) v += *c;
      *c = v;
    } else {
      if (g.aux) v *= act_bwd_from_out(g.act, g.aux[(long long)m * g.ldaux + gn]);
      if (g.beta) v += *c;
      *c = v;
    }
  }
}

// Row-vector variant for M <= 4 rows and a (N, K) row-major weight (the acting path: one environment step is ONE
// row through every layer, src/planet.py:370-403).  One warp per output column: the lanes read the weight row with
// all their loads in flight (one L2 round trip for the whole K) and meet by shuffle; the skinny kernel above walks K
// in 128-column chunks, one round trip each (19 us per layer at K = 1024 against ~3 us here).
constexpr int GV_MAXM = 4;
template <int EPI>
__global__ void __launch_bounds__(256) sgemm_gemv_kernel(GemmArgs g) {
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= g.N) return;
  const int K = g.K1 + g.K2;
  const float* __restrict__ w = g.B + (long long)n * g.ldb;
  float acc[GV_MAXM];
#pragma unroll
  for (int m = 0; m < GV_MAXM; ++m) acc[m] = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float wv = w[k];
    const bool seg1 = k < g.K1;
#pragma unroll
    for (int m = 0; m < GV_MAXM; ++m) {
      if (m < g.M) {
        float a = seg1 ? g.A1[(long long)m * g.lda1 + k] : g.A2[(long long)m * g.lda2 + (k - g.K1)];
        if (seg1 && g.rowscale1) a *= g.rowscale1[m];
        acc[m] = fmaf(a, wv, acc[m]);
      }
    }
  }
#pragma unroll
  for (int m = 0; m < GV_MAXM; ++m)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[m] += __shfl_xor_sync(0xffffffffu, acc[m], o);
  if (lane < g.M) {
    float v = 0.f;
#pragma unroll
    for (int m = 0; m < GV_MAXM; ++m)
      if (lane == m) v = acc[m];
    float* c = g.C + (long long)lane * g.ldc + n;
    if (EPI == EPI_BIAS_ACT) {
      if (g.bias) v += g.bias[n];
      v = act_fwd(g.act, v);
    } else {
      if (g.aux) v *= act_bwd_from_out(g.act, g.aux[(long long)lane * g.ldaux + n]);
    }
    if (g.beta) v += *c;
    *c = v;
  }
}

template <bool A_TRANS, bool B_TRANS, int EPI>
inline int launch_gemm(const GemmArgs& g, cudaStream_t s) {
  if (g.M <= 0 || g.N <= 0) return BD_OK;
  if constexpr (!A_TRANS && B_TRANS && EPI != EPI_ATOMIC) {
    if (g.M <= GV_MAXM) {
      sgemm_gemv_kernel<EPI><<<(g.N + 7) / 8, 256, 0, s>>>(g);
      BD_CUDA_LAUNCH_CHECK();
      return BD_OK;
    }
  }
  if constexpr (!A_TRANS && EPI != EPI_ATOMIC) {
    if (g.M <= 64) {
      sgemm_skinny_kernel<B_TRANS, EPI><<<(g.N + SK_BN - 1) / SK_BN, 256, 0, s>>>(g);
      BD_CUDA_LAUNCH_CHECK();
      return BD_OK;
    }
  }
  dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, EPI == EPI_ATOMIC ? g.ksplit : 1);
  if (gemm_tf32_flag()) sgemm_kernel<A_TRANS, B_TRANS, EPI, true><<<grid, GEMM_THREADS, 0, s>>>(g);
  else sgemm_kernel<A_TRANS, B_TRANS, EPI, false><<<grid, GEMM_THREADS, 0, s>>>(g);
  BD_CUDA_LAUNCH_CHECK();
  return BD_OK;
}

// Y(rows,out) = act([X1|X2] W^T + b)   -- nn.Linear + activation (src/utils.py:399-403)
inline int linear_fwd(const bd_linear& L, int act, const float* x1, int k1, long long ld1,
                      const float* x2, int k2, long long ld2, const float* rowscale1, int rows,
                      float* y, long long ldy, cudaStream_t s) {
  GemmArgs g;
  g.A1 = x1; g.lda1 = ld1; g.K1 = k1; g.A2 = x2; g.lda2 = ld2; g.K2 = k2; g.rowscale1 = rowscale1;
  g.B = L.w; g.ldb = L.in_features; g.C = y; g.ldc = ldy; g.bias = L.b;
  g.M = rows; g.N = L.out_features; g.act = act;
  return launch_gemm<false, true, EPI_BIAS_ACT>(g, s);
}
// plain Y = X W^T + b with explicit weight pointer (GRU input/hidden projections)
inline int matmul_nt_bias(const float* x, int k, long long ldx, const float* w, const float* b,
                          int n_out, int rows, float* y, long long ldy, cudaStream_t s) {
  GemmArgs g;
  g.A1 = x; g.lda1 = ldx; g.K1 = k; g.B = w; g.ldb = k; g.C = y; g.ldc = ldy; g.bias = b;
  g.M = rows; g.N = n_out;
  return launch_gemm<false, true, EPI_BIAS_ACT>(g, s);
}
// dX(rows, ncols) (+)= (dY(rows,out) * W[:, col0:col0+ncols]) (.) act'(aux)
inline int linear_dgrad(const float* dy, int out, long long lddy, const float* w, int in_total,
                        int col0, int ncols, int rows, float* dx, long long lddx, int act,
                        const float* aux, long long ldaux, int beta, cudaStream_t s) {
  GemmArgs g;
  g.A1 = dy; g.lda1 = lddy; g.K1 = out; g.B = w + col0; g.ldb = in_total; g.C = dx; g.ldc = lddx;
  g.M = rows; g.N = ncols; g.act = act; g.aux = aux; g.ldaux = ldaux; g.beta = beta;
  return launch_gemm<false, false, EPI_MUL_DACT>(g, s);
}
// dW[:, col0:col0+ncols] += dY^T X     (rows reduced with split-K atomics)
inline int linear_wgrad(const float* dy, int out, long long lddy, const float* x, int ncols,
                        long long ldx, int rows, float* dw, int in_total, int col0,
                        cudaStream_t s) {
  GemmArgs g;
  g.A1 = dy; g.lda1 = lddy; g.K1 = rows; g.B = x; g.ldb = ldx; g.C = dw + col0; g.ldc = in_total;
  g.M = out; g.N = ncols;
  int tiles = ((out + BM - 1) / BM) * ((ncols + BN - 1) / BN);
  int want = (4 * 148 + tiles - 1) / tiles;
  int maxsplit = (rows + 4 * BK - 1) / (4 * BK);
  g.ksplit = max(1, min(want, maxsplit));
  return launch_gemm<true, false, EPI_ATOMIC>(g, s);
}

// db[n] += sum_rows dY[row, n]
__global__ void colsum_atomic_kernel(const float* __restrict__ dy, long long ld, int rows, int n,
                                     float* __restrict__ db, int rows_per_block) {
  int col = blockIdx.x * blockDim.x + threadIdx.x;
  int r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  if (col >= n) return;
  float s = 0.f;
  for (int r = r0; r < r1; ++r) s += dy[(long long)r * ld + col];
  atomicAdd(db + col, s);
}
inline int bias_grad(const float* dy, int n, long long ld, int rows, float* db, cudaStream_t s) {
  if (rows <= 0) return BD_OK;
  // few rows (time-batched observe pass: 2 450): 32-row slabs keep every SM busy instead of
  // 10 blocks walking 256 rows each (27 us -> a few us per column sum)
  int rpb = rows <= 16384 ? 32 : 256;
  dim3 grid((n + 63) / 64, (rows + rpb - 1) / rpb);
  colsum_atomic_kernel<<<grid, 64, 0, s>>>(dy, ld, rows, n, db, rpb);
  BD_CUDA_LAUNCH_CHECK();
  return BD_OK;
}

// =====================================================================================
// GRUCell gates (nn.GRUCell; src/models.py:252, src/dreamer.py:220).  gi, gh already
// hold W x + b.  Gate order r,z,n;  n = tanh(gi_n + r * gh_n).
// =====================================================================================
__global__ void gru_gate_fwd_kernel(const float* __restrict__ gi, const float* __restrict__ gh,
                                    const float* __restrict__ h_prev, float* __restrict__ h_new,
                                    long long total, int Be) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  long long row = i / Be;
  int j = (int)(i - row * Be);
  const float* gir = gi + row * 3 * Be;
  const float* ghr = gh + row * 3 * Be;
  float r = sigmoidf_(gir[j] + ghr[j]);
  float z = sigmoidf_(gir[Be + j] + ghr[Be + j]);
  float n = tanhf(gir[2 * Be + j] + r * ghr[2 * Be + j]);
  float h = h_prev[i];
  h_new[i] = (1.f - z) * n + z * h;
}
// backward: G = dL/dh_new.  Writes d_gi, d_gh (rows,3Be) and carry = G*z.
__global__ void gru_gate_bwd_kernel(const float* __restrict__ gi, const float* __restrict__ gh,
                                    const float* __restrict__ h_prev, const float* __restrict__ G,
                                    float* __restrict__ d_gi, float* __restrict__ d_gh,
                                    float* __restrict__ carry, long long total, int Be) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  long long row = i / Be;
  int j = (int)(i - row * Be);
  long long o = row * 3 * Be;
  float ghn = gh[o + 2 * Be + j];
  float r = sigmoidf_(gi[o + j] + gh[o + j]);
  float z = sigmoidf_(gi[o + Be + j] + gh[o + Be + j]);
  float n = tanhf(gi[o + 2 * Be + j] + r * ghn);
  float g = G[i], h = h_prev[i];
  float dn = g * (1.f - z);
  float dz = g * (h - n);
  float dpn = dn * (1.f - n * n);
  float dpr = dpn * ghn * r * (1.f - r);
  float dpz = dz * z * (1.f - z);
  d_gi[o + j] = dpr;          d_gh[o + j] = dpr;
  d_gi[o + Be + j] = dpz;     d_gh[o + Be + j] = dpz;
  d_gi[o + 2 * Be + j] = dpn; d_gh[o + 2 * Be + j] = dpn * r;
  carry[i] = g * z;
}

// =====================================================================================
// GaussianBeliefModel tail (src/models.py:70-73): chunk, softplus + min_std, sample.
// =====================================================================================
__global__ void belief_sample_fwd_kernel(const float* __restrict__ pre, const float* __restrict__ eps,
                                         float min_std, float* __restrict__ state,
                                         float* __restrict__ mean, float* __restrict__ stdv,
                                         long long total, int S) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  long long row = i / S;
  int j = (int)(i - row * S);
  float m = pre[row * 2 * S + j];
  float sd = softplusf_(pre[row * 2 * S + S + j]) + min_std;
  mean[i] = m;
  stdv[i] = sd;
  state[i] = m + sd * eps[i];
}
// d_pre (rows,2S) from upstream grads of state / mean / std (each optional) (+ carry on state)
__global__ void belief_sample_bwd_kernel(const float* __restrict__ pre, const float* __restrict__ eps,
                                         const float* __restrict__ g_state,
                                         const float* __restrict__ g_state2,
                                         const float* __restrict__ g_mean,
                                         const float* __restrict__ g_std, float* __restrict__ d_pre,
                                         long long total, int S) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  long long row = i / S;
  int j = (int)(i - row * S);
  float gs = (g_state ? g_state[i] : 0.f) + (g_state2 ? g_state2[i] : 0.f);
  float dm = gs + (g_mean ? g_mean[i] : 0.f);
  float dsd = gs * eps[i] + (g_std ? g_std[i] : 0.f);
  d_pre[row * 2 * S + j] = dm;
  d_pre[row * 2 * S + S + j] = dsd * softplus_gradf_(pre[row * 2 * S + S + j]);
}

// =====================================================================================
// KL(posterior || prior) loss of the dynamics update:
// Planet._kl_loss src/planet.py:288-308 and Dreamer._kl_loss src/dreamer.py:111-146
// (torch.distributions.kl._kl_normal_normal, free nats, optional KL balancing).
//   kl_rows_kernel    one warp per (t, b) row: div[row] = sum_S KL
//   kl_finish_kernel  one block, deterministic tree: loss = mean_rows max(div, free_nats)   (balance < 0)
//                     or max(mean_elements KL, free_nats) (balanced: both detached variants
//                     have the same VALUE; they differ in where the gradient goes)
//   kl_bwd_kernel     one thread per element -> up to four gradients
// torch.max(a, b) passes the gradient to the larger input and half of it on a tie.
// =====================================================================================
__device__ __forceinline__ float kl_normal(float mq, float sq, float mp, float sp) {
  const float vr = (sq / sp) * (sq / sp), t1 = ((mq - mp) / sp) * ((mq - mp) / sp);
  return 0.5f * (vr + t1 - 1.f - logf(vr));
}
__global__ void kl_rows_kernel(const float* __restrict__ mq, const float* __restrict__ sq,
                               const float* __restrict__ mp, const float* __restrict__ sp,
                               long long rows, int S, float* __restrict__ div) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float acc = 0.f;
  for (int j = lane; j < S; j += 32) {
    const long long o = row * S + j;
    acc += kl_normal(mq[o], sq[o], mp[o], sp[o]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) div[row] = acc;
}
__device__ __forceinline__ float max_grad_factor(float a, float b) { return a > b ? 1.f : (a == b ? 0.5f : 0.f); }
__global__ void __launch_bounds__(1024) kl_finish_kernel(const float* __restrict__ div, long long rows, int S,
                                                         const float* __restrict__ free_nats, int balanced,
                                                         float* __restrict__ loss) {
  __shared__ float red[1024];
  const float fn = free_nats[0];
  float acc = 0.f;
  for (long long r = threadIdx.x; r < rows; r += 1024) acc += balanced ? div[r] : fmaxf(div[r], fn);
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    // loss[1] keeps the mean element KL for the backward of the balanced form
    if (balanced) { const float m = red[0] / ((float)rows * (float)S); loss[0] = fmaxf(m, fn); loss[1] = m; }
    else { loss[0] = red[0] / (float)rows; loss[1] = 0.f; }
  }
}
__global__ void kl_bwd_kernel(const float* __restrict__ mq, const float* __restrict__ sq,
                              const float* __restrict__ mp, const float* __restrict__ sp,
                              const float* __restrict__ div, const float* __restrict__ loss,
                              const float* __restrict__ free_nats, const float* __restrict__ g_loss,
                              long long rows, int S, float balance, float* __restrict__ d_mq,
                              float* __restrict__ d_sq, float* __restrict__ d_mp, float* __restrict__ d_sp) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * S) return;
  const float fn = free_nats[0], g = g_loss[0];
  float wq, wp;     // weights of the posterior-side / prior-side gradients
  if (balance < 0.f) {
    wq = wp = g * max_grad_factor(div[i / S], fn) / (float)rows;
  } else {
    const float w = g * max_grad_factor(loss[1], fn) / ((float)rows * (float)S);
    wp = balance * w;            // kl_lhs: posterior detached
    wq = (1.f - balance) * w;    // kl_rhs: prior detached
  }
  const float a = mq[i], b = sq[i], c = mp[i], d = sp[i];
  const float diff = a - c, inv_var = 1.f / (d * d);
  if (d_mq) d_mq[i] = wq * diff * inv_var;
  if (d_sq) d_sq[i] = wq * (b * inv_var - 1.f / b);
  if (d_mp) d_mp[i] = -wp * diff * inv_var;
  if (d_sp) d_sp[i] = wp * (1.f / d - (b * b + diff * diff) * inv_var / d);
}

// =====================================================================================
// Critic regression loss of Dreamer's value update (src/dreamer.py:380-385):
//   loss = -mean( w * Normal(v, 1).log_prob(target) ) = mean( w * (0.5 (v - target)^2 + 0.5 ln 2 pi) )
// (w = the cumulated discount when use_discount, else 1) and its gradient d loss / d v, fused.
//   value_loss_kernel         grid-stride: dv[i] = w (v - t) / n ; per-block partial sums
//   value_loss_finish_kernel  one block, fixed-order tree over the partials (deterministic)
// =====================================================================================
constexpr int kValueLossBlocks = 1024;
__global__ void __launch_bounds__(256) value_loss_kernel(const float* __restrict__ v, const float* __restrict__ t,
                                                         const float* __restrict__ w, long long n,
                                                         float* __restrict__ dv, float* __restrict__ partial) {
  __shared__ float red[256];
  const float inv_n = 1.f / (float)n;
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float d = v[i] - t[i], wi = w ? w[i] : 1.f;
    acc += wi * (0.5f * d * d + 0.9189385332046727f);
    if (dv) dv[i] = wi * d * inv_n;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}
__global__ void __launch_bounds__(1024) value_loss_finish_kernel(const float* __restrict__ partial, int nblocks,
                                                                 long long n, float* __restrict__ loss) {
  __shared__ float red[1024];
  red[threadIdx.x] = (int)threadIdx.x < nblocks ? partial[threadIdx.x] : 0.f;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = red[0] / (float)n;
}

// =====================================================================================
// Actor head: ActorModel squash (src/models.py:513-516), tanh-Normal rsample
// (src/dreamer.py:435-443) and the J-sample Monte-Carlo entropy (src/models.py:725-733,
// 656-673).  One thread per row.  Also emits d entropy / d(mean,std) for backward.
// =====================================================================================
__device__ __forceinline__ void actor_squash(float m_raw, float s_raw, const bd_actor_cfg& c,
                                             float& mean, float& sd) {
  mean = c.mean_scale * tanhf(m_raw / c.mean_scale);
  sd = softplusf_(s_raw + c.raw_init_std) + c.min_std;
}

__global__ void actor_head_fwd_kernel(const float* __restrict__ raw, const float* __restrict__ eps_a,
                                      const float* __restrict__ eps_e, bd_actor_cfg cfg,
                                      long long row0, long long n_total, int rows, int A,
                                      float* __restrict__ action, float* __restrict__ entropy,
                                      float* __restrict__ dent) {
  // raw/eps_a/action/entropy/dent are already offset to this chunk; eps_e is the full
  // (J, n_total, A) slab of this time step and is indexed with row0 + r.
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float kClamp = 0.99999997f;  // rounds to 0.99999994f (src/models.py:663)
  const float kLogSqrt2Pi = 0.9189385332046727f;
  const float kLog2 = 0.6931471805599453f;
  float ent_acc = 0.f;
  const int J = cfg.entropy_samples;
  for (int a = 0; a < A; ++a) {
    float mean, sd;
    actor_squash(raw[(long long)r * 2 * A + a], raw[(long long)r * 2 * A + A + a], cfg, mean, sd);
    action[(long long)r * A + a] = tanhf(mean + eps_a[(long long)r * A + a] * sd);
    float var2 = 2.f * sd * sd, log_sd = logf(sd), inv_var = 1.f / (sd * sd);
    float lp_sum = 0.f, dm_sum = 0.f, ds_sum = 0.f;
    for (int j = 0; j < J; ++j) {
      float e = eps_e[((long long)j * n_total + row0 + r) * A + a];
      float y = tanhf(mean + e * sd);
      float yc = fminf(fmaxf(y, -kClamp), kClamp);
      float gate = (yc == y) ? 1.f : 0.f;                 // clamp passes no gradient when active
      float xh = 0.5f * logf((1.f + yc) / (1.f - yc));    // atanh (src/models.py:622-627)
      float d = xh - mean;
      float lp = -(d * d) / var2 - log_sd - kLogSqrt2Pi
                 - 2.f * (kLog2 - xh - softplusf_(-2.f * xh));
      lp_sum += lp;
      float dlp_dx = -d * inv_var + 2.f * tanhf(xh);      // d lp / d xhat
      dm_sum += d * inv_var + gate * dlp_dx;
      ds_sum += d * d * inv_var / sd - 1.f / sd + gate * e * dlp_dx;
    }
    ent_acc += lp_sum;
    dent[(long long)r * 2 * A + a] = -dm_sum / (float)J;
    dent[(long long)r * 2 * A + A + a] = -ds_sum / (float)J;
  }
  entropy[r] = -ent_acc / (float)J;
}

// Acting (src/dreamer.py:429-444 called from Planet.update_belief_and_act, src/planet.py:370-403): the action
// alone, no entropy.  One warp per row.
//   deterministic == 0: action = tanh(mean + eps sd), eps (rows, A)            (dist.rsample, :443)
//   deterministic == 1: SampleDist.mode (src/models.py:707-723): J samples y_j = tanh(mean + eps_j sd), eps
//     (J, rows, A); the one with the largest log-probability (summed over A; first maximum) is the action.
//     The log-probability is that of actor_head_fwd_kernel (tanh-Normal with the clamped inverse).
__global__ void actor_act_kernel(const float* __restrict__ raw, const float* __restrict__ eps, bd_actor_cfg cfg,
                                 long long rows, int A, int deterministic, float* __restrict__ action) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  if (!deterministic) {
    for (int a = lane; a < A; a += 32) {
      float mean, sd;
      actor_squash(raw[r * 2 * A + a], raw[r * 2 * A + A + a], cfg, mean, sd);
      action[r * A + a] = tanhf(mean + eps[r * A + a] * sd);
    }
    return;
  }
  const float kClamp = 0.99999997f, kLogSqrt2Pi = 0.9189385332046727f, kLog2 = 0.6931471805599453f;
  const int J = cfg.entropy_samples;
  float best = -INFINITY;
  int best_j = 0x7fffffff;
  for (int j = lane; j < J; j += 32) {
    float lp = 0.f;
    for (int a = 0; a < A; ++a) {
      float mean, sd;
      actor_squash(raw[r * 2 * A + a], raw[r * 2 * A + A + a], cfg, mean, sd);
      const float y = tanhf(mean + eps[((long long)j * rows + r) * A + a] * sd);
      const float yc = fminf(fmaxf(y, -kClamp), kClamp);
      const float xh = 0.5f * logf((1.f + yc) / (1.f - yc));
      const float d = xh - mean;
      lp += -(d * d) / (2.f * sd * sd) - logf(sd) - kLogSqrt2Pi - 2.f * (kLog2 - xh - softplusf_(-2.f * xh));
    }
    if (lp > best) { best = lp; best_j = j; }       // (ascending j per lane: keeps the first maximum)
  }
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oj = __shfl_xor_sync(0xffffffffu, best_j, o);
    if (ob > best || (ob == best && oj < best_j)) { best = ob; best_j = oj; }
  }
  for (int a = lane; a < A; a += 32) {
    float mean, sd;
    actor_squash(raw[r * 2 * A + a], raw[r * 2 * A + A + a], cfg, mean, sd);
    action[r * A + a] = tanhf(mean + eps[((long long)best_j * rows + r) * A + a] * sd);
  }
}

// d_raw (rows,2A) from d_action (rows, lda: strided view into d[s;a]) and g_entropy (rows)
__global__ void actor_head_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ eps_a,
                                      const float* __restrict__ action,
                                      const float* __restrict__ dent,
                                      const float* __restrict__ d_action, long long ld_da,
                                      const float* __restrict__ g_entropy, bd_actor_cfg cfg,
                                      int rows, int A, float* __restrict__ d_raw) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * A) return;
  long long r = i / A;
  int a = (int)(i - r * A);
  float m_raw = raw[r * 2 * A + a], s_raw = raw[r * 2 * A + A + a];
  float act = action[i];
  float dy = d_action[r * ld_da + a] * (1.f - act * act);
  float ge = g_entropy ? g_entropy[r] : 0.f;
  float dmean = dy + ge * dent[r * 2 * A + a];
  float dsd = dy * eps_a[i] + ge * dent[r * 2 * A + A + a];
  float th = tanhf(m_raw / cfg.mean_scale);
  d_raw[r * 2 * A + a] = dmean * (1.f - th * th);
  d_raw[r * 2 * A + A + a] = dsd * softplus_gradf_(s_raw + cfg.raw_init_std);
}

// =====================================================================================
// lambda_return (src/dreamer.py:447-471).  One thread per row; T sequential.
//   inputs[t] = r[t] + disc*(1-lam)*next_v[t],  next_v[t] = v[t+1] (t<T-1), bootstrap (t=T-1)
//   R[t] = inputs[t] + disc*lam*R[t+1],  R[T] := bootstrap
// =====================================================================================
__global__ void lambda_return_fwd_kernel(const float* __restrict__ reward,
                                         const float* __restrict__ value,
                                         const float* __restrict__ bootstrap, int T, long long N,
                                         float disc, float lam, float one_minus_lam,
                                         float* __restrict__ ret) {
  long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  // torch evaluates (disc*next)*(1-lam), r + ., (disc*lam)*last, inp + . as separate fp32
  // ops: no FMA contraction here, so the result is bit-identical.
  float last = bootstrap[n];
  float nextv = last;
  const float dl = __fmul_rn(disc, lam);
  for (int t = T - 1; t >= 0; --t) {
    float inp = __fadd_rn(reward[t * N + n], __fmul_rn(__fmul_rn(disc, nextv), one_minus_lam));
    last = __fadd_rn(inp, __fmul_rn(dl, last));
    ret[t * N + n] = last;
    nextv = value[t * N + n];
  }
}
__global__ void lambda_return_bwd_kernel(const float* __restrict__ d_ret, int T, long long N,
                                         float disc, float lam, float* __restrict__ d_reward,
                                         float* __restrict__ d_value,
                                         float* __restrict__ d_bootstrap) {
  long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  // forward-time accumulation of the adjoint of R[t]: G[t] = d_ret[t] + disc*lam*G[t-1]
  float G = 0.f;
  for (int t = 0; t < T; ++t) {
    G = d_ret[t * N + n] + disc * lam * G;
    if (d_reward) d_reward[t * N + n] = G;
    // inputs[t] depends on next_v[t] = value[t+1]  ->  d value[t+1] += G[t]*disc*(1-lam)
    if (d_value) {
      if (t == 0) d_value[n] = 0.f;
      if (t + 1 < T) d_value[(t + 1) * N + n] = G * disc * (1.f - lam);
    }
  }
  // bootstrap enters inputs[T-1] (as next value) and as the initial `last`
  if (d_bootstrap) d_bootstrap[n] = G * disc * (1.f - lam) + G * disc * lam;
}

// =====================================================================================
// small elementwise helpers
// =====================================================================================
__global__ void add2_kernel(const float* __restrict__ a, const float* __restrict__ b,
                            float* __restrict__ out, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = (a ? a[i] : 0.f) + (b ? b[i] : 0.f);
}
// out[r, 0:w] = in[r*ld + c0 : +w] * (rowscale ? rowscale[r] : 1)
__global__ void slice_cols_kernel(const float* __restrict__ in, long long ld, int c0, int w,
                                  const float* __restrict__ rowscale, float* __restrict__ out,
                                  long long rows) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * w) return;
  long long r = i / w;
  int c = (int)(i - r * w);
  float v = in[r * ld + c0 + c];
  if (rowscale) v *= rowscale[r];
  out[i] = v;
}

inline dim3 grid1d(long long n, int block = 256) { return dim3((unsigned)((n + block - 1) / block)); }

// =====================================================================================
// CEM pieces (src/planner.py:53-87)
// =====================================================================================
// actions[h, b, cl, a] = mean[h,b,a] + std[h,b,a] * eps[h, b, c_begin+cl, a]
__global__ void cem_sample_kernel(const float* __restrict__ mean, const float* __restrict__ stdv,
                                  const float* __restrict__ eps, int H, int B, int C, int c_begin,
                                  int Cl, int A, float* __restrict__ actions) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)H * B * Cl * A;
  if (i >= total) return;
  int a = (int)(i % A);
  long long t = i / A;
  int cl = (int)(t % Cl); t /= Cl;
  int b = (int)(t % B);
  int h = (int)(t / B);
  long long ms = ((long long)h * B + b) * A + a;
  actions[i] = mean[ms] + stdv[ms] * eps[(((long long)h * B + b) * C + c_begin + cl) * A + a];
}
// returns[row] = sum_h r[h, row]   (plain sum, no discount; src/planner.py:68-72)
__global__ void cem_sum_rewards_kernel(const float* __restrict__ r, int H, long long rows,
                                       float* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  float s = 0.f;
  for (int h = 0; h < H; ++h) s += r[h * rows + i];
  out[i] = s;
}
// out[(b*Cl + cl), :] = in[b, :]   (expand over candidates, src/planner.py:37-39)
__global__ void cem_expand_kernel(const float* __restrict__ in, int B, int Cl, int D,
                                  float* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * Cl * D) return;
  int d = (int)(i % D);
  int b = (int)(i / ((long long)Cl * D));
  out[i] = in[(long long)b * D + d];
}

// Elite selection + refit.  One CTA per batch row.  Exact top-K:
// elites = the K best under the order (value desc, index asc), i.e. rank(i) < K with
// rank(i) = #{j : v_j > v_i or (v_j == v_i and j < i)}, emitted in ascending
// index order (torch.topk(sorted=False) leaves the order unspecified; the SET is what matters).
// Then mean / population std (unbiased=False) over the K elites per (h, a).
constexpr int CEM_REFIT_THREADS = 1024;
__global__ void __launch_bounds__(CEM_REFIT_THREADS)
cem_refit_kernel(const float* __restrict__ returns, const float* __restrict__ actions, int B, int C,
                 int K, int H, int A, long long* __restrict__ topk_idx, float* __restrict__ mean,
                 float* __restrict__ stdv) {
  // Selection by a bitonic sort of 64-bit keys (value made order-preserving, then ~index so that of
  // equal values the lower index ranks first) -- the same total order as counting
  // rank(i) = #{j : v_j > v_i or (v_j == v_i and j < i)}, in O(log^2) barrier steps instead of C
  // compare iterations per thread (25 us of this kernel at C = 1000 on one CTA).
  extern __shared__ unsigned char smem_raw[];
  unsigned long long* key = reinterpret_cast<unsigned long long*>(smem_raw);   // P = next pow2 >= C
  int P = 1;
  while (P < C) P <<= 1;
  int* flag = reinterpret_cast<int*>(key + P);          // C
  int* elite = flag + C;                                // K
  __shared__ int warp_tot[32];
  const int b = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < P; i += blockDim.x) {
    unsigned long long kk = 0ull;                       // padding sorts last
    if (i < C) {
      const float val = returns[(long long)b * C + i] + 0.f;      // -0 -> +0: equal values tie on the index
      unsigned int u = __float_as_uint(val);
      u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
      kk = ((unsigned long long)u << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned int)i);
      flag[i] = 0;
    }
    key[i] = kk;
  }
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < P; t += blockDim.x) {
        const int x = t ^ j;
        if (x > t) {
          const unsigned long long ka = key[t], kb = key[x];
          const bool desc = (t & k) == 0;
          if (desc ? (ka < kb) : (ka > kb)) { key[t] = kb; key[x] = ka; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < K; i += blockDim.x) flag[0xFFFFFFFFu - (unsigned int)(key[i] & 0xFFFFFFFFull)] = 1;
  __syncthreads();
  // ordered compaction (block-wide exclusive scan over chunks of blockDim.x)
  int base = 0;
  for (int c0 = 0; c0 < C; c0 += blockDim.x) {
    int i = c0 + tid;
    int f = (i < C) ? flag[i] : 0;
    unsigned bal = __ballot_sync(0xffffffffu, f);
    int lane = tid & 31, w = tid >> 5;
    int pre = __popc(bal & ((1u << lane) - 1));
    if (lane == 0) warp_tot[w] = __popc(bal);
    __syncthreads();
    int woff = 0, tot = 0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) {
      int t = warp_tot[k];
      if (k < w) woff += t;
      tot += t;
    }
    if (f) {
      int pos = base + woff + pre;
      elite[pos] = i;
      if (topk_idx) topk_idx[(long long)b * K + pos] = i;
    }
    base += tot;
    __syncthreads();
  }
  // refit: one warp per (h, a), lanes over the K elites (the gather of K scattered actions is an
  // L2 round trip per load: 12 threads walking 2 x 100 dependent-latency loads took 60 of the
  // kernel's 70 us at C = 1000, K = 100)
  const int lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  for (int ha = wid; ha < H * A; ha += nw) {
    int h = ha / A, a = ha - h * A;
    const float* act = actions + (((long long)h * B + b) * C) * A + a;
    float s = 0.f;
    for (int k = lane; k < K; k += 32) s += act[(long long)elite[k] * A];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mu = s / (float)K;
    float q = 0.f;
    for (int k = lane; k < K; k += 32) {
      float d = act[(long long)elite[k] * A] - mu;
      q += d * d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (lane == 0) {
      mean[((long long)h * B + b) * A + a] = mu;
      stdv[((long long)h * B + b) * A + a] = sqrtf(q / (float)K);
    }
  }
}

}  // namespace f32
}  // namespace bd

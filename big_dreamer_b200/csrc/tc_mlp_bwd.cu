// Host side of the tensor-core MLP backward (tc_mlp_bwd.cuh): program / pack tables, scratch
// carving, row chunking, and the three launches (dgrad chain, streaming wgrad, unpack).
#include "api_internal.h"
#include "tc_mlp_bwd.cuh"
#include "tc_builder.cuh"

namespace bd {
namespace tc {

static inline size_t al256(size_t x) { return (x + 255) & ~size_t(255); }

struct BwdPlan {
  int L, k1, k2, Kp_b, Ks, Kp_h, Kp_g;
  int n[BD_MAX_LAYERS], kp_xs[BD_MAX_LAYERS], kp_ds[BD_MAX_LAYERS];
  size_t pack_elems;       // fwd images + transposed dgrad images
  size_t per_tile_bytes;   // scratch images per 128-row tile
  size_t dwp_bytes;        // fp32 gradient images
  bool want_w;
};

static void make_plan(const bd_mlp& m, int k1, int k2, bool want_w, BwdPlan& p) {
  p.L = m.n_layers; p.k1 = k1; p.k2 = k2; p.want_w = want_w;
  p.Kp_b = r16(k1 + 1); p.Ks = k2 > 0 ? r16(k2) : 16;
  p.Kp_h = 16; p.Kp_g = 16; p.pack_elems = 0; p.per_tile_bytes = 0; p.dwp_bytes = 0;
  for (int l = 0; l < p.L; ++l) {
    p.n[l] = m.layer[l].out_features;
    p.kp_xs[l] = r16(p.n[l] + 1);
    p.kp_ds[l] = r16(p.n[l]);
    if (l + 1 < p.L) { p.Kp_h = max(p.Kp_h, p.kp_xs[l]); p.per_tile_bytes += 128 * p.kp_xs[l] * 2; }
    p.Kp_g = max(p.Kp_g, p.kp_ds[l]);
    const int kin = m.layer[l].in_features;
    p.pack_elems += (size_t)r16(p.n[l]) * (l == 0 ? p.Kp_b + p.Ks : r16(kin + 1));   // forward images
    p.pack_elems += (size_t)r16(kin) * r16(p.n[l]);                                   // W^T images
    if (want_w) {
      p.per_tile_bytes += 128 * p.kp_ds[l] * 2;
      p.dwp_bytes += al256((size_t)256 * (l == 0 ? p.Kp_b : r16(kin + 1)) * 4);
      if (l == 0 && k2 > 0) p.dwp_bytes += al256((size_t)256 * p.Ks * 4);
    }
  }
  if (want_w) p.per_tile_bytes += 128 * (p.Kp_b + p.Ks) * 2;
}

bool mlp_backward_supported(const bd_mlp& m, int k1, int k2, int precision) {
  if (!mlp_supported(m, k1, k2, precision)) return false;
  if (k1 + k2 > 256) return false;                    // dX accumulator width
  for (int l = 0; l < m.n_layers; ++l)
    if (m.layer[l].out_features > 255) return false;
  return m.n_layers >= 1;
}

static const int64_t kMaxBwdRows = 1 << 18;
size_t mlp_backward_workspace_bytes(const bd_mlp& m, int k1, int k2, int64_t rows) {
  BwdPlan p;
  make_plan(m, k1, k2, true, p);
  int64_t r = rows < kMaxBwdRows ? rows : kMaxBwdRows;
  size_t tiles = (size_t)((r + 127) / 128);
  return al256(p.pack_elems * 2) + 256 + p.dwp_bytes + tiles * p.per_tile_bytes + 65536;
}

template <typename K>
static cudaError_t launch_bwd_pair(K kernel, unsigned grid, const MlpBwdArgs& ba, cudaStream_t s) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kBwdThreads);
  cfg.dynamicSmemBytes = ba.sm.total; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, ba);
}
template <int FMT>
static int launch_bwd_act(int act, unsigned grid, const MlpBwdArgs& ba, cudaStream_t s) {
#define BD_LAUNCH_BWD(ACTV)                                                                          \
  do {                                                                                               \
    if (ba.c2pair) {                                                                                 \
      set_smem_attr(mlp_bwd_kernel<FMT, ACTV, true>, ba.sm.total);                                   \
      cudaError_t e_ = launch_bwd_pair(mlp_bwd_kernel<FMT, ACTV, true>, grid, ba, s);                \
      if (e_ != cudaSuccess) BD_FAIL(BD_ERR_CUDA, "mlp_backward pair launch: %s", cudaGetErrorString(e_)); \
    } else {                                                                                         \
      set_smem_attr(mlp_bwd_kernel<FMT, ACTV, false>, ba.sm.total);                                  \
      mlp_bwd_kernel<FMT, ACTV, false><<<grid, kBwdThreads, ba.sm.total, s>>>(ba);                   \
    }                                                                                                \
  } while (0)
  switch (act) {
    case BD_ACT_ELU: BD_LAUNCH_BWD(BD_ACT_ELU); break;
    case BD_ACT_RELU: BD_LAUNCH_BWD(BD_ACT_RELU); break;
    case BD_ACT_TANH: BD_LAUNCH_BWD(BD_ACT_TANH); break;
    default: BD_LAUNCH_BWD(BD_ACT_IDENTITY); break;
  }
#undef BD_LAUNCH_BWD
  BD_CUDA_LAUNCH_CHECK();
  return BD_OK;
}

int mlp_backward(const bd_mlp* m, const bd_mlp_bwd_args* a, void* ws, size_t ws_bytes, int precision,
                 bd_stream_t stream, const float* x1b, const float* x2b, int64_t split, int64_t seg_rows,
                 const void* x0b_img, const void* x0s_img) {
  const int k1 = a->k1, k2 = a->k2, L = m->n_layers;
  if (!mlp_backward_supported(*m, k1, k2, precision))
    BD_FAIL(BD_ERR_UNSUPPORTED, "tensor-core mlp_backward: sizes/activation not supported");
  bool want_w = false;
  for (int l = 0; l < L; ++l) want_w |= (a->dw[l] != nullptr) || (a->db[l] != nullptr);
  const bool want_dx = a->dx1 || a->dx2;
  const bool have_saved = a->saved != nullptr;   // hidden images from bd_mlp_forward_save: no recompute
  BwdPlan p;
  make_plan(*m, k1, k2, want_w, p);
  if (have_saved)
    for (int l = 0; l + 1 < L; ++l) p.per_tile_bytes -= 128 * p.kp_xs[l] * 2;
  // layer-0 input images made by the caller (the rollout copies them out of its operand tiles): with the saved hidden
  // images the kernel then needs no inputs at all -- its tile prologue (fp32 rows -> images) disappears
  const bool have_x0 = have_saved && want_w && x0b_img != nullptr && (k2 == 0 || x0s_img != nullptr);
  if (have_x0) p.per_tile_bytes -= 128 * (p.Kp_b + p.Ks) * 2;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int fmt = precision == BD_PREC_FP16 ? 0 : 1;

  // ---- program: forward recompute of the hidden layers, then the dgrad chain
  Builder b;
  for (int l = 0; l < L; ++l) {
    const bd_linear& Lr = m->layer[l];
    const int n = Lr.out_features, Np = r16(n);
    if (l + 1 < L && !have_saved) {
      const int d = b.dcol();
      if (l == 0) {
        uint32_t wb = b.add_pack(Lr.w, k1 + k2, 0, n, Np, p.Kp_b, 0, k1, Lr.b, k1);
        b.add_gemm(wb, Np, p.Kp_b, TILE_BCUR, 0, d, 0);
        if (k2 > 0) {
          uint32_t wsx = b.add_pack(Lr.w, k1 + k2, 0, n, Np, p.Ks, k1, k2, nullptr, -1);
          b.add_gemm(wsx, Np, p.Ks, TILE_SA, 0, d, 1);
        }
      } else {
        const int kin = Lr.in_features, Kp = r16(kin + 1);
        uint32_t w = b.add_pack(Lr.w, kin, 0, n, Np, Kp, 0, kin, Lr.b, kin);
        b.add_gemm(w, Np, Kp, TILE_H, 0, d, 0);
      }
      b.end_phase(EPI_B_ACT_SAVE, 1, n, Np, p.kp_xs[l], d, l, TILE_H);
    }
  }
  // Paired tiles (see MlpBwdArgs::pair): with saved hidden images the chain needs one operand tile per row tile, so two
  // row tiles fit a CTA and their chains are interleaved phase by phase, each phase depending on the epilogue TWO
  // phases back (its own tile's previous layer).
  const bool seg_ = seg_rows > 0 && a->rows % seg_rows == 0;
  const long long total_tiles_ = seg_ ? (a->rows / seg_rows) * ((seg_rows + 127) / 128) : (a->rows + 127) / 128;
  int sms_ = 148;
  {
    int dev_ = 0;
    cudaGetDevice(&dev_);
    cudaDeviceGetAttribute(&sms_, cudaDevAttrMultiProcessorCount, dev_);
  }
  // (measured at 2^17 start states: 1.78 ms paired vs 1.75 ms single -- the 12 epilogue warps, not the MMAs, bound this
  // kernel, and the two tiles' epilogues still run one after the other; kept as an opt-in, BD_BWD_PAIR=1)
  bool pair = false;
  if (const char* e = getenv("BD_BWD_PAIR")) pair = have_saved && atoi(e) != 0;
  (void)total_tiles_; (void)sms_;
  // CTA-pair mode (cta_group::2, see issuer_pair_role): opt-in, BD_TC_PAIR2=1; needs an even tile count per launch
  const bool c2 = getenv("BD_TC_PAIR2") && atoi(getenv("BD_TC_PAIR2")) != 0 && !pair && total_tiles_ % 2 == 0 &&
                  total_tiles_ >= 2;
  const int nsub = pair ? 2 : 1;
  const int sub_tile[2] = {TILE_H2, TILE_H};
  const int dep = pair ? 2 : 1;
  for (int sb = 0; sb < nsub; ++sb) {
    b.end_phase(EPI_B_LOAD_DY, 1, m->layer[L - 1].out_features, 0, 0, 0, L - 1, sub_tile[sb]);
    b.prog.p[b.prog.n_phases - 1].pad = (uint8_t)sb;
  }
  for (int l = L - 1; l >= 0; --l) {
    const bd_linear& Lr = m->layer[l];
    const int kin = Lr.in_features, n = Lr.out_features;
    if (l == 0 && !want_dx) break;
    const int Np = r16(kin), Kp = r16(n);
    if (b.pack.njobs >= kMaxPackJobs) BD_FAIL(BD_ERR_UNSUPPORTED, "mlp_backward: too many layers");
    // W^T image: packed(n_in, k_out) = W[k_out, n_in]
    PackJob& j = b.pack.job[b.pack.njobs++];
    j = PackJob{};
    j.w = Lr.w; j.bias = nullptr; j.dst_off = b.w_elems; j.ld = kin; j.row0 = 0; j.N = kin; j.Np = Np;
    j.Kp = Kp; j.bias_k = -1; j.nseg = 1; j.seg[0] = {0, 0, n}; j.transpose = 1;
    const uint32_t woff = (uint32_t)b.w_elems;
    b.w_elems += (long long)Np * Kp;
    for (int sb = 0; sb < nsub; ++sb) {
      const int d = b.dcol();
      b.add_gemm(woff, Np, Kp, sub_tile[sb], 0, d, 0);
      if (l > 0) b.end_phase(EPI_B_DACT, dep, kin, Np, p.kp_ds[l - 1], d, l - 1, sub_tile[sb]);
      else b.end_phase(EPI_B_DX, dep, kin, Np, 0, d, 0, sub_tile[sb]);
      b.prog.p[b.prog.n_phases - 1].pad = (uint8_t)sb;
    }
  }
  if (!b.ok) BD_FAIL(BD_ERR_UNSUPPORTED, "tensor-core mlp_backward: program too large");

  // ---- workspace carve
  char* base = static_cast<char*>(ws);
  size_t off = 0;
  auto take = [&](size_t bytes) { char* r = base + off; off += al256(bytes); return r; };
  uint16_t* wpack = reinterpret_cast<uint16_t*>(take((size_t)b.w_elems * 2));
  unsigned int* amax = reinterpret_cast<unsigned int*>(take(256));
  cudaMemsetAsync(amax, 0, 256, s);
  {
    const long long n = a->rows * m->layer[L - 1].out_features;
    long long g = (n + 255) / 256;
    if (g > 1184) g = 1184;
    absmax_kernel<<<(unsigned)(g < 1 ? 1 : g), 256, 0, s>>>(a->dy, n, amax);
    BD_CUDA_LAUNCH_CHECK();
  }
  float* dwp[BD_MAX_LAYERS][2] = {};
  int dwp_kp[BD_MAX_LAYERS][2] = {};
  if (want_w) {
    const size_t start = off;
    for (int l = 0; l < L; ++l) {
      dwp_kp[l][0] = (l == 0) ? p.Kp_b : r16(m->layer[l].in_features + 1);
      dwp[l][0] = reinterpret_cast<float*>(take((size_t)256 * dwp_kp[l][0] * 4));
      if (l == 0 && k2 > 0) {
        dwp_kp[l][1] = p.Ks;
        dwp[l][1] = reinterpret_cast<float*>(take((size_t)256 * p.Ks * 4));
      }
    }
    cudaMemsetAsync(base + start, 0, off - start, s);
  }
  if (off + p.per_tile_bytes + 4096 > ws_bytes)
    BD_FAIL(BD_ERR_WORKSPACE, "tensor-core mlp_backward: workspace %zu too small", ws_bytes);
  long long chunk_tiles = (long long)((ws_bytes - off - 4096) / (p.per_tile_bytes ? p.per_tile_bytes : 1));
  // segmented tiling (see MlpBwdArgs): rows = T steps of seg_rows rows, each step tiled on its own
  const bool seg = seg_rows > 0 && a->rows % seg_rows == 0;
  const long long seg_tiles = seg ? (seg_rows + 127) / 128 : 0;
  const long long total_tiles = seg ? (a->rows / seg_rows) * seg_tiles : (a->rows + 127) / 128;
  if (chunk_tiles > total_tiles) chunk_tiles = total_tiles;
  if (c2) chunk_tiles &= ~1LL;       // (CTA pairs: every launch takes an even number of tiles)
  if (chunk_tiles < 1) BD_FAIL(BD_ERR_WORKSPACE, "tensor-core mlp_backward: workspace too small");
  char* scratch = base + off;

  // ---- pack weights once
  if (c2)
    for (int i = 0; i < b.pack.njobs; ++i) b.pack.job[i].split2 = 1;
  {
    long long max_img = 0;
    for (int i = 0; i < b.pack.njobs; ++i)
      max_img = max(max_img, (long long)b.pack.job[i].Np * b.pack.job[i].Kp);
    long long pgx = (max_img + 255) / 256;
    if (pgx > 64) pgx = 64;
    dim3 pgrid((unsigned)pgx, (unsigned)b.pack.njobs);
    if (fmt == 0) pack_weights_kernel<0><<<pgrid, 256, 0, s>>>(b.pack, wpack);
    else pack_weights_kernel<1><<<pgrid, 256, 0, s>>>(b.pack, wpack);
    BD_CUDA_LAUNCH_CHECK();
  }

  MlpBwdArgs ba{};
  {
    // tiles: B0 | SA | H | G
    SmemPlan& sm = ba.sm;
    uint32_t o = 0;
    auto tk = [&](uint32_t bytes) { uint32_t r = o; o += (bytes + 1023) & ~1023u; return r; };
    if (pair) {        // one G tile per sub-tile (TILE_H2, TILE_H); no forward recompute, hence no input / hidden tiles
      sm.off_tile[4] = tk(kTileRows * p.Kp_g * 2);
      sm.off_tile[3] = tk(kTileRows * p.Kp_g * 2);
      sm.off_tile[0] = sm.off_tile[1] = sm.off_tile[2] = sm.off_tile[4];
    } else {
    sm.off_tile[0] = tk(kTileRows * p.Kp_b * 2);
    sm.off_tile[1] = sm.off_tile[0];
    sm.off_tile[2] = tk(kTileRows * p.Ks * 2);
    sm.off_tile[3] = tk(kTileRows * p.Kp_h * 2);
    sm.off_tile[4] = tk(kTileRows * p.Kp_g * 2);
    }
    sm.stage_bytes = align_stage(b.max_stage);
    sm.off_ring = o;
    const uint32_t budget = 227 * 1024 - 4096;
    if (o + 2 * sm.stage_bytes > budget)
      BD_FAIL(BD_ERR_UNSUPPORTED, "tensor-core mlp_backward: tiles do not fit shared memory");
    sm.nstage = min(8u, (budget - o) / sm.stage_bytes);
    sm.total = o + sm.nstage * sm.stage_bytes + 1024;
  }
  // (CTA pair: a stage holds HALF of the weight rows, so it takes twice the K columns)
  b.finalize_blocks(c2 ? 2 * ba.sm.stage_bytes : ba.sm.stage_bytes);
  ba.prog = b.prog;
  ba.wpack = wpack; ba.T = 1; ba.prof = nullptr; ba.amax_bits = amax;
  ba.k1 = k1; ba.k2 = k2; ba.out = m->layer[L - 1].out_features; ba.n_layers = L; ba.act = m->activation;
  ba.Kp_b = p.Kp_b; ba.Ks = p.Ks; ba.Kp_h = p.Kp_h; ba.Kp_g = p.Kp_g; ba.want_images = want_w ? 1 : 0;
  ba.need_x = ((want_w && !have_x0) || !have_saved) ? 1 : 0;
  for (int l = 0; l < L; ++l) { ba.kp_xs[l] = p.kp_xs[l]; ba.kp_ds[l] = p.kp_ds[l]; }

  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int out = ba.out;

  for (long long t0 = 0; t0 < total_tiles; t0 += chunk_tiles) {
    const long long nt = chunk_tiles < total_tiles - t0 ? chunk_tiles : total_tiles - t0;
    const long long r0 = t0 * 128;
    long long nrows = a->rows - r0;
    if (nrows > nt * 128) nrows = nt * 128;
    // scratch images for this chunk
    size_t so = 0;
    auto st = [&](size_t bytes) { char* r = scratch + so; so += bytes; return reinterpret_cast<uint16_t*>(r); };
    if (have_saved) {
      const size_t all_tiles = (size_t)total_tiles;
      const char* sb = static_cast<const char*>(a->saved);
      for (int l = 0; l + 1 < L; ++l) {
        ba.xs[l] = reinterpret_cast<uint16_t*>(const_cast<char*>(sb) + (size_t)t0 * 128 * p.kp_xs[l] * 2);
        sb += all_tiles * 128 * p.kp_xs[l] * 2;
      }
    } else {
      for (int l = 0; l + 1 < L; ++l) ba.xs[l] = st((size_t)nt * 128 * p.kp_xs[l] * 2);
    }
    for (int l = 0; l < L; ++l) ba.ds[l] = want_w ? st((size_t)nt * 128 * p.kp_ds[l] * 2) : nullptr;
    if (!want_w) ba.ds[L - 1] = nullptr;
    if (have_x0) {
      ba.x0b = reinterpret_cast<uint16_t*>(const_cast<char*>(static_cast<const char*>(x0b_img)) + (size_t)t0 * 128 * p.Kp_b * 2);
      ba.x0s = k2 > 0 ? reinterpret_cast<uint16_t*>(const_cast<char*>(static_cast<const char*>(x0s_img)) + (size_t)t0 * 128 * p.Ks * 2)
                      : nullptr;
    } else {
    ba.x0b = want_w ? st((size_t)nt * 128 * p.Kp_b * 2) : nullptr;
    ba.x0s = want_w ? st((size_t)nt * 128 * p.Ks * 2) : nullptr;
    }
    ba.N = nrows; ba.ntiles = nt;
    ba.pair = pair ? 1 : 0; ba.nloop = pair ? (nt + 1) / 2 : nt;
    ba.c2pair = (c2 && nt % 2 == 0) ? 1 : 0;
    if (c2 && !ba.c2pair) BD_FAIL(BD_ERR_UNSUPPORTED, "mlp_backward: CTA-pair mode needs an even tile count per launch");
    if (ba.c2pair) ba.nloop = nt / 2;
#ifdef BD_BWD_DBG
    ba.dbg = getenv("BD_BWD_DBGV") ? atoi(getenv("BD_BWD_DBGV")) : 0;
#endif
    ba.seg_tiles = (int)seg_tiles; ba.seg_rows = seg ? seg_rows : 0; ba.tile_base = seg ? t0 : 0;
    if (seg) {      // row pointers stay those of row 0: the kernel maps (tile_base + tile) to its rows
      ba.N = a->rows;
      ba.x1 = a->x1; ba.x2 = a->x2;
    } else {
    ba.x1 = a->x1 + r0 * k1;
    ba.x2 = a->x2 ? a->x2 + r0 * k2 : nullptr;
    }
    // two-segment input: rows [0, split) from x1 / x2, the rest from x1b / x2b (chunk-relative here)
    const bool two_seg = x1b != nullptr && split >= 0 && split < a->rows;
    if (seg) {
      ba.split = two_seg ? split : a->rows;
      ba.x1b = two_seg ? x1b : ba.x1;
      ba.x2b = (two_seg && x2b) ? x2b : ba.x2;
      ba.dy = a->dy; ba.dx1 = a->dx1; ba.dx2 = a->dx2;
    } else {
    ba.split = two_seg ? (split > r0 ? split - r0 : 0) : nrows;
    ba.x1b = two_seg ? x1b + (r0 > split ? (r0 - split) * k1 : 0) : ba.x1;
    ba.x2b = (two_seg && x2b) ? x2b + (r0 > split ? (r0 - split) * k2 : 0) : ba.x2;
    ba.dy = a->dy + r0 * out;
    ba.dx1 = a->dx1 ? a->dx1 + r0 * k1 : nullptr;
    ba.dx2 = a->dx2 ? a->dx2 + r0 * k2 : nullptr;
    }
    unsigned grid = (unsigned)(ba.nloop < sms ? ba.nloop : sms);
    if (ba.c2pair) grid = 2u * (unsigned)(ba.nloop < sms / 2 ? ba.nloop : sms / 2);
    {
      // the producer warp pulls the NEXT tile's inputs (rows of x1 / x2 / dy, saved hidden images) into L2
      PrefetchPlan& pf = ba.pf;
      pf.n = 0; pf.reverse = 0;
      auto add = [&](const void* ptr, size_t tile_bytes) {      // (paired tiles: a loop item = two consecutive tiles)
        if (!ptr || pf.n >= 6) return;
        tile_bytes *= (size_t)nsub;
        pf.base[pf.n] = static_cast<const char*>(ptr); pf.step_stride[pf.n] = 0;
        pf.tile_stride[pf.n] = (long long)tile_bytes; pf.bytes[pf.n] = (unsigned int)tile_bytes; ++pf.n;
      };
      if (ba.need_x && !two_seg && !seg) {     // (the L2 prefetch assumes one contiguous input per tile index)
        add(ba.x1, (size_t)128 * k1 * 4);
        add(ba.x2, (size_t)128 * k2 * 4);
      }
      if (!seg) add(ba.dy, (size_t)128 * out * 4);
      if (have_saved)
        for (int l = 0; l + 1 < L && pf.n < 6; ++l) add(ba.xs[l], (size_t)128 * p.kp_xs[l] * 2);
    }
    {
      ProfScope ps(BD_PROF_MLP_BWD, s);
      if (fmt == 0) BD_TRY(launch_bwd_act<0>(m->activation, grid, ba, s));
      else BD_TRY(launch_bwd_act<1>(m->activation, grid, ba, s));
    }

    if (want_w) {
      WgradArgs wa{};
      int nj = 0;
      uint32_t max_x = 0, max_dy = 0;
      for (int l = 0; l < L; ++l) {
        if (!a->dw[l] && !a->db[l]) continue;
        for (int part = 0; part < 2; ++part) {
          if (!dwp[l][part]) continue;
          WgradJob& j = wa.job[nj++];
          j.dyimg = ba.ds[l]; j.kp_dy = p.kp_ds[l]; j.m_valid = m->layer[l].out_features;
          j.dwp = dwp[l][part]; j.kp_x = dwp_kp[l][part];
          j.ximg = (l == 0) ? (part == 0 ? ba.x0b : ba.x0s) : ba.xs[l - 1];
          max_x = max(max_x, (uint32_t)j.kp_x);
          max_dy = max(max_dy, (uint32_t)j.kp_dy);
        }
      }
      if (nj > 0) {
        wa.ntiles = nt;
        wa.amax_bits = amax;
        // a stage = [dY image | X image], packed (the widest of each over the jobs); two stages so the copies of
        // tile i + 1 run under the MMAs of tile i (a stage with a fixed 64 KB dY slot left room for one: the
        // kernel alternated between loading and contracting and streamed at 4.1 TB/s)
        wa.x_off = (128 * max_dy * 2 + 1023) & ~1023u;
        wa.stage_bytes = wa.x_off + ((128 * max_x * 2 + 1023) & ~1023u);
        wa.nstage = (2 * wa.stage_bytes <= 220 * 1024) ? 2 : 1;
        // the second M tile of a 129..256-feature dY reads 64 KB from the stage start (see wgrad_kernel)
        const size_t smem = max((size_t)wa.nstage * wa.stage_bytes,
                                (size_t)(wa.nstage - 1) * wa.stage_bytes + 128 * 256 * 2);
        long long per = (sms + nj - 1) / nj;
        if (per > nt) per = nt;
        if (per < 1) per = 1;
        dim3 wgrid((unsigned)per, (unsigned)nj);
        ProfScope ps(BD_PROF_WGRAD, s);
        if (fmt == 0) {
          set_smem_attr(wgrad_kernel<0>, smem);
          wgrad_kernel<0><<<wgrid, 128, smem, s>>>(wa);
        } else {
          set_smem_attr(wgrad_kernel<1>, smem);
          wgrad_kernel<1><<<wgrid, 128, smem, s>>>(wa);
        }
        BD_CUDA_LAUNCH_CHECK();
      }
    }
  }

  if (want_w) {
    UnpackTable ut{};
    for (int l = 0; l < L; ++l) {
      if (!a->dw[l] && !a->db[l]) continue;
      for (int part = 0; part < 2; ++part) {
        if (!dwp[l][part]) continue;
        UnpackJob& u = ut.job[ut.njobs++];
        const bd_linear& Lr = m->layer[l];
        u.dwp = dwp[l][part]; u.dw = a->dw[l]; u.db = part == 0 ? a->db[l] : nullptr;
        u.kp = dwp_kp[l][part]; u.n = Lr.out_features; u.ld = Lr.in_features; u.nseg = 1;
        if (l == 0) {
          u.seg[0] = part == 0 ? PackSeg{0, 0, k1} : PackSeg{0, k1, k2};
          u.bias_k = part == 0 ? k1 : -1;
        } else {
          u.seg[0] = PackSeg{0, 0, Lr.in_features};
          u.bias_k = Lr.in_features;
        }
      }
    }
    if (ut.njobs > 0) {
      unpack_dw_kernel<<<dim3(32, ut.njobs), 256, 0, s>>>(ut);
      BD_CUDA_LAUNCH_CHECK();
    }
  }
  return BD_OK;
}

// ---------------------------------------------------------------------------------------------
// Backward of two scalar heads on the same rows (the pair of bd_heads_forward; frozen heads: only d x1 / d x2) as
// ONE launch: the two dgrad chains are interleaved phase by phase (each depending on the epilogue two phases back:
// the MMAs of one head run under the epilogue of the other), each in its own G tile and TMEM half, and their last
// GEMMs accumulate into ONE dX accumulator -- the sum of the two heads' input gradients, which is what flows into
// the beliefs / states (src/dreamer.py:321-335).
// ---------------------------------------------------------------------------------------------
size_t heads_pair_backward_workspace_bytes(const bd_mlp& reward, const bd_mlp& value, int k1, int k2) {
  BwdPlan pr, pv;
  make_plan(reward, k1, k2, false, pr);
  make_plan(value, k1, k2, false, pv);
  return al256((pr.pack_elems + pv.pack_elems) * 2) + 256 + 65536;
}
int heads_pair_backward(const bd_mlp* reward, const bd_mlp* value, int k1, int k2, int64_t rows,
                        const float* dy_reward, const float* dy_value, const void* saved_reward,
                        const void* saved_value, float* dx1, float* dx2, void* ws, size_t ws_bytes, int precision,
                        bd_stream_t stream) {
  const bd_mlp* hs[2] = {reward, value};
  const int L = reward->n_layers;
  if (!mlp_backward_supported(*reward, k1, k2, precision) || !mlp_backward_supported(*value, k1, k2, precision) ||
      value->n_layers != L || L < 2 || reward->activation != value->activation)
    BD_FAIL(BD_ERR_UNSUPPORTED, "tensor-core heads_pair_backward: configuration not supported");
  for (int l = 0; l < L; ++l)
    if (reward->layer[l].out_features != value->layer[l].out_features ||
        reward->layer[l].in_features != value->layer[l].in_features)
      BD_FAIL(BD_ERR_UNSUPPORTED, "tensor-core heads_pair_backward: the two heads differ in shape");
  BwdPlan p;
  make_plan(*reward, k1, k2, false, p);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int fmt = precision == BD_PREC_FP16 ? 0 : 1;

  Builder b;
  const int sub_tile[2] = {TILE_H2, TILE_H};
  for (int sb = 0; sb < 2; ++sb) {
    b.end_phase(EPI_B_LOAD_DY, 1, 1, 0, 0, 0, L - 1, sub_tile[sb]);
    b.prog.p[b.prog.n_phases - 1].pad = (uint8_t)sb;
  }
  auto pack_T = [&](const bd_linear& Lr) -> uint32_t {     // W^T image: packed(n_in, k_out) = W[k_out, n_in]
    const int kin = Lr.in_features, n = Lr.out_features;
    PackJob& j = b.pack.job[b.pack.njobs++];
    j = PackJob{};
    j.w = Lr.w; j.bias = nullptr; j.dst_off = b.w_elems; j.ld = kin; j.row0 = 0; j.N = kin; j.Np = r16(kin);
    j.Kp = r16(n); j.bias_k = -1; j.nseg = 1; j.seg[0] = {0, 0, n}; j.transpose = 1;
    const uint32_t woff = (uint32_t)b.w_elems;
    b.w_elems += (long long)r16(kin) * r16(n);
    return woff;
  };
  if (2 * L + 2 > kMaxPackJobs) BD_FAIL(BD_ERR_UNSUPPORTED, "heads_pair_backward: too many layers");
  for (int l = L - 1; l >= 1; --l)
    for (int sb = 0; sb < 2; ++sb) {
      const bd_linear& Lr = hs[sb]->layer[l];
      const int kin = Lr.in_features, n = Lr.out_features;
      const uint32_t woff = pack_T(Lr);
      const int d = b.dcol();
      b.add_gemm(woff, r16(kin), r16(n), sub_tile[sb], 0, d, 0);
      b.end_phase(EPI_B_DACT, 2, kin, r16(kin), p.kp_ds[l - 1], d, l - 1, sub_tile[sb]);
      b.prog.p[b.prog.n_phases - 1].pad = (uint8_t)sb;
    }
  {   // dX = dY_0(reward) W_0(reward) + dY_0(value) W_0(value): two GEMMs into one accumulator, one epilogue
    const int kin = k1 + k2, n = reward->layer[0].out_features;
    const int d = b.dcol();
    const uint32_t w0 = pack_T(reward->layer[0]), w1 = pack_T(value->layer[0]);
    b.add_gemm(w0, r16(kin), r16(n), sub_tile[0], 0, d, 0);
    b.add_gemm(w1, r16(kin), r16(n), sub_tile[1], 0, d, 1);
    b.end_phase(EPI_B_DX, 1, kin, r16(kin), 0, d, 0, sub_tile[0]);
    b.prog.g[b.prog.n_gemms - 2].dep_back = 2;      // the first head's last epilogue is two phases back
  }
  if (!b.ok) BD_FAIL(BD_ERR_UNSUPPORTED, "tensor-core heads_pair_backward: program too large");

  char* base = static_cast<char*>(ws);
  size_t off = 0;
  auto take = [&](size_t bytes) { char* r = base + off; off += al256(bytes); return r; };
  uint16_t* wpack = reinterpret_cast<uint16_t*>(take((size_t)b.w_elems * 2));
  unsigned int* amax = reinterpret_cast<unsigned int*>(take(256));
  if (off + 4096 > ws_bytes) BD_FAIL(BD_ERR_WORKSPACE, "tensor-core heads_pair_backward: workspace %zu too small", ws_bytes);
  cudaMemsetAsync(amax, 0, 256, s);
  {   // one power-of-two scale for both heads' upstream gradients
    AbsmaxJobs jobs{};
    jobs.x[0] = dy_reward; jobs.n[0] = rows; jobs.x[1] = dy_value; jobs.n[1] = rows;
    long long g = (rows + 255) / 256;
    if (g > 592) g = 592;
    absmax_multi_kernel<<<dim3((unsigned)(g < 1 ? 1 : g), 2), 256, 0, s>>>(jobs, amax);
    BD_CUDA_LAUNCH_CHECK();
  }
  {
    long long max_img = 0;
    for (int i = 0; i < b.pack.njobs; ++i)
      max_img = max(max_img, (long long)b.pack.job[i].Np * b.pack.job[i].Kp);
    long long pgx = (max_img + 255) / 256;
    if (pgx > 64) pgx = 64;
    dim3 pgrid((unsigned)pgx, (unsigned)b.pack.njobs);
    if (fmt == 0) pack_weights_kernel<0><<<pgrid, 256, 0, s>>>(b.pack, wpack);
    else pack_weights_kernel<1><<<pgrid, 256, 0, s>>>(b.pack, wpack);
    BD_CUDA_LAUNCH_CHECK();
  }
  MlpBwdArgs ba{};
  {
    SmemPlan& sm = ba.sm;
    uint32_t o = 0;
    auto tk = [&](uint32_t bytes) { uint32_t r = o; o += (bytes + 1023) & ~1023u; return r; };
    sm.off_tile[4] = tk(kTileRows * p.Kp_g * 2);          // G tile of the first head (TILE_H2)
    sm.off_tile[3] = tk(kTileRows * p.Kp_g * 2);          // G tile of the second head (TILE_H)
    sm.off_tile[0] = sm.off_tile[1] = sm.off_tile[2] = sm.off_tile[4];
    // (room for the dX tile staged row-major over both dead G tiles by the EPI_B_DX epilogue)
    const int st1 = k1 + ((12 - k1 % 8) % 8), st2 = k2 | 1;
    const uint32_t stage_dx = (uint32_t)kTileRows * (st1 + st2) * 4;
    if (o < stage_dx) o = (stage_dx + 1023) & ~1023u;
    sm.stage_bytes = align_stage(b.max_stage);
    sm.off_ring = o;
    const uint32_t budget = 227 * 1024 - 4096;
    if (o + 2 * sm.stage_bytes > budget)
      BD_FAIL(BD_ERR_UNSUPPORTED, "tensor-core heads_pair_backward: tiles do not fit shared memory");
    sm.nstage = min(8u, (budget - o) / sm.stage_bytes);
    sm.total = o + sm.nstage * sm.stage_bytes + 1024;
  }
  b.finalize_blocks(ba.sm.stage_bytes);
  ba.prog = b.prog;
  ba.wpack = wpack; ba.T = 1; ba.prof = nullptr; ba.amax_bits = amax;
  ba.k1 = k1; ba.k2 = k2; ba.out = 1; ba.n_layers = L; ba.act = reward->activation;
  ba.Kp_b = p.Kp_b; ba.Ks = p.Ks; ba.Kp_h = p.Kp_h; ba.Kp_g = p.Kp_g; ba.want_images = 0; ba.need_x = 0;
  for (int l = 0; l < L; ++l) { ba.kp_xs[l] = p.kp_xs[l]; ba.kp_ds[l] = p.kp_ds[l]; }
  const long long nt = (rows + 127) / 128;
  ba.N = rows; ba.ntiles = nt; ba.pair = 2; ba.nloop = nt; ba.c2pair = 0;
  ba.seg_tiles = 0; ba.seg_rows = 0; ba.tile_base = 0; ba.split = rows;
  ba.dy = dy_reward; ba.dy_b = dy_value; ba.dx1 = dx1; ba.dx2 = dx2;
  {
    const char* sr = static_cast<const char*>(saved_reward);
    const char* sv = static_cast<const char*>(saved_value);
    for (int l = 0; l + 1 < L; ++l) {
      ba.xs[l] = reinterpret_cast<uint16_t*>(const_cast<char*>(sr));
      ba.xs_b[l] = reinterpret_cast<uint16_t*>(const_cast<char*>(sv));
      sr += (size_t)nt * 128 * p.kp_xs[l] * 2;
      sv += (size_t)nt * 128 * p.kp_xs[l] * 2;
    }
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = (unsigned)(nt < sms ? nt : sms);
  {
    PrefetchPlan& pf = ba.pf;
    pf.n = 0; pf.reverse = 0;
    auto add = [&](const void* ptr, size_t tile_bytes) {
      if (!ptr || pf.n >= 6) return;
      pf.base[pf.n] = static_cast<const char*>(ptr); pf.step_stride[pf.n] = 0;
      pf.tile_stride[pf.n] = (long long)tile_bytes; pf.bytes[pf.n] = (unsigned int)tile_bytes; ++pf.n;
    };
    for (int l = 0; l + 1 < L && pf.n < 6; ++l) {
      add(ba.xs[l], (size_t)128 * p.kp_xs[l] * 2);
      add(ba.xs_b[l], (size_t)128 * p.kp_xs[l] * 2);
    }
  }
  ProfScope ps(BD_PROF_MLP_BWD, s);
  if (fmt == 0) BD_TRY(launch_bwd_act<0>(reward->activation, grid, ba, s));
  else BD_TRY(launch_bwd_act<1>(reward->activation, grid, ba, s));
  return BD_OK;
}

}  // namespace tc
}  // namespace bd

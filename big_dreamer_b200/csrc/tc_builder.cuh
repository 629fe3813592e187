// Host-side builder of engine programs and weight-packing tables (shared by all tensor-core
// entry points).
#pragma once
#include "tc_engine.cuh"
#include <stdlib.h>
#include <vector>

namespace bd {
namespace tc {

inline int r16(int x) { return (x + 15) / 16 * 16; }
// K columns of the widest GEMM that one weight-ring stage holds (BD_TC_KC overrides, multiples of 16), and the
// stage size rounding (descriptors and bulk copies need 16 bytes; 256 keeps the stages on 128-byte lines)
inline uint32_t stage_kc() {
  static const uint32_t v = [] {
    const char* e = getenv("BD_TC_KC");
    const int k = e ? atoi(e) : 32;
    return (uint32_t)((k >= 16 && k <= 256) ? k / 16 * 16 : 32);
  }();
  return v;
}
inline uint32_t align_stage(uint32_t bytes) { return (bytes + 255u) & ~255u; }

struct Builder {
  PackTable pack{};
  Program prog{};
  long long w_elems = 0;
  int cur_phase_g0 = 0;
  uint32_t max_stage = 0;
  bool ok = true;

  // packs rows [row0,row0+n) of w (ld cols) into an image (Np x Kp); returns element offset
  uint32_t add_pack(const float* w, int ld, int row0, int n, int Np, int Kp, int src_c0, int len,
                    const float* bias, int bias_k) {
    if (pack.njobs >= kMaxPackJobs) { ok = false; return 0; }
    PackJob& j = pack.job[pack.njobs++];
    j = PackJob{};
    j.w = w; j.bias = bias; j.dst_off = w_elems; j.ld = ld; j.row0 = row0; j.N = n; j.Np = Np;
    j.Kp = Kp; j.bias_k = bias ? bias_k : -1; j.nseg = 1; j.seg[0] = {0, src_c0, len}; j.transpose = 0;
    uint32_t off = (uint32_t)w_elems;
    w_elems += (long long)Np * Kp;
    return off;
  }
  // image whose rows are stacked from up to three source row ranges (same column mapping)
  uint32_t add_pack_rows(const float* w, int ld, int nrseg, const PackSeg* rsegs, int Np, int Kp,
                         int src_c0, int len, const float* bias, int bias_k) {
    uint32_t off = add_pack(w, ld, 0, 0, Np, Kp, src_c0, len, bias, bias_k);
    if (!ok) return off;
    PackJob& j = pack.job[pack.njobs - 1];
    j.nrseg = nrseg;
    for (int i = 0; i < nrseg; ++i) j.rseg[i] = rsegs[i];
    return off;
  }
  void add_gemm(uint32_t w_off, int Np, int Kp, int a_tile, int a_k0, int d_col, int accumulate) {
    if (prog.n_gemms >= kMaxGemms) { ok = false; return; }
    Gemm& g = prog.g[prog.n_gemms++];
    g.w_off = w_off; g.Np = (uint16_t)Np; g.Kp = (uint16_t)Kp; g.a_k0 = (uint16_t)a_k0;
    g.d_col = (uint16_t)d_col; g.a_tile = (uint8_t)a_tile; g.accumulate = (uint8_t)accumulate;
    g.kc = 32; g.pad = 0; g.dep_back = 1;
    max_stage = max(max_stage, (uint32_t)Np * stage_kc() * 2);
  }
  // K columns per ring stage: as many as fit (narrow GEMMs move their whole K in one or two copies)
  void finalize_blocks(uint32_t stage_bytes) { finalize_blocks_of(prog, stage_bytes); }
  static void finalize_blocks_of(Program& prog, uint32_t stage_bytes) {
    for (int i = 0; i < prog.n_gemms; ++i) {
      Gemm& g = prog.g[i];
      uint32_t kc = stage_bytes / (g.Np * 2u) / 16u * 16u;
      if (kc > g.Kp) kc = g.Kp;
      if (kc < 16) kc = 16;
      g.kc = (uint16_t)kc;
    }
  }
  void end_phase(int epi, int dep_back, int n_valid, int Np, int Kp_out, int d_col, int aux0,
                 int out_tile) {
    if (prog.n_phases >= kMaxPhases) { ok = false; return; }
    Phase& p = prog.p[prog.n_phases++];
    p.g0 = (uint8_t)cur_phase_g0; p.ng = (uint8_t)(prog.n_gemms - cur_phase_g0); p.epi = (uint8_t)epi;
    p.dep_back = (uint8_t)dep_back; p.n_valid = (uint16_t)n_valid; p.Np = (uint16_t)Np;
    p.Kp_out = (uint16_t)Kp_out; p.d_col = (uint16_t)d_col; p.aux0 = (uint16_t)aux0;
    p.out_tile = (uint8_t)out_tile; p.pad = 0; p.n_sub = 1; p.pad2 = 0; p.split = 0; p.col0 = 0; p.pad3 = 0;
    for (int gi = cur_phase_g0; gi < prog.n_gemms; ++gi) prog.g[gi].dep_back = (uint8_t)dep_back;
    for (int gi : chained) prog.g[gi].dep_back = 2;     // the previous phase published 2 sub-epilogues
    chained.clear();
    cur_phase_g0 = prog.n_gemms;
  }
  int dcol() const { return (prog.n_phases & 1) * 256; }

  // ---- column-split (cluster) builds: one program per rank, shared pack table / weight buffer
  Program ranks[kMaxRanks];
  void end_rank(int r) {
    ranks[r] = prog;
    prog = Program{};
    cur_phase_g0 = 0;
    chained.clear();
  }

  // Sub-epilogue pipelining between two chained layers.  Call right after end_phase() of an ACT
  // epilogue that writes an operand tile of Kp_out columns: it will publish columns [0, split)
  // first.  chain_gemm() then adds the consumer GEMM as two K-slabs, the first of which only
  // depends on that early publication.
  int split_last_phase() {
    Phase& p = prog.p[prog.n_phases - 1];
    const int split = (p.Kp_out / 2 + 31) / 32 * 32;
    if (getenv("BD_TC_NOSPLIT") || p.Kp_out < 96 || split >= p.Kp_out) return 0;
    p.n_sub = 2; p.split = (uint16_t)split;
    return split;
  }
  // consumer of the operand tile written by the previous phase (K = Kp columns from a_tile)
  void chain_gemm(uint32_t w_off, int Np, int Kp, int a_tile, int d_col, int split) {
    if (split <= 0 || split >= Kp) { add_gemm(w_off, Np, Kp, a_tile, 0, d_col, 0); return; }
    add_gemm(w_off, Np, split, a_tile, 0, d_col, 0);
    chained.push_back(prog.n_gemms - 1);
    add_gemm(w_off + (uint32_t)split * Np, Np, Kp - split, a_tile, split, d_col, 1);
  }
  std::vector<int> chained;   // GEMMs that may start after the FIRST sub-epilogue of the previous phase
};


}  // namespace tc
}  // namespace bd

// Weight packing for the tensor-core path: nn.Linear weights (out,in) fp32 row-major ->
// 16-bit KM8 images (tc_common.cuh) that a bulk TMA copy can drop into shared memory as-is.
// The bias is folded into the contraction: the activation tiles carry a constant-1 column and
// the packed weight carries the bias in the matching K column.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace bd {
namespace tc {

struct PackSeg { int dst_k0, src_c0, len; };     // dst cols [dst_k0, +len) <- src cols [src_c0, +len)
struct PackJob {
  const float* w;       // source weight, row-major, leading dim ld
  const float* bias;    // optional (N) -> column bias_k
  long long dst_off;    // element offset of this matrix in the packed buffer
  int ld, row0, N;      // rows [row0, row0+N) of the source
  int Np, Kp;           // padded rows (mult of 16) and cols (mult of 16) of the packed image
  int bias_k;           // -1: none
  int nseg;
  PackSeg seg[3];
  int transpose;        // 1: packed(n, k) = w[(src row = k-mapped), (src col = row0 + n)]  (dgrad)
  int nrseg;            // > 0: image rows are stacked from several source row ranges (GRU gates)
  PackSeg rseg[3];      // dst rows [dst_k0, +len) <- src rows [src_c0, +len)   (fields reused)
  int split2;           // 1: CTA-pair layout -- two half images (rows [0, Np/2) and [Np/2, Np)) one after the other,
                        //    each a KM8 image of Np/2 rows: a K range of one half is one contiguous byte range
};
constexpr int kMaxPackJobs = 64;
struct PackTable { int njobs; PackJob job[kMaxPackJobs]; };

template <int FMT>
__global__ void pack_weights_kernel(const __grid_constant__ PackTable tab, uint16_t* __restrict__ dst) {
  const PackJob& j = tab.job[blockIdx.y];
  if ((int)blockIdx.y >= tab.njobs) return;
  const long long total = (long long)j.Np * j.Kp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    // iterate in DESTINATION order (coalesced writes): decode KM8 offset -> (n, k)
    const int rows_img = j.split2 ? j.Np / 2 : j.Np;         // rows per KM8 image
    const long long img_elems = (long long)rows_img * j.Kp;
    const int half = j.split2 ? (int)(i / img_elems) : 0;
    const long long ii = i - (long long)half * img_elems;
    const long long per_kgroup = (long long)rows_img * 8;    // elements per 8-col group
    int kg = (int)(ii / per_kgroup);
    int rem = (int)(ii - (long long)kg * per_kgroup);
    int n = half * rows_img + (rem >> 3), k = kg * 8 + (rem & 7);
    float v = 0.f;
    int srow = -1;                       // source row of image row n
    if (j.nrseg > 0) {
      for (int s = 0; s < j.nrseg; ++s) {
        int c = n - j.rseg[s].dst_k0;
        if (c >= 0 && c < j.rseg[s].len) srow = j.rseg[s].src_c0 + c;
      }
    } else if (n < j.N) {
      srow = j.row0 + n;
    }
    if (srow >= 0) {
      if (k == j.bias_k && j.bias) v = j.bias[srow];
      for (int s = 0; s < j.nseg; ++s) {
        int c = k - j.seg[s].dst_k0;
        if (c >= 0 && c < j.seg[s].len) {
          int sc = j.seg[s].src_c0 + c;
          v = j.transpose ? j.w[(long long)sc * j.ld + srow]
                          : j.w[(long long)srow * j.ld + sc];
        }
      }
    }
    dst[j.dst_off + i] = Half16<FMT>::cvt(v);
  }
}

}  // namespace tc
}  // namespace bd

/* bd_b200_test.h -- debug / self-test entry points of the tensor-core building blocks.
 *
 * NOT part of the product library: these live in libbd_b200_test.so (built from
 * csrc/tc_selftest.cu and csrc/tc_mmabench.cu), which only tests/ and scripts/ load.
 */
#ifndef BD_B200_TEST_H
#define BD_B200_TEST_H
#include "bd_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- self-test ---- */
/* One 128-row tile of y (128,N) = x (128,K) w(N,K)^T + b through the tcgen05 building block
 * (16-bit operands: fmt 0 = fp16, 1 = bf16; fp32 accumulation).  ws: >= 2*roundup16(N)*roundup16(K+1)
 * bytes.  Used by the GPU tests to validate the UMMA descriptors / TMEM layout in isolation. */
int bd_tc_selftest(const float* x, const float* w, const float* b, int K, int N, int fmt,
                   int swap_lbo_sbo, void* ws, size_t ws_bytes, float* y, bd_stream_t stream);

/* Debug micro-benchmarks (not on the product path; scripts/mmabench*.py, scripts/dsmembench.py):
 * cycles per tcgen05.mma for different operand layouts / issue structures / concurrent traffic,
 * and the cost of exchanging an operand-tile slice between the CTAs of a cluster. */
int bd_tc_mmabench(int N, int nmma, int layout, int dep, long long* out_cycles, bd_stream_t stream);
int bd_tc_mmabench2(int N, int ksteps, int nrep, int tma, int mode, const void* gsrc, long long* out,
                    bd_stream_t stream);
int bd_tc_dsmembench(int R, int bytes, int mode, int iters, long long* out_cycles, bd_stream_t stream);

/* last error message of THIS library (same contract as bd_last_error) */
const char* bd_test_last_error(void);

#ifdef __cplusplus
}
#endif
#endif

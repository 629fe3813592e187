// Shared helpers for libbd_b200: error reporting, activation math, workspace carving.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/bd_b200.h"

namespace bd {

void set_error(const char* fmt, ...);
extern std::atomic<unsigned long long> g_launch_count;   // kernels launched by this library (bd_launch_count)

#define BD_FAIL(code, ...)        \
  do {                            \
    bd::set_error(__VA_ARGS__);   \
    return (code);                \
  } while (0)

#define BD_CHECK_ARG(cond, ...)                     \
  do {                                              \
    if (!(cond)) BD_FAIL(BD_ERR_BAD_ARG, __VA_ARGS__); \
  } while (0)

#define BD_CUDA_LAUNCH_CHECK()                                                        \
  do {                                                                                \
    ++bd::g_launch_count;                                                             \
    cudaError_t e__ = cudaGetLastError();                                             \
    if (e__ != cudaSuccess)                                                           \
      BD_FAIL(BD_ERR_CUDA, "%s:%d CUDA launch failed: %s", __FILE__, __LINE__,        \
              cudaGetErrorString(e__));                                               \
  } while (0)

#define BD_TRY(expr)               \
  do {                             \
    int rc__ = (expr);             \
    if (rc__ != BD_OK) return rc__; \
  } while (0)

// Dynamic shared-memory limit of a kernel that needs more than 48 KB.  The attribute is per-function
// global state: setting it to the size of every launch made it depend on the LAST launch, which
// breaks any tool or graph path that re-issues an earlier, larger launch of the same kernel (seen
// with ncu's per-node profiling of a replayed CUDA graph: LaunchFailed on the heads'
// mlp_bwd_kernel after the smaller actor launch).  It only ever grows here.
constexpr int kMaxOptinSmem = 227 * 1024 - 4096;     // device opt-in maximum less static barriers
void grow_smem_attr(const void* kernel, int bytes);  // api.cu
template <typename K>
inline void set_smem_attr(K kernel, size_t bytes) { grow_smem_attr(reinterpret_cast<const void*>(kernel), (int)bytes); }

// ---------------------------------------------------------------- optional per-kernel timing
bool prof_enabled();
void prof_begin(int kernel, cudaStream_t s);   // records a start event (no-op when disabled)
void prof_end(int kernel, cudaStream_t s);     // records the matching stop event
struct ProfScope {
  int k; cudaStream_t s;
  ProfScope(int kernel, cudaStream_t stream) : k(kernel), s(stream) { prof_begin(k, s); }
  ~ProfScope() { prof_end(k, s); }
};

// Packed fp32 FMA (sm_100 FFMA2): {d0, d1} += a * {b0, b1}, each lane rounded exactly like fmaf.
// Scalar FFMA issues every other cycle per scheduler (64 FMA/clk/SM); the packed form is the
// only way to the SM's 128 FMA/clk, and takes the row operand as a broadcast scalar.
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a, float b0, float b1) {
  uint64_t d, av, bv;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(d0), "f"(d1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
  asm("mov.b64 %0, {%1, %2};" : "=l"(bv) : "f"(b0), "f"(b1));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(av), "l"(bv));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}

__device__ __forceinline__ uint32_t f32_to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// ---------------------------------------------------------------- activations
// torch semantics: ELU(alpha=1) uses expm1; softplus(beta=1, threshold=20).
__device__ __forceinline__ float act_fwd(int act, float x) {
  switch (act) {
    case BD_ACT_ELU: return x > 0.f ? x : expm1f(x);
    case BD_ACT_RELU: return x > 0.f ? x : 0.f;
    case BD_ACT_TANH: return tanhf(x);
    case BD_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    default: return x;
  }
}
// derivative expressed through the activation OUTPUT y (what backward has at hand)
__device__ __forceinline__ float act_bwd_from_out(int act, float y) {
  switch (act) {
    case BD_ACT_ELU: return y > 0.f ? 1.f : y + 1.f;
    case BD_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case BD_ACT_TANH: return 1.f - y * y;
    case BD_ACT_SIGMOID: return y * (1.f - y);
    default: return 1.f;
  }
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float softplusf_(float x) { return x > 20.f ? x : log1pf(expf(x)); }
// d softplus / dx with the same threshold rule as torch's softplus_backward
__device__ __forceinline__ float softplus_gradf_(float x) { return x > 20.f ? 1.f : sigmoidf_(x); }

// ---------------------------------------------------------------- workspace
struct Arena {
  char* base;
  size_t cap, off;
  Arena(void* p, size_t bytes) : base(static_cast<char*>(p)), cap(bytes), off(0) {}
  float* f32(size_t n) {
    size_t bytes = (n * sizeof(float) + 255) & ~size_t(255);
    if (off + bytes > cap) { off = cap + 1; return nullptr; }
    float* r = reinterpret_cast<float*>(base + off);
    off += bytes;
    return r;
  }
  bool ok() const { return off <= cap; }
};
inline size_t pad256(size_t n_floats) { return (n_floats * sizeof(float) + 255) & ~size_t(255); }

inline bool valid_act(int a) { return a >= BD_ACT_IDENTITY && a <= BD_ACT_SIGMOID; }

}  // namespace bd

// Micro-benchmark (debug): issue rate of legacy mma.sync on sm_100a, TF32 m16n8k8 and FP16 m16n8k16,
// against packed FFMA2, one CTA of 8 warps per SM.  Prints cycles per instruction per scheduler.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k_tf32(float* out, long long* cyc, int iters) {
  float c[8][4] = {};
  uint32_t a[4] = {0x3f800000u + threadIdx.x, 0x3f900000u, 0x3fa00000u, 0x3fb00000u}, b[2] = {0x3f800000u, 0x3f000000u + threadIdx.x};
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  long long t1 = clock64();
  float s = 0; for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_f16(float* out, long long* cyc, int iters) {
  float c[8][4] = {};
  uint32_t a[4] = {0x3c003c00u + threadIdx.x, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u}, b[2] = {0x3c003c00u, 0x38003800u + threadIdx.x};
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  long long t1 = clock64();
  float s = 0; for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_ffma2(float* out, long long* cyc, int iters) {
  unsigned long long c[16];
  for (int j = 0; j < 16; ++j) c[j] = 0;
  unsigned long long a = 0x3f8000003f800000ull + threadIdx.x, b = 0x3f0000003f000000ull;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c[j]) : "l"(a), "l"(b));
  }
  long long t1 = clock64();
  unsigned long long s = 0; for (int j = 0; j < 16; ++j) s += c[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 8);
  const int iters = 20000; long long h;
  for (int rep = 0; rep < 2; ++rep) {
    k_tf32<<<16, 256>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    if (rep) printf("tf32 m16n8k8 : %.2f cycles per mma per scheduler (2 warps each) -> %.0f MAC/clk/SM\n", (double)h / (iters * 8.0 * 2), 1024.0 * 4 / ((double)h / (iters * 8.0 * 2)));
    k_f16<<<16, 256>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    if (rep) printf("f16 m16n8k16 : %.2f cycles per mma per scheduler -> %.0f MAC/clk/SM\n", (double)h / (iters * 8.0 * 2), 2048.0 * 4 / ((double)h / (iters * 8.0 * 2)));
    k_ffma2<<<16, 256>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    if (rep) printf("ffma2        : %.2f cycles per instr per scheduler -> %.0f MAC/clk/SM\n", (double)h / (iters * 16.0 * 2), 64.0 * 4 / ((double)h / (iters * 16.0 * 2)));
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}

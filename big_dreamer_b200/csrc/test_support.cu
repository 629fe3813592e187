// Support code of libbd_b200_test.so (self-test and micro-benchmark entry points, include/bd_b200_test.h):
// the error channel the BD_FAIL / BD_CHECK_ARG macros write to.  The product library has its own (api.cu).
#include "common.cuh"
#include <cstdarg>
#include <cstdio>

namespace bd {
std::atomic<unsigned long long> g_launch_count{0};   // BD_CUDA_LAUNCH_CHECK counts launches per library
static thread_local char g_test_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_test_err, sizeof(g_test_err), fmt, ap);
  va_end(ap);
}
}  // namespace bd

extern "C" const char* bd_test_last_error(void) { return bd::g_test_err; }

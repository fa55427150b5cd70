#include <stdlib.h>
// extern "C" boundary: error reporting + the generic GEMM entry point.
#include "common.cuh"
#include "host.h"
#include "../../include/magb200.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>

namespace magb200 {
static thread_local char g_last_error[1024] = "";
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}
static unsigned long long g_launches = 0;
int pdl_level() {
  static int level = -1;
  if (level < 0) {
    const char* e = getenv("MAGB200_PDL");
    level = (e && *e) ? atoi(e) : 1;
  }
  return level;
}
void count_launch() { __atomic_add_fetch(&g_launches, 1ull, __ATOMIC_RELAXED); }
}  // namespace magb200

using namespace magb200;

extern "C" {

const char* magb200_last_error(void) { return g_last_error; }

int magb200_version(void) { return MAGB200_VERSION; }

unsigned long long magb200_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

int magb200_device_check(void) {
  int dev = 0;
  MAGB_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  MAGB_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
  MAGB_REQUIRE(prop.major == 10, "libmagb200 needs an sm_100 (B200) device, found sm_%d%d (%s)", prop.major, prop.minor,
               prop.name);
  return 0;
}

void magb200_gemm_set_debug(void* trace_dev, int cfg) {
  gemm_set_debug(reinterpret_cast<unsigned long long*>(trace_dev), cfg);
}

int magb200_gemm(const void* A, const void* B, int M, int N, int K, int lda, int ldb, int a_dtype, int b_dtype,
                 int a_mn_major, int b_mn_major, int epilogue, const float* bias, void* out0, void* out1,
                 const void* aux, int ldo, float alpha, int splits, int n_split, int k_start_hi, void* stream) {
  GemmProblem p;
  p.A = A; p.B = B; p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldb = ldb;
  p.a_dtype = a_dtype; p.b_dtype = b_dtype; p.a_mn = a_mn_major != 0; p.b_mn = b_mn_major != 0;
  p.epi = epilogue; p.bias = bias; p.out0 = out0; p.out1 = out1; p.aux = aux; p.ldo = ldo; p.alpha = alpha;
  p.splits = splits; p.n_split = n_split; p.k_start_hi = k_start_hi;
  return gemm_launch(p, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"

// Micro-benchmark (B200): how fast can one CTA per SM push a 64 KB fp32 tile from shared memory to global?
//   mode 0: 4 warps, coalesced st.global.v4 (512 B contiguous per instruction)   [the GEMM epilogue today]
//   mode 1: 8 warps, same
//   mode 2: 4 warps, cp.async.bulk.global.shared::cta, one 512 B row per lane-issued instruction
//   mode 3: one elected thread issues all 128 row copies
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_bw store_bw.cu && ./store_bw
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int ld, long long* cyc) {
  extern __shared__ __align__(128) uint8_t sm[];
  const int rows = 128, rowbytes = 512;
  for (int i = threadIdx.x; i < rows * rowbytes / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = (float)i;
  __syncthreads();
  float* base = out + (size_t)blockIdx.x * rows * ld;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long t0 = clock64();
  if (MODE == 0 || MODE == 1) {
    const int nw = MODE == 0 ? 4 : 8;
    if (warp < nw) {
      for (int r = warp; r < rows; r += nw) {
        const uint4 v = *reinterpret_cast<const uint4*>(sm + r * rowbytes + lane * 16);
        asm volatile("st.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(base + (size_t)r * ld + lane * 4), "r"(v.x), "r"(v.y),
                     "r"(v.z), "r"(v.w) : "memory");
      }
    }
  } else if (MODE == 2) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp < 4) {
      const int r = warp * 32 + lane;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + (size_t)r * ld),
                   "r"(smem_u32(sm + r * rowbytes)), "r"(rowbytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  } else {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
      for (int r = 0; r < rows; ++r)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + (size_t)r * ld),
                     "r"(smem_u32(sm + r * rowbytes)), "r"(rowbytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, float* out, int ld, long long* dcyc, int nblk) {
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int it = 0; it < 3; ++it) k<MODE><<<nblk, 256, 65536>>>(out, ld, dcyc);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int it = 0; it < 20; ++it) k<MODE><<<nblk, 256, 65536>>>(out, ld, dcyc);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[148]; cudaMemcpy(h, dcyc, sizeof(long long) * nblk, cudaMemcpyDeviceToHost);
  long long mx = 0, sum = 0; for (int i = 0; i < nblk; ++i) { sum += h[i]; if (h[i] > mx) mx = h[i]; }
  printf("%-46s blocks=%3d  store phase cycles avg=%6lld max=%6lld  -> %.1f B/clk/SM   kernel %.2f us  err=%s\n", name, nblk,
         sum / nblk, mx, 65536.0 / (sum / (double)nblk), ms * 1000 / 20, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const int ld = 768, rows = 128;
  float* out; long long* dcyc;
  cudaMalloc(&out, (size_t)148 * rows * ld * 4); cudaMalloc(&dcyc, 148 * 8);
  for (int nblk : {114, 148, 16}) {
    run<0>("st.global.v4, 4 warps", out, ld, dcyc, nblk);
    run<1>("st.global.v4, 8 warps", out, ld, dcyc, nblk);
    run<2>("cp.async.bulk rows, 128 lanes issue", out, ld, dcyc, nblk);
    run<3>("cp.async.bulk rows, 1 thread issues", out, ld, dcyc, nblk);
  }
  return 0;
}

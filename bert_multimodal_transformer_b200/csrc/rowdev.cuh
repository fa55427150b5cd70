// Device helpers of the row-wise (per-token) kernels: one warp owns one token row of hidden size 768 = 32 lanes x 24
// floats; every global access is a float4 / fp16x4 vector, coalesced across the warp.  Shared by rowops.cu and
// mag_fused.cu.
#pragma once
#include "common.cuh"

namespace magb200 {

constexpr int H = 768;
constexpr int VEC = 6;  // float4 per lane per row

struct Row {
  float v[4 * VEC];
};

__device__ __forceinline__ int col_of(int lane, int j) { return 4 * (lane + 32 * j); }

__device__ __forceinline__ void load_row(Row& r, const float* __restrict__ p, int lane) {
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const float4 t = *reinterpret_cast<const float4*>(p + col_of(lane, j));
    r.v[4 * j] = t.x; r.v[4 * j + 1] = t.y; r.v[4 * j + 2] = t.z; r.v[4 * j + 3] = t.w;
  }
}
__device__ __forceinline__ void store_row(const Row& r, float* __restrict__ p, int lane) {
#pragma unroll
  for (int j = 0; j < VEC; ++j)
    *reinterpret_cast<float4*>(p + col_of(lane, j)) = make_float4(r.v[4 * j], r.v[4 * j + 1], r.v[4 * j + 2], r.v[4 * j + 3]);
}
__device__ __forceinline__ void store_row_f16(const Row& r, __half* __restrict__ p, int lane) {
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    __half2 a = __floats2half2_rn(r.v[4 * j], r.v[4 * j + 1]);
    __half2 b = __floats2half2_rn(r.v[4 * j + 2], r.v[4 * j + 3]);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p + col_of(lane, j)) = u;
  }
}
__device__ __forceinline__ void store_row_bf16(const Row& r, __nv_bfloat16* __restrict__ p, int lane) {
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    __nv_bfloat162 a = __floats2bfloat162_rn(r.v[4 * j], r.v[4 * j + 1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(r.v[4 * j + 2], r.v[4 * j + 3]);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p + col_of(lane, j)) = u;
  }
}
__device__ __forceinline__ float row_sum(const Row& r) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4 * VEC; ++i) s += r.v[i];
  return warp_sum(s);
}
__device__ __forceinline__ float row_dot(const Row& a, const Row& b) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4 * VEC; ++i) s += a.v[i] * b.v[i];
  return warp_sum(s);
}
// two-pass LayerNorm statistics (matches torch.nn.functional.layer_norm on fp32)
__device__ __forceinline__ void row_stats(const Row& r, float eps, float& mean, float& rstd) {
  mean = row_sum(r) * (1.0f / H);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4 * VEC; ++i) {
    const float d = r.v[i] - mean;
    s += d * d;
  }
  const float var = warp_sum(s) * (1.0f / H);
  rstd = 1.0f / sqrtf(var + eps);
}
// in place: r <- keep ? r * scale : 0 for the row's elements (element index = row*H + col)
__device__ __forceinline__ void row_dropout(Row& r, uint64_t seed, uint32_t site, int64_t row, int lane, uint32_t thresh,
                                            float scale) {
  if (thresh == 0) return;
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const uint64_t e4 = ((uint64_t)row * H + col_of(lane, j)) >> 2;
    const uint4 u = philox4x32(seed, site, e4);
    r.v[4 * j] = u.x >= thresh ? r.v[4 * j] * scale : 0.f;
    r.v[4 * j + 1] = u.y >= thresh ? r.v[4 * j + 1] * scale : 0.f;
    r.v[4 * j + 2] = u.z >= thresh ? r.v[4 * j + 2] * scale : 0.f;
    r.v[4 * j + 3] = u.w >= thresh ? r.v[4 * j + 3] * scale : 0.f;
  }
}

// Column reductions (parameter gradients): each warp adds its row's contribution into a
// block-shared accumulator with shared-memory atomics (columns differ across lanes: conflict-free),
// and the block flushes one global atomicAdd per column at the end.
__device__ __forceinline__ void acc_zero(float* acc, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
}
__device__ __forceinline__ void acc_add_row(float* acc, const Row& r, int lane) {
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const int c = col_of(lane, j);
#pragma unroll
    for (int k = 0; k < 4; ++k) atomicAdd(acc + c + k, r.v[4 * j + k]);
  }
}
__device__ __forceinline__ void acc_flush(const float* acc, float* __restrict__ gout, int n, float scale) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(gout + i, acc[i] * scale);
}

__device__ __forceinline__ void red_add_f32x4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Register-accumulator variant for the per-layer kernels (shared-memory atomics cost ~64 cycles per
// warp instruction, measured 26 us for ln_bwd): each warp keeps its column partial sums in registers
// over all the rows it visits; at the end the NW warps of the block combine through shared memory and
// the block issues one global atomicAdd per column.
__device__ __forceinline__ void row_zero(Row& r) {
#pragma unroll
  for (int i = 0; i < 4 * VEC; ++i) r.v[i] = 0.f;
}
template <int NW>
__device__ __forceinline__ void block_reduce_flush(const Row& part, float* smem /*[NW][H]*/, float* __restrict__ gout,
                                                   float scale, int warp, int lane) {
  __syncthreads();
  store_row(part, smem + warp * H, lane);
  __syncthreads();
  for (int c = threadIdx.x; c < H; c += NW * 32) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) s += smem[w * H + c];
    atomicAdd(gout + c, s * scale);
  }
}

}  // namespace magb200

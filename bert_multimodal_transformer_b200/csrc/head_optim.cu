// Classification head (bert.py:304-307, 313-322; xlnet.py:506-509), MSE loss
// (multimodal_driver.py:371-373) and the fused AdamW step (transformers 3.0.2 AdamW as configured at
// multimodal_driver.py:328-350) over the flat parameter buffer.
#include "common.cuh"
#include "host.h"

namespace magb200 {

constexpr int H = 768;

// pooled = tanh(pre)  [saved, fp32] ; logits = dropout(pooled) . Wc^T + bc        one warp per batch row
__global__ void __launch_bounds__(128) cls_fwd_kernel(const float* __restrict__ pre, const float* __restrict__ Wc,
                                                      const float* __restrict__ bc, int B, int L, uint32_t thresh,
                                                      float scale, Seed seed_in, uint32_t site,
                                                      float* __restrict__ pooled, float* __restrict__ logits) {
  const uint64_t seed = seed_in.ptr ? *seed_in.ptr : seed_in.value;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + warp;
  if (b >= B) return;
  float v[24];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const int c = 4 * (lane + 32 * j);
    const float4 t = *reinterpret_cast<const float4*>(pre + (int64_t)b * H + c);
    float4 o = make_float4(tanhf(t.x), tanhf(t.y), tanhf(t.z), tanhf(t.w));
    *reinterpret_cast<float4*>(pooled + (int64_t)b * H + c) = o;
    if (thresh) {
      const uint4 u = philox4x32(seed, site, ((uint64_t)b * H + c) >> 2);
      o.x = u.x >= thresh ? o.x * scale : 0.f;
      o.y = u.y >= thresh ? o.y * scale : 0.f;
      o.z = u.z >= thresh ? o.z * scale : 0.f;
      o.w = u.w >= thresh ? o.w * scale : 0.f;
    }
    v[4 * j] = o.x; v[4 * j + 1] = o.y; v[4 * j + 2] = o.z; v[4 * j + 3] = o.w;
  }
  for (int l = 0; l < L; ++l) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const float4 w = *reinterpret_cast<const float4*>(Wc + (int64_t)l * H + 4 * (lane + 32 * j));
      s += v[4 * j] * w.x + v[4 * j + 1] * w.y + v[4 * j + 2] * w.z + v[4 * j + 3] * w.w;
    }
    s = warp_sum(s);
    if (lane == 0) logits[(int64_t)b * L + l] = s + bc[l];
  }
}

// dlogits [B][L] -> dpre (fp16 [B][768], gradient wrt the pooler pre-activation), dWc, dbc (+=)
__global__ void __launch_bounds__(128) cls_bwd_kernel(const float* __restrict__ dlogits, const float* __restrict__ pooled,
                                                      const float* __restrict__ Wc, int B, int L, uint32_t thresh,
                                                      float scale, Seed seed_in, uint32_t site, float oscale,
                                                      __half* __restrict__ dpre, float* __restrict__ dWc,
                                                      float* __restrict__ dbc) {
  pdl_trigger();   // the next kernel in the stream (a GEMM) may start its private prologue
  const uint64_t seed = seed_in.ptr ? *seed_in.ptr : seed_in.value;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + warp;
  if (b >= B) return;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const int c = 4 * (lane + 32 * j);
    const float4 pl = *reinterpret_cast<const float4*>(pooled + (int64_t)b * H + c);
    float keep[4] = {1.f, 1.f, 1.f, 1.f};
    if (thresh) {
      const uint4 u = philox4x32(seed, site, ((uint64_t)b * H + c) >> 2);
      keep[0] = u.x >= thresh ? scale : 0.f;
      keep[1] = u.y >= thresh ? scale : 0.f;
      keep[2] = u.z >= thresh ? scale : 0.f;
      keep[3] = u.w >= thresh ? scale : 0.f;
    }
    const float p[4] = {pl.x, pl.y, pl.z, pl.w};
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    for (int l = 0; l < L; ++l) {
      const float g = dlogits[(int64_t)b * L + l];
      const float4 w = *reinterpret_cast<const float4*>(Wc + (int64_t)l * H + c);
      d[0] += g * w.x; d[1] += g * w.y; d[2] += g * w.z; d[3] += g * w.w;
#pragma unroll
      for (int k = 0; k < 4; ++k) atomicAdd(dWc + (int64_t)l * H + c + k, g * p[k] * keep[k] * oscale);
    }
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = d[k] * keep[k] * (1.f - p[k] * p[k]);
    __half2 a = __floats2half2_rn(o[0], o[1]), bb = __floats2half2_rn(o[2], o[3]);
    uint2 u2;
    u2.x = *reinterpret_cast<uint32_t*>(&a);
    u2.y = *reinterpret_cast<uint32_t*>(&bb);
    *reinterpret_cast<uint2*>(dpre + (int64_t)b * H + c) = u2;
  }
  if (lane < L) atomicAdd(dbc + lane, dlogits[(int64_t)b * L + lane] * oscale);
}

// loss = mean((logits - labels)^2); dlogits = 2 (logits - labels) / n * gscale      single block
__global__ void __launch_bounds__(256) mse_kernel(const float* __restrict__ logits, const float* __restrict__ labels,
                                                  int n, float gscale, float* __restrict__ loss,
                                                  float* __restrict__ dlogits) {
  __shared__ float red[8];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = logits[i] - labels[i];
    s += d * d;
    if (dlogits) dlogits[i] = 2.f * d / (float)n * gscale;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    *loss = t / (float)n;
  }
}

// Fused AdamW (transformers 3.0.2 `AdamW.step`, correct_bias=True):
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= step_size * m / (sqrt(v) + eps)
//   step_size = lr * sqrt(1-b2^t) / (1-b1^t) ;  then  p -= lr * wd * p  where the 256-element chunk is flagged
// One pass over the flat buffers: 16 B read + 12 B written per parameter (+2 B fp16 shadow), HBM bound.
// Also clears the gradient (zero_grad) and refreshes the fp16 operand copy the GEMMs read.
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, __half* __restrict__ p16,
                                                    const uint8_t* __restrict__ decay_chunk, int64_t n4, float lr,
                                                    float b1, float b2, float eps, float wd, float step_size,
                                                    float gscale, int zero_grad,
                                                    const float* __restrict__ hyper) {
  if (hyper != nullptr) {
    lr = hyper[0];
    step_size = hyper[1];
  }
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 gg = reinterpret_cast<float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    const float dec = decay_chunk[i >> 6] ? lr * wd : 0.f;   // 64 float4 = 256 elements per chunk
    float* P = &pp.x; float* G = &gg.x; float* M = &mm.x; float* V = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gr = G[k] * gscale;
      M[k] = b1 * M[k] + (1.f - b1) * gr;
      V[k] = b2 * V[k] + (1.f - b2) * gr * gr;
      P[k] -= step_size * M[k] / (sqrtf(V[k]) + eps);
      P[k] -= dec * P[k];
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p16) {
      __half2 a = __floats2half2_rn(pp.x, pp.y), b = __floats2half2_rn(pp.z, pp.w);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&a);
      u.y = *reinterpret_cast<uint32_t*>(&b);
      reinterpret_cast<uint2*>(p16)[i] = u;
    }
  }
}

static inline void drop_params(float p, uint32_t& thresh, float& scale) {
  if (p <= 0.f) { thresh = 0; scale = 1.f; return; }
  double t = (double)p * 4294967296.0;
  thresh = t >= 4294967295.0 ? 4294967295u : (uint32_t)t;
  scale = 1.0f / (1.0f - p);
}

int cls_fwd(const float* pre, const float* Wc, const float* bc, int B, int L, float p, Seed seed, uint32_t site,
            float* pooled, float* logits, cudaStream_t st) {
  MAGB_REQUIRE(L >= 1 && L <= 32, "classifier: num_labels=%d out of range [1,32]", L);
  uint32_t th; float sc;
  drop_params(p, th, sc);
  cls_fwd_kernel<<<(B + 3) / 4, 128, 0, st>>>(pre, Wc, bc, B, L, th, sc, seed, site, pooled, logits);
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int cls_bwd(const float* dlogits, const float* pooled, const float* Wc, int B, int L, float p, Seed seed,
            uint32_t site, float oscale, __half* dpre, float* dWc, float* dbc, cudaStream_t st) {
  MAGB_REQUIRE(L >= 1 && L <= 32, "classifier: num_labels=%d out of range [1,32]", L);
  uint32_t th; float sc;
  drop_params(p, th, sc);
  cls_bwd_kernel<<<(B + 3) / 4, 128, 0, st>>>(dlogits, pooled, Wc, B, L, th, sc, seed, site, oscale, dpre, dWc, dbc);
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int mse_loss(const float* logits, const float* labels, int n, float gscale, float* loss, float* dlogits,
             cudaStream_t st) {
  mse_kernel<<<1, 256, 0, st>>>(logits, labels, n, gscale, loss, dlogits);
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int adamw_step(float* p, float* g, float* m, float* v, __half* p16, const uint8_t* decay_chunk, int64_t n, float lr,
               float b1, float b2, float eps, float wd, int step, float gscale, int zero_grad, const float* hyper_dev,
               cudaStream_t st) {
  MAGB_REQUIRE(n % 256 == 0, "adamw: flat buffer length must be a multiple of 256");
  MAGB_REQUIRE(step >= 1, "adamw: step must be >= 1");
  const double bc1 = 1.0 - pow((double)b1, step), bc2 = 1.0 - pow((double)b2, step);
  const float step_size = (float)((double)lr * sqrt(bc2) / bc1);
  const int64_t n4 = n / 4;
  int blocks = (int)((n4 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  adamw_kernel<<<blocks, 256, 0, st>>>(p, g, m, v, p16, decay_chunk, n4, lr, b1, b2, eps, wd, step_size, gscale,
                                       zero_grad, hyper_dev);
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace magb200

// Scaled-dot-product attention core of BertSelfAttention (transformers 3.0.2, called through
// bert.py:221-229):
//     P = dropout(softmax(Q K^T / sqrt(d) + (1 - mask) * -10000));  ctx = P V
// and its backward.  One CTA per (head, batch element); Q/K/V rows of that head are staged in
// shared memory as fp16 with a 66-element pitch (bank-conflict-free for per-lane key rows),
// scores/probabilities live in registers + a per-warp shared row.  At S = 50 this is 1.1 % of the
// layer FLOPs: the kernel is latency/occupancy bound, not tensor bound.
#include "common.cuh"
#include "host.h"

namespace magb200 {

constexpr int DH = 64;        // head dim
constexpr int KP = DH + 2;    // smem pitch in halfs (33 words: consecutive rows shift one bank)
constexpr int AW = 4;         // warps per CTA

__device__ __forceinline__ float mask_add(const int64_t* mrow, int j) {
  // get_extended_attention_mask (transformers 3.0.2; bert.py:180-182): (1 - m) * -10000
  return mrow == nullptr ? 0.f : (1.0f - (float)mrow[j]) * -10000.0f;
}

// load one head's [S][64] slice of the fused qkv activation into smem (fp16, pitch KP)
__device__ __forceinline__ void load_head(const __half* __restrict__ base, int ld, int S, __half* dst) {
  // 8 halfs (16 B) per thread per step; 8 steps per row
  for (int i = threadIdx.x; i < S * 8; i += blockDim.x) {
    const int r = i >> 3, c = (i & 7) * 8;
    const uint4 v = *reinterpret_cast<const uint4*>(base + (int64_t)r * ld + c);
    uint32_t* d = reinterpret_cast<uint32_t*>(dst + r * KP + c);
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
}
__device__ __forceinline__ void load_head_f32(const __half* __restrict__ base, int ld, int S, float* dst /*[S][DH+1]*/) {
  for (int i = threadIdx.x; i < S * 32; i += blockDim.x) {
    const int r = i >> 5, c = (i & 31) * 2;
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(base + (int64_t)r * ld + c));
    dst[r * (DH + 1) + c] = f.x;
    dst[r * (DH + 1) + c + 1] = f.y;
  }
}

__device__ __forceinline__ float dot64(const __half* a, const __half* b) {
  float s = 0.f;
#pragma unroll
  for (int d = 0; d < DH; d += 2) {
    const float2 x = __half22float2(*reinterpret_cast<const __half2*>(a + d));
    const float2 y = __half22float2(*reinterpret_cast<const __half2*>(b + d));
    s = fmaf(x.x, y.x, s);
    s = fmaf(x.y, y.y, s);
  }
  return s;
}

// element index of P[b][h][i][j] for the dropout counter
__device__ __forceinline__ bool keep_prob(uint64_t seed, uint32_t site, uint64_t idx, uint32_t thresh) {
  if (thresh == 0) return true;
  const uint4 u = philox4x32(seed, site, idx >> 2);
  const uint32_t w = (idx & 3) == 0 ? u.x : (idx & 3) == 1 ? u.y : (idx & 3) == 2 ? u.z : u.w;
  return w >= thresh;
}

// ------------------------------------------------------------------------------------ forward
// qkv: [B*S][3*heads*64] fp16 (q | k | v), ctx: [B*S][heads*64] fp16
template <int MAXJ>  // keys per lane = ceil(S / 32) <= MAXJ
__global__ void __launch_bounds__(AW * 32) sdpa_fwd_kernel(const __half* __restrict__ qkv,
                                                           const int64_t* __restrict__ mask, int S, int heads,
                                                           float scale, uint32_t thresh, float dscale, Seed seed_in,
                                                           uint32_t site, __half* __restrict__ ctx,
                                                           float* __restrict__ probs_out) {
  const uint64_t seed = seed_in.ptr ? *seed_in.ptr : seed_in.value;
  extern __shared__ uint8_t smem_att[];
  __half* sK = reinterpret_cast<__half*>(smem_att);
  __half* sV = sK + S * KP;
  __half* sQ = sV + S * KP;
  float* sP = reinterpret_cast<float*>(sQ + S * KP);  // [AW][S]
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Hd = heads * DH, ld = 3 * Hd;
  const __half* base = qkv + (int64_t)b * S * ld + h * DH;
  load_head(base, ld, S, sQ);
  load_head(base + Hd, ld, S, sK);
  load_head(base + 2 * Hd, ld, S, sV);
  __syncthreads();
  const int64_t* mrow = mask ? mask + (int64_t)b * S : nullptr;
  float* prow = sP + warp * S;
  for (int i = warp; i < S; i += AW) {
    float sc[MAXJ];
    float mx = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < MAXJ; ++jj) {
      const int j = lane + 32 * jj;
      sc[jj] = -INFINITY;
      if (j < S) {
        sc[jj] = dot64(sQ + i * KP, sK + j * KP) * scale + mask_add(mrow, j);
        mx = fmaxf(mx, sc[jj]);
      }
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < MAXJ; ++jj) {
      const int j = lane + 32 * jj;
      sc[jj] = j < S ? expf(sc[jj] - mx) : 0.f;
      sum += sc[jj];
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    __syncwarp();
#pragma unroll
    for (int jj = 0; jj < MAXJ; ++jj) {
      const int j = lane + 32 * jj;
      if (j < S) {
        float p = sc[jj] * inv;
        const uint64_t idx = (((uint64_t)b * heads + h) * S + i) * S + j;
        p = keep_prob(seed, site, idx, thresh) ? p * dscale : 0.f;
        prow[j] = p;
        if (probs_out != nullptr) probs_out[idx] = p;          // output_attentions: the (dropped) probabilities
      }
    }
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < S; ++j) {
      const float p = prow[j];
      const float2 v = __half22float2(*reinterpret_cast<const __half2*>(sV + j * KP + 2 * lane));
      o0 = fmaf(p, v.x, o0);
      o1 = fmaf(p, v.y, o1);
    }
    if (ctx != nullptr)
      *reinterpret_cast<__half2*>(ctx + ((int64_t)b * S + i) * Hd + h * DH + 2 * lane) = __floats2half2_rn(o0, o1);
  }
}

// ------------------------------------------------------------------------------------ backward
// dctx: [B*S][heads*64] fp16 -> dqkv: [B*S][3*heads*64] fp16 (loss-scaled gradients).  Probabilities are recomputed from
// Q, K (and the dropout counter), so the forward stores nothing but qkv.
template <int MAXJ>
__global__ void __launch_bounds__(AW * 32) sdpa_bwd_kernel(const __half* __restrict__ qkv,
                                                           const __half* __restrict__ dctx,
                                                           const int64_t* __restrict__ mask, int S, int heads,
                                                           float scale, uint32_t thresh, float dscale, Seed seed_in,
                                                           uint32_t site, __half* __restrict__ dqkv) {
  const uint64_t seed = seed_in.ptr ? *seed_in.ptr : seed_in.value;
  extern __shared__ uint8_t smem_att[];
  __half* sK = reinterpret_cast<__half*>(smem_att);
  __half* sV = sK + S * KP;
  __half* sQ = sV + S * KP;
  float* sdO = reinterpret_cast<float*>(sQ + S * KP);  // [S][DH+1]
  float* sPd = sdO + S * (DH + 1);                      // [S][S+1] dropped probabilities
  float* sdS = sPd + S * (S + 1);                       // [S][S+1] dS
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Hd = heads * DH, ld = 3 * Hd;
  const __half* base = qkv + (int64_t)b * S * ld + h * DH;
  load_head(base, ld, S, sQ);
  load_head(base + Hd, ld, S, sK);
  load_head(base + 2 * Hd, ld, S, sV);
  load_head_f32(dctx + (int64_t)b * S * Hd + h * DH, Hd, S, sdO);
  __syncthreads();
  const int64_t* mrow = mask ? mask + (int64_t)b * S : nullptr;
  const int SP = S + 1;
  // phase 1: per query row i -> Pd[i][:], dS[i][:]
  for (int i = warp; i < S; i += AW) {
    float sc[MAXJ], dp[MAXJ];
    float mx = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < MAXJ; ++jj) {
      const int j = lane + 32 * jj;
      sc[jj] = -INFINITY;
      dp[jj] = 0.f;
      if (j < S) {
        sc[jj] = dot64(sQ + i * KP, sK + j * KP) * scale + mask_add(mrow, j);
        mx = fmaxf(mx, sc[jj]);
        // dPd[i][j] = sum_d dO[i][d] V[j][d]
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < DH; d += 2) {
          const float2 v = __half22float2(*reinterpret_cast<const __half2*>(sV + j * KP + d));
          s = fmaf(sdO[i * (DH + 1) + d], v.x, s);
          s = fmaf(sdO[i * (DH + 1) + d + 1], v.y, s);
        }
        dp[jj] = s;
      }
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < MAXJ; ++jj) {
      const int j = lane + 32 * jj;
      sc[jj] = j < S ? expf(sc[jj] - mx) : 0.f;
      sum += sc[jj];
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    float dot = 0.f;
#pragma unroll
    for (int jj = 0; jj < MAXJ; ++jj) {
      const int j = lane + 32 * jj;
      if (j < S) {
        const float p = sc[jj] * inv;
        const uint64_t idx = (((uint64_t)b * heads + h) * S + i) * S + j;
        const float k = keep_prob(seed, site, idx, thresh) ? dscale : 0.f;
        sPd[i * SP + j] = p * k;
        dp[jj] *= k;          // dP = dPd * keep / (1-p)
        sc[jj] = p;
        dot += dp[jj] * p;
      }
    }
    dot = warp_sum(dot);
#pragma unroll
    for (int jj = 0; jj < MAXJ; ++jj) {
      const int j = lane + 32 * jj;
      if (j < S) sdS[i * SP + j] = sc[jj] * (dp[jj] - dot) * scale;   // includes the 1/sqrt(d) factor
    }
  }
  __syncthreads();
  // phase 2: dQ[i] = sum_j dS[i][j] K[j];  dK[j] = sum_i dS[i][j] Q[i];  dV[j] = sum_i Pd[i][j] dO[i]
  __half* obase = dqkv + (int64_t)b * S * ld + h * DH;
  for (int r = warp; r < S; r += AW) {
    float q0 = 0.f, q1 = 0.f, k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
    for (int x = 0; x < S; ++x) {
      const float ds_rx = sdS[r * SP + x];   // row r as query
      const float ds_xr = sdS[x * SP + r];   // row r as key
      const float pd_xr = sPd[x * SP + r];
      const float2 kk = __half22float2(*reinterpret_cast<const __half2*>(sK + x * KP + 2 * lane));
      const float2 qq = __half22float2(*reinterpret_cast<const __half2*>(sQ + x * KP + 2 * lane));
      q0 = fmaf(ds_rx, kk.x, q0); q1 = fmaf(ds_rx, kk.y, q1);
      k0 = fmaf(ds_xr, qq.x, k0); k1 = fmaf(ds_xr, qq.y, k1);
      v0 = fmaf(pd_xr, sdO[x * (DH + 1) + 2 * lane], v0);
      v1 = fmaf(pd_xr, sdO[x * (DH + 1) + 2 * lane + 1], v1);
    }
    __half* o = obase + (int64_t)r * ld + 2 * lane;
    *reinterpret_cast<__half2*>(o) = __floats2half2_rn(q0, q1);
    *reinterpret_cast<__half2*>(o + Hd) = __floats2half2_rn(k0, k1);
    *reinterpret_cast<__half2*>(o + 2 * Hd) = __floats2half2_rn(v0, v1);
  }
}

static inline void drop_params(float p, uint32_t& thresh, float& scale) {
  if (p <= 0.f) { thresh = 0; scale = 1.f; return; }
  double t = (double)p * 4294967296.0;
  thresh = t >= 4294967295.0 ? 4294967295u : (uint32_t)t;
  scale = 1.0f / (1.0f - p);
}

template <int MAXJ>
static int launch_fwd(const __half* qkv, const int64_t* mask, int B, int S, int heads, float p, Seed seed,
                      uint32_t site, __half* ctx, cudaStream_t st, float* probs_out = nullptr) {
  uint32_t th; float ds;
  drop_params(p, th, ds);
  const size_t smem = (size_t)3 * S * KP * sizeof(__half) + (size_t)AW * S * sizeof(float);
  auto k = sdpa_fwd_kernel<MAXJ>;
  static size_t configured_by_dev[kMaxDevices] = {0};
  size_t& configured = configured_by_dev[current_device_slot()];
  if (smem > configured) {
    MAGB_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  k<<<dim3(heads, B), AW * 32, smem, st>>>(qkv, mask, S, heads, 0.125f, th, ds, seed, site, ctx, probs_out);
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// attention probabilities P[b][h][i][j] (after dropout, as BertSelfAttention returns them with output_attentions) from the
// saved q | k | v activations; not on the training path (the fused kernels never materialise P)
int sdpa_probs(const __half* qkv, const int64_t* mask, int B, int S, int heads, float p, Seed seed, uint32_t site,
               float* probs, cudaStream_t st) {
  MAGB_REQUIRE(S >= 1 && S <= 512 && probs != nullptr, "sdpa_probs: bad arguments (S=%d)", S);
  if (S <= 64) return launch_fwd<2>(qkv, mask, B, S, heads, p, seed, site, nullptr, st, probs);
  if (S <= 128) return launch_fwd<4>(qkv, mask, B, S, heads, p, seed, site, nullptr, st, probs);
  if (S <= 256) return launch_fwd<8>(qkv, mask, B, S, heads, p, seed, site, nullptr, st, probs);
  return launch_fwd<16>(qkv, mask, B, S, heads, p, seed, site, nullptr, st, probs);
}

int sdpa_fwd(const __half* qkv, const int64_t* mask, int B, int S, int heads, int head_dim, float p, Seed seed,
             uint32_t site, __half* ctx, float* lse, cudaStream_t st) {
  MAGB_REQUIRE(head_dim == DH, "sdpa: head_dim must be 64");
  MAGB_REQUIRE(S >= 1 && S <= 512, "sdpa: sequence length %d out of range [1,512]", S);
  // short sequences with an lse buffer: warp-MMA kernel (attention_mma.cu); otherwise the general kernel
  if (S <= 64 && lse != nullptr) return sdpa_fwd_mma(qkv, mask, B, S, heads, p, seed, site, ctx, lse, st);
  // longer sequences with an lse buffer: non-materialising tensor-core kernel (attention_flash.cu)
  if (lse != nullptr) return sdpa_fwd_flash(qkv, mask, B, S, heads, p, seed, site, ctx, lse, st);
  if (S <= 64) return launch_fwd<2>(qkv, mask, B, S, heads, p, seed, site, ctx, st);
  if (S <= 128) return launch_fwd<4>(qkv, mask, B, S, heads, p, seed, site, ctx, st);
  if (S <= 256) return launch_fwd<8>(qkv, mask, B, S, heads, p, seed, site, ctx, st);
  return launch_fwd<16>(qkv, mask, B, S, heads, p, seed, site, ctx, st);
}

template <int MAXJ>
static int launch_bwd(const __half* qkv, const __half* dctx, const int64_t* mask, int B, int S, int heads, float p,
                      Seed seed, uint32_t site, __half* dqkv, cudaStream_t st) {
  uint32_t th; float ds;
  drop_params(p, th, ds);
  const size_t smem = (size_t)3 * S * KP * sizeof(__half) + (size_t)S * (DH + 1) * sizeof(float) +
                      (size_t)2 * S * (S + 1) * sizeof(float);
  auto k = sdpa_bwd_kernel<MAXJ>;
  static size_t configured_by_dev[kMaxDevices] = {0};
  size_t& configured = configured_by_dev[current_device_slot()];
  if (smem > configured) {
    MAGB_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  k<<<dim3(heads, B), AW * 32, smem, st>>>(qkv, dctx, mask, S, heads, 0.125f, th, ds, seed, site, dqkv);
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int sdpa_bwd(const __half* qkv, const __half* dctx, const __half* ctx, const float* lse, const int64_t* mask, int B,
             int S, int heads, int head_dim, float p, Seed seed, uint32_t site, float oscale, __half* dqkv, float* dbias,
             float* scratch, cudaStream_t st) {
  MAGB_REQUIRE(head_dim == DH, "sdpa: head_dim must be 64");
  MAGB_REQUIRE(S >= 1 && S <= 512, "sdpa backward: sequence length %d out of range [1,512]", S);
  if (S <= 64 && lse != nullptr && ctx != nullptr)
    return sdpa_bwd_mma(qkv, dctx, ctx, lse, mask, B, S, heads, p, seed, site, oscale, dqkv, dbias, st);
  if (lse != nullptr && ctx != nullptr && scratch != nullptr) {
    int rc = sdpa_bwd_flash(qkv, dctx, ctx, lse, mask, B, S, heads, p, seed, site, scratch, dqkv, st);
    if (rc) return rc;
    if (dbias != nullptr) return colsum_f16(dqkv, B * S, 3 * heads * DH, oscale, dbias, st);
    return 0;
  }
  MAGB_REQUIRE(S <= 128, "sdpa backward: S=%d > 128 needs lse, ctx and a workspace (the general kernel materialises S x S)", S);
  int rc = S <= 64 ? launch_bwd<2>(qkv, dctx, mask, B, S, heads, p, seed, site, dqkv, st)
                   : launch_bwd<4>(qkv, dctx, mask, B, S, heads, p, seed, site, dqkv, st);
  if (rc) return rc;
  if (dbias != nullptr) return colsum_f16(dqkv, B * S, 3 * heads * DH, oscale, dbias, st);
  return 0;
}

}  // namespace magb200

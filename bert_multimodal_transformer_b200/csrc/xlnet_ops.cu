// MAG-XLNet specific kernels (content stream only: target_mapping / perm_mask / mems are None on the
// path the driver exercises, xlnet.py:306-313):
//   word-embedding gather + dropout                                  xlnet.py:304-305
//   relative positional encoding (sin | cos halves) + dropout        xlnet.py:93-146, 332-333
//   XLNetRelativeAttention core: ac + rel-shifted bd + segment ef scores, -1e30 mask (pad_j and i != j),
//     softmax, dropout, P V                                         (transformers 3.0.2 rel_attn_core; App. A.3)
//   last-token summary input with the final dropout                  xlnet.py:396, 506
// and their backward passes.  The relative shift is folded into indexing:
//   bd[i][j] = (q_i + r_r_bias) . kr[j + S - i]      (checked in the oracle tests)
// The attention kernels are CUDA-core kernels with the head resident in shared memory (S <= 64): they
// are 1-2 % of the FLOPs and correctness-first this round.
#include "common.cuh"
#include "host.h"

namespace magb200 {

namespace {
constexpr int H = 768;
constexpr int DH = 64;
constexpr int KP = DH + 2;   // smem pitch (halfs), conflict-free for per-lane rows
constexpr int AW = 4;

inline void drop_params(float p, uint32_t& thresh, float& scale) {
  if (p <= 0.f) { thresh = 0; scale = 1.f; return; }
  double t = (double)p * 4294967296.0;
  thresh = t >= 4294967295.0 ? 4294967295u : (uint32_t)t;
  scale = 1.0f / (1.0f - p);
}

__device__ __forceinline__ float keep_scale(uint64_t seed, uint32_t site, uint64_t idx, uint32_t thresh, float dscale) {
  if (thresh == 0) return 1.f;
  const uint4 u = philox4x32(seed, site, idx >> 2);
  const uint32_t w = (idx & 3) == 0 ? u.x : (idx & 3) == 1 ? u.y : (idx & 3) == 2 ? u.z : u.w;
  return w >= thresh ? dscale : 0.f;
}

// ---------------------------------------------------------------------------------------------
// x[t] = dropout(word[ids[t]])  (fp32 + fp16 copies); one warp per token
__global__ void __launch_bounds__(128) gather_fwd_kernel(const int64_t* __restrict__ ids, const float* __restrict__ word,
                                                         int T, int vocab, uint32_t thresh, float scale, Seed seed_in,
                                                         uint32_t site, float* __restrict__ x32, __half* __restrict__ x16) {
  const uint64_t seed = seed_in.ptr ? *seed_in.ptr : seed_in.value;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * 4 + warp;
  if (t >= T) return;
  int64_t id = ids[t];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const int c = 4 * (lane + 32 * j);
    float4 v = *reinterpret_cast<const float4*>(word + id * H + c);
    if (thresh) {
      const uint4 u = philox4x32(seed, site, ((uint64_t)t * H + c) >> 2);
      v.x = u.x >= thresh ? v.x * scale : 0.f; v.y = u.y >= thresh ? v.y * scale : 0.f;
      v.z = u.z >= thresh ? v.z * scale : 0.f; v.w = u.w >= thresh ? v.w * scale : 0.f;
    }
    *reinterpret_cast<float4*>(x32 + (int64_t)t * H + c) = v;
    __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 w2;
    w2.x = *reinterpret_cast<uint32_t*>(&a); w2.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(x16 + (int64_t)t * H + c) = w2;
  }
}
// dword[ids[t]] += oscale * keep * dx[t]
__global__ void __launch_bounds__(128) gather_bwd_kernel(const float* __restrict__ dx, const int64_t* __restrict__ ids,
                                                         int T, int vocab, uint32_t thresh, float scale, Seed seed_in,
                                                         uint32_t site, float oscale, float* __restrict__ dword) {
  const uint64_t seed = seed_in.ptr ? *seed_in.ptr : seed_in.value;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * 4 + warp;
  if (t >= T) return;
  int64_t id = ids[t];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const int c = 4 * (lane + 32 * j);
    float4 v = *reinterpret_cast<const float4*>(dx + (int64_t)t * H + c);
    if (thresh) {
      const uint4 u = philox4x32(seed, site, ((uint64_t)t * H + c) >> 2);
      v.x = u.x >= thresh ? v.x * scale : 0.f; v.y = u.y >= thresh ? v.y * scale : 0.f;
      v.z = u.z >= thresh ? v.z * scale : 0.f; v.w = u.w >= thresh ? v.w * scale : 0.f;
    }
    float* d = dword + id * H + c;
    atomicAdd(d, v.x * oscale); atomicAdd(d + 1, v.y * oscale); atomicAdd(d + 2, v.z * oscale); atomicAdd(d + 3, v.w * oscale);
  }
}

// pos16[bp][p][:] = dropout([sin(pos_p * inv_freq) | cos(pos_p * inv_freq)]), pos_p = S - p, p in [0, 2S)
__global__ void __launch_bounds__(128) pos_emb_kernel(int S, int Bp, uint32_t thresh, float scale, Seed seed_in,
                                                      uint32_t site, __half* __restrict__ pos16) {
  const uint64_t seed = seed_in.ptr ? *seed_in.ptr : seed_in.value;
  const int row = blockIdx.x;                 // bp * 2S + p
  const int p = row % (2 * S);
  const float pos = (float)(S - p);           // arange(klen, -qlen, -1.0)   xlnet.py:111,139
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    const int f = c < H / 2 ? c : c - H / 2;
    // inv_freq = 1 / 10000^(2f / d_model)   (xlnet.py:106-107), evaluated like torch.pow on fp32
    const float inv_freq = 1.0f / powf(10000.0f, (float)(2 * f) / (float)H);
    const float ang = pos * inv_freq;
    float v = c < H / 2 ? sinf(ang) : cosf(ang);
    v *= keep_scale(seed, site, (uint64_t)row * H + c, thresh, scale);
    pos16[(int64_t)row * H + c] = __float2half_rn(v);
  }
}

// out16[b][:] = dropout(x32[b][S-1][:])  : the summary reads the last token of the dropped output
__global__ void __launch_bounds__(128) last_token_fwd_kernel(const float* __restrict__ x32, int B, int S, uint32_t thresh,
                                                             float scale, Seed seed_in, uint32_t site,
                                                             __half* __restrict__ out16) {
  const uint64_t seed = seed_in.ptr ? *seed_in.ptr : seed_in.value;
  const int b = blockIdx.x;
  const int64_t t = (int64_t)b * S + (S - 1);
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    const float v = x32[t * H + c] * keep_scale(seed, site, (uint64_t)t * H + c, thresh, scale);
    out16[(int64_t)b * H + c] = __float2half_rn(v);
  }
}
// dx32[T][768] = 0 except the last-token rows: keep * dlast[b]
__global__ void __launch_bounds__(128) last_token_bwd_kernel(const float* __restrict__ dlast, int B, int S, uint32_t thresh,
                                                             float scale, Seed seed_in, uint32_t site,
                                                             float* __restrict__ dx32) {
  const uint64_t seed = seed_in.ptr ? *seed_in.ptr : seed_in.value;
  const int t = blockIdx.x;
  const int s = t % S, b = t / S;
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float v = 0.f;
    if (s == S - 1) v = dlast[(int64_t)b * H + c] * keep_scale(seed, site, (uint64_t)t * H + c, thresh, scale);
    dx32[(int64_t)t * H + c] = v;
  }
}

// Wqkv16[k][0:768 | 768:1536 | 1536:2304] = fp16(q[k][:] | k[k][:] | v[k][:])   ([in][out] blocks side by side)
__global__ void pack_qkv_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                                __half* __restrict__ out) {
  const int row = blockIdx.x;
  for (int c = threadIdx.x; c < 3 * H; c += blockDim.x) {
    const float* src = c < H ? q : (c < 2 * H ? k : v);
    out[(int64_t)row * 3 * H + c] = __float2half_rn(src[(int64_t)row * H + (c % H)]);
  }
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_head(const __half* __restrict__ base, int ld, int rows, __half* dst) {
  for (int i = threadIdx.x; i < rows * 8; i += blockDim.x) {
    const int r = i >> 3, c = (i & 7) * 8;
    const uint4 v = *reinterpret_cast<const uint4*>(base + (int64_t)r * ld + c);
    uint32_t* d = reinterpret_cast<uint32_t*>(dst + r * KP + c);
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
}

struct RelSmemPtrs {
  __half *q, *k, *v, *kr;
  float *rw, *rr, *rs, *seg;   // [64], [64], [64], [2][64]
  int* tt;
  int* pad;
};
__device__ __forceinline__ size_t rel_carve(uint8_t* base, int S, RelSmemPtrs& p) {
  size_t off = 0;
  p.q = reinterpret_cast<__half*>(base + off); off += (size_t)S * KP * 2;
  p.k = reinterpret_cast<__half*>(base + off); off += (size_t)S * KP * 2;
  p.v = reinterpret_cast<__half*>(base + off); off += (size_t)S * KP * 2;
  p.kr = reinterpret_cast<__half*>(base + off); off += (size_t)2 * S * KP * 2;
  off = (off + 15) & ~size_t(15);
  p.rw = reinterpret_cast<float*>(base + off); off += 64 * 4;
  p.rr = reinterpret_cast<float*>(base + off); off += 64 * 4;
  p.rs = reinterpret_cast<float*>(base + off); off += 64 * 4;
  p.seg = reinterpret_cast<float*>(base + off); off += 128 * 4;
  p.tt = reinterpret_cast<int*>(base + off); off += (size_t)S * 4;
  p.pad = reinterpret_cast<int*>(base + off); off += (size_t)S * 4;
  return (off + 15) & ~size_t(15);
}
__device__ __forceinline__ void rel_load_common(RelSmemPtrs& sm, const __half* qkv, const __half* kr16, int Bp,
                                                const float* r_w_bias, const float* r_r_bias, const float* r_s_bias,
                                                const float* seg_embed, const int64_t* mask, const int64_t* tt, int S,
                                                int heads, int h, int b) {
  const int Hd = heads * DH, ld = 3 * Hd;
  const __half* base = qkv + (int64_t)b * S * ld + h * DH;
  load_head(base, ld, S, sm.q);
  load_head(base + Hd, ld, S, sm.k);
  load_head(base + 2 * Hd, ld, S, sm.v);
  load_head(kr16 + (int64_t)(Bp > 1 ? b : 0) * 2 * S * Hd + h * DH, Hd, 2 * S, sm.kr);
  for (int i = threadIdx.x; i < DH; i += blockDim.x) {
    sm.rw[i] = r_w_bias[h * DH + i];
    sm.rr[i] = r_r_bias[h * DH + i];
    sm.rs[i] = r_s_bias[h * DH + i];
    sm.seg[i] = seg_embed[h * DH + i];
    sm.seg[64 + i] = seg_embed[Hd + h * DH + i];
  }
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    sm.tt[i] = tt ? (int)tt[(int64_t)b * S + i] : 0;
    sm.pad[i] = mask ? (mask[(int64_t)b * S + i] == 0) : 0;   // input_mask = 1 - attention_mask  (xlnet.py:263-264)
  }
}

// score row i for key j (pre-softmax), given qw = q_i + r_w_bias, qr = q_i + r_r_bias and ef[2]
__device__ __forceinline__ float rel_score(const RelSmemPtrs& sm, const float* qw, const float* qr, const float* ef, int i,
                                           int j, int S) {
  float ac = 0.f, bd = 0.f;
  const __half* kj = sm.k + j * KP;
  const __half* rj = sm.kr + (j + S - i) * KP;
#pragma unroll 8
  for (int d = 0; d < DH; d += 2) {
    const float2 kk = __half22float2(*reinterpret_cast<const __half2*>(kj + d));
    const float2 rr = __half22float2(*reinterpret_cast<const __half2*>(rj + d));
    ac = fmaf(qw[d], kk.x, ac); ac = fmaf(qw[d + 1], kk.y, ac);
    bd = fmaf(qr[d], rr.x, bd); bd = fmaf(qr[d + 1], rr.y, bd);
  }
  const float e = ef[sm.tt[i] != sm.tt[j] ? 1 : 0];
  float sc = (ac + bd + e) * 0.125f;
  if (sm.pad[j] && i != j) sc -= 1e30f;          // non_tgt_mask  (xlnet.py:288-296)
  return sc;
}

template <int MAXJ>
__global__ void __launch_bounds__(AW * 32) rel_attn_fwd_kernel(
    const __half* __restrict__ qkv, const __half* __restrict__ kr16, int Bp, const float* __restrict__ r_w_bias,
    const float* __restrict__ r_r_bias, const float* __restrict__ r_s_bias, const float* __restrict__ seg_embed,
    const int64_t* __restrict__ mask, const int64_t* __restrict__ tt, int S, int heads, uint32_t thresh, float dscale,
    Seed seed_in, uint32_t site, __half* __restrict__ ctx) {
  const uint64_t seed = seed_in.ptr ? *seed_in.ptr : seed_in.value;
  extern __shared__ __align__(16) uint8_t smem_rel[];
  RelSmemPtrs sm;
  size_t off = rel_carve(smem_rel, S, sm);
  float* sQW = reinterpret_cast<float*>(smem_rel + off);   // [AW][64] q_i + r_w_bias
  float* sQR = sQW + AW * 64;                              // [AW][64] q_i + r_r_bias
  float* sEF = sQR + AW * 64;                              // [AW][2]
  float* sP = sEF + AW * 2;                                // [AW][S]
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Hd = heads * DH;
  rel_load_common(sm, qkv, kr16, Bp, r_w_bias, r_r_bias, r_s_bias, seg_embed, mask, tt, S, heads, h, b);
  __syncthreads();
  float* qw = sQW + warp * 64;
  float* qr = sQR + warp * 64;
  float* ef = sEF + warp * 2;
  float* prow = sP + warp * S;
  for (int i = warp; i < S; i += AW) {
    __syncwarp();
    float e0 = 0.f, e1 = 0.f;
    for (int d = lane; d < DH; d += 32) {
      const float q = __half2float(sm.q[i * KP + d]);
      qw[d] = q + sm.rw[d];
      qr[d] = q + sm.rr[d];
      e0 += (q + sm.rs[d]) * sm.seg[d];
      e1 += (q + sm.rs[d]) * sm.seg[64 + d];
    }
    e0 = warp_sum(e0); e1 = warp_sum(e1);
    if (lane == 0) { ef[0] = e0; ef[1] = e1; }
    __syncwarp();
    float sc[MAXJ];
    float mx = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < MAXJ; ++jj) {
      const int j = lane + 32 * jj;
      sc[jj] = -INFINITY;
      if (j < S) { sc[jj] = rel_score(sm, qw, qr, ef, i, j, S); mx = fmaxf(mx, sc[jj]); }
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
#pragma unroll
    for (int jj = 0; jj < MAXJ; ++jj) {
      const int j = lane + 32 * jj;
      if (j < S) {
        const uint64_t idx = (((uint64_t)b * heads + h) * S + i) * S + j;
        prow[j] = sc[jj] * inv * keep_scale(seed, site, idx, thresh, dscale);
      }
    }
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < S; ++j) {
      const float p = prow[j];
      const float2 v = __half22float2(*reinterpret_cast<const __half2*>(sm.v + j * KP + 2 * lane));
      o0 = fmaf(p, v.x, o0);
      o1 = fmaf(p, v.y, o1);
    }
    *reinterpret_cast<__half2*>(ctx + ((int64_t)b * S + i) * Hd + h * DH + 2 * lane) = __floats2half2_rn(o0, o1);
  }
}

// Backward.  dctx (fp16, loss-scaled) -> dqkv (fp16, scaled), dkr32 (fp32 +=, scaled), and the small
// per-head parameter gradients (+= oscale * ...): r_w_bias, r_r_bias, r_s_bias, seg_embed.
template <int MAXJ>
__global__ void __launch_bounds__(AW * 32) rel_attn_bwd_kernel(
    const __half* __restrict__ qkv, const __half* __restrict__ kr16, int Bp, const float* __restrict__ r_w_bias,
    const float* __restrict__ r_r_bias, const float* __restrict__ r_s_bias, const float* __restrict__ seg_embed,
    const int64_t* __restrict__ mask, const int64_t* __restrict__ tt, const __half* __restrict__ dctx, int S, int heads,
    uint32_t thresh, float dscale, Seed seed_in, uint32_t site, float oscale, __half* __restrict__ dqkv,
    float* __restrict__ dkr32, float* __restrict__ d_rw, float* __restrict__ d_rr, float* __restrict__ d_rs,
    float* __restrict__ d_seg) {
  const uint64_t seed = seed_in.ptr ? *seed_in.ptr : seed_in.value;
  extern __shared__ __align__(16) uint8_t smem_rel[];
  RelSmemPtrs sm;
  size_t off = rel_carve(smem_rel, S, sm);
  const int SP1 = S + 1;
  float* sdO = reinterpret_cast<float*>(smem_rel + off);   // [S][65]
  float* sPd = sdO + S * 65;                               // [S][S+1] dropped probabilities
  float* sdS = sPd + S * SP1;                              // [S][S+1] d(score) * 0.125
  float* sdEF = sdS + S * SP1;                             // [S][2]
  float* sQW = sdEF + S * 2;                               // [AW][64]
  float* sQR = sQW + AW * 64;
  float* sEF = sQR + AW * 64;                              // [AW][2]
  float* sAcc = sEF + AW * 2;                              // [5][64]: d_rw | d_rr | d_rs | d_seg[0] | d_seg[1]
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Hd = heads * DH, ld = 3 * Hd;
  rel_load_common(sm, qkv, kr16, Bp, r_w_bias, r_r_bias, r_s_bias, seg_embed, mask, tt, S, heads, h, b);
  for (int i = threadIdx.x; i < S * 32; i += blockDim.x) {
    const int r = i >> 5, c = (i & 31) * 2;
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(dctx + ((int64_t)b * S + r) * Hd + h * DH + c));
    sdO[r * 65 + c] = f.x;
    sdO[r * 65 + c + 1] = f.y;
  }
  for (int i = threadIdx.x; i < 5 * 64; i += blockDim.x) sAcc[i] = 0.f;
  __syncthreads();
  float* qw = sQW + warp * 64;
  float* qr = sQR + warp * 64;
  float* ef = sEF + warp * 2;
  // ---- phase 1: per query row: P, dP, dS, dEF
  for (int i = warp; i < S; i += AW) {
    __syncwarp();
    float e0 = 0.f, e1 = 0.f;
    for (int d = lane; d < DH; d += 32) {
      const float q = __half2float(sm.q[i * KP + d]);
      qw[d] = q + sm.rw[d];
      qr[d] = q + sm.rr[d];
      e0 += (q + sm.rs[d]) * sm.seg[d];
      e1 += (q + sm.rs[d]) * sm.seg[64 + d];
    }
    e0 = warp_sum(e0); e1 = warp_sum(e1);
    if (lane == 0) { ef[0] = e0; ef[1] = e1; }
    __syncwarp();
    float sc[MAXJ], dp[MAXJ];
    float mx = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < MAXJ; ++jj) {
      const int j = lane + 32 * jj;
      sc[jj] = -INFINITY; dp[jj] = 0.f;
      if (j < S) {
        sc[jj] = rel_score(sm, qw, qr, ef, i, j, S);
        mx = fmaxf(mx, sc[jj]);
        float s = 0.f;
#pragma unroll 8
        for (int d = 0; d < DH; d += 2) {
          const float2 v = __half22float2(*reinterpret_cast<const __half2*>(sm.v + j * KP + d));
          s = fmaf(sdO[i * 65 + d], v.x, s);
          s = fmaf(sdO[i * 65 + d + 1], v.y, s);
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
        const float k = keep_scale(seed, site, idx, thresh, dscale);
        sPd[i * SP1 + j] = p * k;
        dp[jj] *= k;
        sc[jj] = p;
        dot += dp[jj] * p;
      }
    }
    dot = warp_sum(dot);
    float de0 = 0.f, de1 = 0.f;
#pragma unroll
    for (int jj = 0; jj < MAXJ; ++jj) {
      const int j = lane + 32 * jj;
      if (j < S) {
        const float ds = sc[jj] * (dp[jj] - dot) * 0.125f;
        sdS[i * SP1 + j] = ds;
        if (sm.tt[i] != sm.tt[j]) de1 += ds; else de0 += ds;
      }
    }
    de0 = warp_sum(de0); de1 = warp_sum(de1);
    if (lane == 0) { sdEF[i * 2] = de0; sdEF[i * 2 + 1] = de1; }
  }
  __syncthreads();
  // ---- phase 2: one warp per row r (as query for dQ, as key for dK/dV); lane owns head dims 2*lane, 2*lane+1
  __half* obase = dqkv + (int64_t)b * S * ld + h * DH;
  const int d0 = 2 * lane;
  float a_rw0 = 0.f, a_rw1 = 0.f, a_rr0 = 0.f, a_rr1 = 0.f, a_rs0 = 0.f, a_rs1 = 0.f;
  float a_s00 = 0.f, a_s01 = 0.f, a_s10 = 0.f, a_s11 = 0.f;
  for (int r = warp; r < S; r += AW) {
    float qa0 = 0.f, qa1 = 0.f, qb0 = 0.f, qb1 = 0.f, k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
    for (int x = 0; x < S; ++x) {
      const float ds_rx = sdS[r * SP1 + x];                 // r query, x key
      const float ds_xr = sdS[x * SP1 + r];                 // x query, r key
      const float pd_xr = sPd[x * SP1 + r];
      const float2 kk = __half22float2(*reinterpret_cast<const __half2*>(sm.k + x * KP + d0));
      const float2 rr = __half22float2(*reinterpret_cast<const __half2*>(sm.kr + (x + S - r) * KP + d0));
      const float2 qq = __half22float2(*reinterpret_cast<const __half2*>(sm.q + x * KP + d0));
      qa0 = fmaf(ds_rx, kk.x, qa0); qa1 = fmaf(ds_rx, kk.y, qa1);
      qb0 = fmaf(ds_rx, rr.x, qb0); qb1 = fmaf(ds_rx, rr.y, qb1);
      k0 = fmaf(ds_xr, qq.x + sm.rw[d0], k0); k1 = fmaf(ds_xr, qq.y + sm.rw[d0 + 1], k1);
      v0 = fmaf(pd_xr, sdO[x * 65 + d0], v0); v1 = fmaf(pd_xr, sdO[x * 65 + d0 + 1], v1);
    }
    const float de0 = sdEF[r * 2], de1 = sdEF[r * 2 + 1];
    const float qs0 = de0 * sm.seg[d0] + de1 * sm.seg[64 + d0], qs1 = de0 * sm.seg[d0 + 1] + de1 * sm.seg[64 + d0 + 1];
    const float2 qr_ = __half22float2(*reinterpret_cast<const __half2*>(sm.q + r * KP + d0));
    a_rw0 += qa0; a_rw1 += qa1; a_rr0 += qb0; a_rr1 += qb1; a_rs0 += qs0; a_rs1 += qs1;
    a_s00 += de0 * (qr_.x + sm.rs[d0]); a_s01 += de0 * (qr_.y + sm.rs[d0 + 1]);
    a_s10 += de1 * (qr_.x + sm.rs[d0]); a_s11 += de1 * (qr_.y + sm.rs[d0 + 1]);
    __half* o = obase + (int64_t)r * ld + d0;
    *reinterpret_cast<__half2*>(o) = __floats2half2_rn(qa0 + qb0 + qs0, qa1 + qb1 + qs1);
    *reinterpret_cast<__half2*>(o + Hd) = __floats2half2_rn(k0, k1);
    *reinterpret_cast<__half2*>(o + 2 * Hd) = __floats2half2_rn(v0, v1);
  }
  atomicAdd(sAcc + d0, a_rw0); atomicAdd(sAcc + d0 + 1, a_rw1);
  atomicAdd(sAcc + 64 + d0, a_rr0); atomicAdd(sAcc + 64 + d0 + 1, a_rr1);
  atomicAdd(sAcc + 128 + d0, a_rs0); atomicAdd(sAcc + 128 + d0 + 1, a_rs1);
  atomicAdd(sAcc + 192 + d0, a_s00); atomicAdd(sAcc + 192 + d0 + 1, a_s01);
  atomicAdd(sAcc + 256 + d0, a_s10); atomicAdd(sAcc + 256 + d0 + 1, a_s11);
  // ---- phase 3: dkr[p] += sum_i dS[i][p - S + i] * (q_i + r_r_bias)
  float* dkr_base = dkr32 + (int64_t)(Bp > 1 ? b : 0) * 2 * S * Hd + h * DH;
  for (int p = warp; p < 2 * S; p += AW) {
    float r0 = 0.f, r1 = 0.f;
    for (int i = 0; i < S; ++i) {
      const int j = p - S + i;
      if (j < 0 || j >= S) continue;
      const float ds = sdS[i * SP1 + j];
      const float2 qq = __half22float2(*reinterpret_cast<const __half2*>(sm.q + i * KP + d0));
      r0 = fmaf(ds, qq.x + sm.rr[d0], r0);
      r1 = fmaf(ds, qq.y + sm.rr[d0 + 1], r1);
    }
    atomicAdd(dkr_base + (int64_t)p * Hd + d0, r0);
    atomicAdd(dkr_base + (int64_t)p * Hd + d0 + 1, r1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64; i += blockDim.x) {
    atomicAdd(d_rw + h * DH + i, sAcc[i] * oscale);
    atomicAdd(d_rr + h * DH + i, sAcc[64 + i] * oscale);
    atomicAdd(d_rs + h * DH + i, sAcc[128 + i] * oscale);
    atomicAdd(d_seg + h * DH + i, sAcc[192 + i] * oscale);
    atomicAdd(d_seg + Hd + h * DH + i, sAcc[256 + i] * oscale);
  }
}

size_t rel_common_bytes(int S) {
  size_t off = (size_t)S * KP * 2 * 3 + (size_t)2 * S * KP * 2;
  off = (off + 15) & ~size_t(15);
  off += 64 * 4 * 3 + 128 * 4 + (size_t)S * 4 * 2;
  return (off + 15) & ~size_t(15);
}

}  // namespace

int xl_gather_fwd(const int64_t* ids, const float* word, int T, int vocab, float p, Seed seed, uint32_t site, float* x32,
                  __half* x16, cudaStream_t st) {
  uint32_t th; float sc;
  drop_params(p, th, sc);
  gather_fwd_kernel<<<(T + 3) / 4, 128, 0, st>>>(ids, word, T, vocab, th, sc, seed, site, x32, x16);
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
int xl_gather_bwd(const float* dx, const int64_t* ids, int T, int vocab, float p, Seed seed, uint32_t site, float oscale,
                  float* dword, cudaStream_t st) {
  uint32_t th; float sc;
  drop_params(p, th, sc);
  gather_bwd_kernel<<<(T + 3) / 4, 128, 0, st>>>(dx, ids, T, vocab, th, sc, seed, site, oscale, dword);
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
int xl_pos_emb(int S, int Bp, float p, Seed seed, uint32_t site, __half* pos16, cudaStream_t st) {
  uint32_t th; float sc;
  drop_params(p, th, sc);
  pos_emb_kernel<<<Bp * 2 * S, 128, 0, st>>>(S, Bp, th, sc, seed, site, pos16);
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
int xl_last_token_fwd(const float* x32, int B, int S, float p, Seed seed, uint32_t site, __half* out16, cudaStream_t st) {
  uint32_t th; float sc;
  drop_params(p, th, sc);
  last_token_fwd_kernel<<<B, 128, 0, st>>>(x32, B, S, th, sc, seed, site, out16);
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
int xl_last_token_bwd(const float* dlast, int B, int S, float p, Seed seed, uint32_t site, float* dx32, cudaStream_t st) {
  uint32_t th; float sc;
  drop_params(p, th, sc);
  last_token_bwd_kernel<<<B * S, 128, 0, st>>>(dlast, B, S, th, sc, seed, site, dx32);
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
int xl_pack_qkv(const float* q, const float* k, const float* v, __half* out, cudaStream_t st) {
  pack_qkv_kernel<<<H, 256, 0, st>>>(q, k, v, out);
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int xl_rel_attn_fwd(const RelAttnArgs& a, __half* ctx, cudaStream_t st) {
  MAGB_REQUIRE(a.head_dim == DH, "rel_attn: head_dim must be 64");
  MAGB_REQUIRE(a.S >= 1 && a.S <= 64, "rel_attn: sequence length %d out of range [1,64]", a.S);
  uint32_t th; float ds;
  drop_params(a.p, th, ds);
  const size_t smem = rel_common_bytes(a.S) + (size_t)(AW * 64 * 2 + AW * 2 + AW * a.S) * 4;
  auto k = rel_attn_fwd_kernel<2>;
  static size_t configured_by_dev[kMaxDevices] = {0};
  size_t& configured = configured_by_dev[current_device_slot()];
  if (smem > configured) {
    MAGB_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  k<<<dim3(a.heads, a.B), AW * 32, smem, st>>>(a.qkv, a.kr16, a.Bp, a.r_w_bias, a.r_r_bias, a.r_s_bias, a.seg_embed, a.mask,
                                              a.tt, a.S, a.heads, th, ds, a.seed, a.site, ctx);
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int xl_rel_attn_bwd(const RelAttnArgs& a, const __half* dctx, float oscale, __half* dqkv, float* dkr32, float* d_rw,
                    float* d_rr, float* d_rs, float* d_seg, cudaStream_t st) {
  MAGB_REQUIRE(a.head_dim == DH, "rel_attn: head_dim must be 64");
  MAGB_REQUIRE(a.S >= 1 && a.S <= 64, "rel_attn: sequence length %d out of range [1,64]", a.S);
  uint32_t th; float ds;
  drop_params(a.p, th, ds);
  const int S = a.S;
  const size_t smem = rel_common_bytes(S) +
                      (size_t)(S * 65 + 2 * S * (S + 1) + S * 2 + AW * 64 * 2 + AW * 2 + 5 * 64) * 4;
  auto k = rel_attn_bwd_kernel<2>;
  static size_t configured_by_dev[kMaxDevices] = {0};
  size_t& configured = configured_by_dev[current_device_slot()];
  if (smem > configured) {
    MAGB_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  k<<<dim3(a.heads, a.B), AW * 32, smem, st>>>(a.qkv, a.kr16, a.Bp, a.r_w_bias, a.r_r_bias, a.r_s_bias, a.seg_embed, a.mask,
                                              a.tt, dctx, S, a.heads, th, ds, a.seed, a.site, oscale, dqkv, dkr32, d_rw,
                                              d_rr, d_rs, d_seg);
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace magb200

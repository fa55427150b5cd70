// Row-wise (per-token) kernels of the MAG-BERT path, hidden size 768 = 32 lanes x 24 floats:
//   embedding gather + LayerNorm (+dropout)           BertEmbeddings  (bert.py:211-216)
//   MAG tail: gates, norms, alpha, shift, LayerNorm   MAG.forward     (modeling.py:27-49)
//   dropout + residual + LayerNorm                    BertSelfOutput / BertOutput (App. A.2)
// and their backward passes.  One warp owns one token row; every global access is a float4 /
// 8-byte bf16x4 / fp16x4 vector, coalesced across the warp.  HBM-bound: algorithmic bytes per
// token are listed in DESIGN.md.
#include "common.cuh"
#include "host.h"
#include "rowdev.cuh"

namespace magb200 {

// --------------------------------------------------------------------------------------------
// Embedding gather + LayerNorm + dropout.  Also packs the MAG GEMM A operand
//   Amag[t] = [ fp16(emb[t]) (768) | fp16(visual[t]) (V) | fp16(acoustic[t]) (A) | 0 ... ] (pitch KA)
__global__ void __launch_bounds__(128) embed_ln_fwd_kernel(
    const int64_t* __restrict__ ids, const int64_t* __restrict__ tt, const float* __restrict__ visual,
    const float* __restrict__ acoustic, const float* __restrict__ word, const float* __restrict__ pos,
    const float* __restrict__ type, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
    uint32_t thresh, float scale, Seed seed_in, uint32_t site, int V, int A, int KA, int T, int S, int vocab,
    float* __restrict__ emb32, __half* __restrict__ amag) {
  pdl_trigger();   // the next kernel in the stream (a GEMM) may start its private prologue
  const uint64_t seed = seed_in.ptr ? *seed_in.ptr : seed_in.value;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * 4 + warp;
  if (t >= T) return;
  int64_t id = ids[t];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const int64_t ty = tt ? tt[t] : 0;
  const int s = t % S;
  Row r, a;
  load_row(r, word + id * H, lane);
  load_row(a, pos + (int64_t)s * H, lane);
#pragma unroll
  for (int i = 0; i < 4 * VEC; ++i) r.v[i] += a.v[i];
  load_row(a, type + ty * H, lane);
#pragma unroll
  for (int i = 0; i < 4 * VEC; ++i) r.v[i] += a.v[i];
  float mean, rstd;
  row_stats(r, eps, mean, rstd);
  Row g, b;
  load_row(g, gamma, lane);
  load_row(b, beta, lane);
#pragma unroll
  for (int i = 0; i < 4 * VEC; ++i) r.v[i] = (r.v[i] - mean) * rstd * g.v[i] + b.v[i];
  row_dropout(r, seed, site, t, lane, thresh, scale);
  store_row(r, emb32 + (int64_t)t * H, lane);
  if (amag != nullptr) {
    __half* arow = amag + (int64_t)t * KA;
    store_row_f16(r, arow, lane);
    for (int c = lane; c < KA - H; c += 32) {
      float x = 0.f;
      if (c < V) x = visual[(int64_t)t * V + c];
      else if (c < V + A) x = acoustic[(int64_t)t * A + (c - V)];
      arow[H + c] = __float2half_rn(x);
    }
  }
}

// Packs [text | visual | acoustic | 0] for the standalone MAG op (text given, no embedding stage).
__global__ void __launch_bounds__(128) mag_pack_inputs_kernel(const float* __restrict__ text,
                                                              const float* __restrict__ visual,
                                                              const float* __restrict__ acoustic, int V, int A, int KA,
                                                              int T, __half* __restrict__ amag) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * 4 + warp;
  if (t >= T) return;
  Row r;
  load_row(r, text + (int64_t)t * H, lane);
  __half* arow = amag + (int64_t)t * KA;
  store_row_f16(r, arow, lane);
  for (int c = lane; c < KA - H; c += 32) {
    float x = 0.f;
    if (c < V) x = visual[(int64_t)t * V + c];
    else if (c < V + A) x = acoustic[(int64_t)t * A + (c - V)];
    arow[H + c] = __float2half_rn(x);
  }
}

// Split-precision packing for the exact MAG forward: x = hi + lo with hi = x truncated to TF32 (low 13 mantissa bits
// cleared, so the tensor core reads it exactly) and lo = x - hi (exact in fp32; the tensor core truncates it to 10
// mantissa bits: relative error 2^-21 of x).  A*B ~= A_hi*B_hi + A_hi*B_lo + A_lo*B_hi, the dropped lo*lo term is 2^-22.
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  lo = x - hi;
}
__global__ void __launch_bounds__(128) mag_pack_inputs_split_kernel(const float* __restrict__ text,
                                                                    const float* __restrict__ visual,
                                                                    const float* __restrict__ acoustic, int V, int A,
                                                                    int KA, int T, float* __restrict__ a_hi,
                                                                    float* __restrict__ a_lo) {
  const int t = blockIdx.x;
  if (t >= T) return;
  for (int c = threadIdx.x; c < KA; c += blockDim.x) {
    float x = 0.f;
    if (c < H) x = text[(int64_t)t * H + c];
    else if (c < H + V) x = visual[(int64_t)t * V + (c - H)];
    else if (c < H + V + A) x = acoustic[(int64_t)t * A + (c - H - V)];
    float hi, lo;
    split_tf32(x, hi, lo);
    a_hi[(int64_t)t * KA + c] = hi;
    a_lo[(int64_t)t * KA + c] = lo;
  }
}
__global__ void mag_pack_weights_split_kernel(const float* __restrict__ W_hv, const float* __restrict__ W_ha,
                                              const float* __restrict__ W_v, const float* __restrict__ W_a, int V, int A,
                                              int KA, float* __restrict__ b_hi, float* __restrict__ b_lo) {
  const int row = blockIdx.x;  // 0 .. 4H, same layout as mag_pack_weights_kernel
  const int blk = row / H, n = row % H;
  for (int k = threadIdx.x; k < KA; k += blockDim.x) {
    float x = 0.f;
    if (k < H) {
      if (blk == 0) x = W_hv[(int64_t)n * (V + H) + V + k];
      else if (blk == 1) x = W_ha[(int64_t)n * (A + H) + A + k];
    } else if (k < H + V) {
      const int j = k - H;
      if (blk == 0) x = W_hv[(int64_t)n * (V + H) + j];
      else if (blk == 2) x = W_v[(int64_t)n * V + j];
    } else if (k < H + V + A) {
      const int j = k - H - V;
      if (blk == 1) x = W_ha[(int64_t)n * (A + H) + j];
      else if (blk == 3) x = W_a[(int64_t)n * A + j];
    }
    float hi, lo;
    split_tf32(x, hi, lo);
    b_hi[(int64_t)row * KA + k] = hi;
    b_lo[(int64_t)row * KA + k] = lo;
  }
}

// Backward of embedding LayerNorm+dropout and scatter-add into the three tables.
// Block (s, z) owns sequence position s: its 8 warps walk the batch rows b = z*8 + warp, z*8 + warp + 8*gridDim.y, ...
// so the position-table gradient, the LayerNorm parameter gradients and the token-type gradients stay in register
// accumulators until one block-wide reduction at the end.  (The first version issued three scalar atomics per
// element: 5.5 M atomics, 141 us under ncu; the type table alone took T-way contention on 2 x 768 addresses.)
// Only the word-table scatter remains atomic, as 16-byte vector reductions.
// Token types: type 1 has its own accumulator, type 0 is "everything minus type 1"; a type >= 2 (not produced by
// the reference's data pipeline, but legal for the table) goes through atomics and is subtracted from type 0 again.
__global__ void __launch_bounds__(256) embed_ln_bwd_kernel(
    const float* __restrict__ dout, const int64_t* __restrict__ ids, const int64_t* __restrict__ tt,
    const float* __restrict__ word, const float* __restrict__ pos, const float* __restrict__ type,
    const float* __restrict__ gamma, float eps, uint32_t thresh, float scale, Seed seed_in, uint32_t site, int B,
    int S, int vocab, int type_rows, float oscale, float* __restrict__ dword, float* __restrict__ dpos, float* __restrict__ dtype,
    float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const uint64_t seed = seed_in.ptr ? *seed_in.ptr : seed_in.value;
  __shared__ float red[8 * H];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x;
  Row accp, acc1, accg, accb, prow;
  row_zero(accp); row_zero(acc1); row_zero(accg); row_zero(accb);
  load_row(prow, pos + (int64_t)s * H, lane);
  for (int b = blockIdx.y * 8 + warp; b < B; b += gridDim.y * 8) {
    const int t = b * S + s;
    int64_t id = ids[t];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    int64_t ty = tt ? tt[t] : 0;
    ty = ty < 0 ? 0 : (ty >= type_rows ? type_rows - 1 : ty);
    Row z, a;
    load_row(z, word + id * H, lane);
    load_row(a, type + ty * H, lane);
#pragma unroll
    for (int i = 0; i < 4 * VEC; ++i) z.v[i] += prow.v[i] + a.v[i];
    float mean, rstd;
    row_stats(z, eps, mean, rstd);
    Row dy;
    load_row(dy, dout + (int64_t)t * H, lane);
    row_dropout(dy, seed, site, t, lane, thresh, scale);
    load_row(a, gamma, lane);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 4 * VEC; ++i) {
      z.v[i] = (z.v[i] - mean) * rstd;  // xhat
      accg.v[i] += dy.v[i] * z.v[i];
      accb.v[i] += dy.v[i];
      dy.v[i] *= a.v[i];
      s1 += dy.v[i];
      s2 += dy.v[i] * z.v[i];
    }
    const float m1 = warp_sum(s1) * (1.0f / H);
    const float m2 = warp_sum(s2) * (1.0f / H);
    const float w1 = ty == 1 ? 1.f : 0.f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      float dz[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        dz[k] = rstd * (dy.v[4 * j + k] - m1 - z.v[4 * j + k] * m2);
        accp.v[4 * j + k] += dz[k];
        acc1.v[4 * j + k] += w1 * dz[k];
        dz[k] *= oscale;
      }
      const int c = col_of(lane, j);
      red_add_f32x4(dword + id * H + c, dz[0], dz[1], dz[2], dz[3]);
      if (ty >= 2) {
        red_add_f32x4(dtype + ty * H + c, dz[0], dz[1], dz[2], dz[3]);
        red_add_f32x4(dtype + c, -dz[0], -dz[1], -dz[2], -dz[3]);
      }
    }
  }
  block_reduce_flush<8>(accg, red, dgamma, oscale, warp, lane);
  block_reduce_flush<8>(accb, red, dbeta, oscale, warp, lane);
  block_reduce_flush<8>(accp, red, dpos + (int64_t)s * H, oscale, warp, lane);
  if (type_rows >= 2) block_reduce_flush<8>(acc1, red, dtype + H, oscale, warp, lane);
#pragma unroll
  for (int i = 0; i < 4 * VEC; ++i) accp.v[i] -= acc1.v[i];
  block_reduce_flush<8>(accp, red, dtype, oscale, warp, lane);
}

// --------------------------------------------------------------------------------------------
// MAG tail (modeling.py:27-49) on the gate GEMM output G[t] = [gate_v_pre | gate_a_pre | W_v v | W_a a]
// (biases already added by the GEMM epilogue).
struct MagRow {
  Row hm;       // h_m
  float alpha;  // min(thresh, 1)
  float thr, em, hm_raw, hm_n;
};

__device__ __forceinline__ void mag_row_forward(const float* __restrict__ Gt, const Row& text, float beta_shift,
                                                int lane, Row& gv, Row& ga, Row& hv, Row& ha, MagRow& m) {
  load_row(gv, Gt, lane);
  load_row(ga, Gt + H, lane);
  load_row(hv, Gt + 2 * H, lane);
  load_row(ha, Gt + 3 * H, lane);
#pragma unroll
  for (int i = 0; i < 4 * VEC; ++i)
    m.hm.v[i] = fmaxf(gv.v[i], 0.f) * hv.v[i] + fmaxf(ga.v[i], 0.f) * ha.v[i];       // modeling.py:27-30
  m.em = sqrtf(row_dot(text, text));                                                   // :32
  m.hm_raw = sqrtf(row_dot(m.hm, m.hm));                                               // :33
  m.hm_n = m.hm_raw == 0.f ? 1.f : m.hm_raw;                                           // :35-36
  m.thr = m.em / (m.hm_n + 1e-6f) * beta_shift;                                        // :38 (eps :26)
  m.alpha = fminf(m.thr, 1.f);                                                         // :40-42
}

__global__ void __launch_bounds__(128) mag_tail_fwd_kernel(const float* __restrict__ G, const float* __restrict__ text32,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float beta_shift,
                                                           float eps, uint32_t thresh, float scale, Seed seed_in,
                                                           uint32_t site, int T, float* __restrict__ out32,
                                                           __half* __restrict__ out16) {
  pdl_trigger();   // the next kernel in the stream (a GEMM) may start its private prologue
  const uint64_t seed = seed_in.ptr ? *seed_in.ptr : seed_in.value;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * 4 + warp;
  if (t >= T) return;
  Row text, gv, ga, hv, ha;
  MagRow m;
  load_row(text, text32 + (int64_t)t * H, lane);
  mag_row_forward(G + (int64_t)t * 4 * H, text, beta_shift, lane, gv, ga, hv, ha, m);
  Row s;
#pragma unroll
  for (int i = 0; i < 4 * VEC; ++i) s.v[i] = m.alpha * m.hm.v[i] + text.v[i];          // :43-48
  float mean, rstd;
  row_stats(s, eps, mean, rstd);
  load_row(gv, gamma, lane);
  load_row(ga, beta, lane);
#pragma unroll
  for (int i = 0; i < 4 * VEC; ++i) s.v[i] = (s.v[i] - mean) * rstd * gv.v[i] + ga.v[i];
  row_dropout(s, seed, site, t, lane, thresh, scale);                                  // :47-49
  store_row(s, out32 + (int64_t)t * H, lane);
  if (out16) store_row_f16(s, out16 + (int64_t)t * H, lane);
}

__global__ void __launch_bounds__(128) mag_tail_bwd_kernel(
    const float* __restrict__ dout, const float* __restrict__ G, const float* __restrict__ text32,
    const float* __restrict__ gamma, float beta_shift, float eps, uint32_t thresh, float scale, Seed seed_in,
    uint32_t site, int T, float oscale, __half* __restrict__ dG, float* __restrict__ dtext,
    float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias4) {
  pdl_trigger();   // the next kernel in the stream (a GEMM) may start its private prologue
  const uint64_t seed = seed_in.ptr ? *seed_in.ptr : seed_in.value;
  __shared__ float acc[6 * H];   // dgamma | dbeta | dbias of the four gate-GEMM column blocks
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  acc_zero(acc, 6 * H);
  for (int t = blockIdx.x * 4 + warp; t < T; t += gridDim.x * 4) {
    Row text, gv, ga, hv, ha;
    MagRow m;
    load_row(text, text32 + (int64_t)t * H, lane);
    mag_row_forward(G + (int64_t)t * 4 * H, text, beta_shift, lane, gv, ga, hv, ha, m);
    Row xh;
#pragma unroll
    for (int i = 0; i < 4 * VEC; ++i) xh.v[i] = m.alpha * m.hm.v[i] + text.v[i];
    float mean, rstd;
    row_stats(xh, eps, mean, rstd);
    Row dy, gm;
    load_row(dy, dout + (int64_t)t * H, lane);
    row_dropout(dy, seed, site, t, lane, thresh, scale);
    load_row(gm, gamma, lane);
    Row gy;
#pragma unroll
    for (int i = 0; i < 4 * VEC; ++i) {
      xh.v[i] = (xh.v[i] - mean) * rstd;
      gy.v[i] = dy.v[i] * xh.v[i];
    }
    acc_add_row(acc, gy, lane);
    acc_add_row(acc + H, dy, lane);
#pragma unroll
    for (int i = 0; i < 4 * VEC; ++i) gy.v[i] = dy.v[i] * gm.v[i];
    const float m1 = row_sum(gy) * (1.0f / H);
    const float m2 = row_dot(gy, xh) * (1.0f / H);
    Row ds;
#pragma unroll
    for (int i = 0; i < 4 * VEC; ++i) ds.v[i] = rstd * (gy.v[i] - m1 - xh.v[i] * m2);
    // alpha path: alpha = min(thr, 1); torch.min splits the gradient evenly on an exact tie
    const float dalpha = row_dot(ds, m.hm);
    const float dthr = m.thr < 1.f ? dalpha : (m.thr == 1.f ? 0.5f * dalpha : 0.f);
    const float inv = 1.0f / (m.hm_n + 1e-6f);
    const float dem = dthr * beta_shift * inv;
    const float dhmn = -dthr * beta_shift * m.em * inv * inv;
    const float ctext = m.em > 0.f ? dem / m.em : 0.f;
    const float chm = m.hm_raw != 0.f ? dhmn / m.hm_raw : 0.f;   // where(hm==0, ones, hm): constant branch
    Row o0, o1, o2, o3;
#pragma unroll
    for (int i = 0; i < 4 * VEC; ++i) {
      const float dt = ds.v[i] + ctext * text.v[i];
      const float dhm = m.alpha * ds.v[i] + chm * m.hm.v[i];
      text.v[i] = dt;
      o0.v[i] = gv.v[i] > 0.f ? dhm * hv.v[i] : 0.f;
      o1.v[i] = ga.v[i] > 0.f ? dhm * ha.v[i] : 0.f;
      o2.v[i] = dhm * fmaxf(gv.v[i], 0.f);
      o3.v[i] = dhm * fmaxf(ga.v[i], 0.f);
    }
    acc_add_row(acc + 2 * H, o0, lane);
    acc_add_row(acc + 3 * H, o1, lane);
    acc_add_row(acc + 4 * H, o2, lane);
    acc_add_row(acc + 5 * H, o3, lane);
    store_row(text, dtext + (int64_t)t * H, lane);
    __half* dg = dG + (int64_t)t * 4 * H;
    store_row_f16(o0, dg, lane);
    store_row_f16(o1, dg + H, lane);
    store_row_f16(o2, dg + 2 * H, lane);
    store_row_f16(o3, dg + 3 * H, lane);
  }
  __syncthreads();
  acc_flush(acc, dgamma, H, oscale);
  acc_flush(acc + H, dbeta, H, oscale);
  acc_flush(acc + 2 * H, dbias4, 4 * H, oscale);
}

// --------------------------------------------------------------------------------------------
// y = LayerNorm(dropout(pre) + res)   (BertSelfOutput / BertOutput; XLNet post_attention / ff)
__global__ void __launch_bounds__(128) dropout_residual_ln_fwd_kernel(
    const float* __restrict__ pre, const float* __restrict__ res, const float* __restrict__ gamma,
    const float* __restrict__ beta, float eps, uint32_t thresh, float scale, Seed seed_in, uint32_t site, int T,
    float* __restrict__ y32, __half* __restrict__ y16, float* __restrict__ rstd_out, float* zero_after) {
  pdl_wait();      // may be launched with the PDL attribute: the predecessor can still be running
  pdl_trigger();   // the next kernel in the stream (a GEMM) may start its private prologue
  const uint64_t seed = seed_in.ptr ? *seed_in.ptr : seed_in.value;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * 4 + warp;
  if (t >= T) return;
  Row z, r;
  if (zero_after == nullptr) {
    load_row(z, pre + (int64_t)t * H, lane);
  } else {
    // `pre` is a reduction target (stream-K GEMM, red.add epilogue): hand it back zeroed for the next producer.
    // zero_after aliases pre, so both accesses are volatile asm with a memory clobber: the compiler must not move
    // the (restrict-qualified) row load below the store.
    const float* src = pre + (int64_t)t * H;
    float* dst = zero_after + (int64_t)t * H;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      float4 v;
      asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                   : "l"(src + col_of(lane, j)) : "memory");
      z.v[4 * j] = v.x; z.v[4 * j + 1] = v.y; z.v[4 * j + 2] = v.z; z.v[4 * j + 3] = v.w;
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j)
      asm volatile("st.global.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(dst + col_of(lane, j)), "f"(0.f) : "memory");
  }
  row_dropout(z, seed, site, t, lane, thresh, scale);
  load_row(r, res + (int64_t)t * H, lane);
#pragma unroll
  for (int i = 0; i < 4 * VEC; ++i) z.v[i] += r.v[i];
  float mean, rstd;
  row_stats(z, eps, mean, rstd);
  Row g, b;
  load_row(g, gamma, lane);
  load_row(b, beta, lane);
#pragma unroll
  for (int i = 0; i < 4 * VEC; ++i) z.v[i] = (z.v[i] - mean) * rstd * g.v[i] + b.v[i];
  store_row(z, y32 + (int64_t)t * H, lane);
  if (y16) store_row_f16(z, y16 + (int64_t)t * H, lane);
  if (rstd_out && lane == 0) rstd_out[t] = rstd;
}

// Backward of the above from the saved output y (xhat = (y - beta) / gamma) and rstd.
//   dz   (fp32)  : gradient wrt the LayerNorm input = gradient of the residual branch
//   dpre (fp16)  : dropout-masked dz = gradient wrt the GEMM output (operand of dgrad / wgrad GEMMs)
//   dgamma, dbeta, dbias (+= column sums)
constexpr int LNB_WARPS = 16;
__global__ void __launch_bounds__(LNB_WARPS * 32, 1) ln_bwd_kernel(const float* __restrict__ dy32, const float* __restrict__ y32,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     const float* __restrict__ rstd_in, uint32_t thresh, float scale,
                                                     Seed seed_in, uint32_t site, int T, float oscale,
                                                     float* __restrict__ dz32, __half* __restrict__ dpre,
                                                     float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                     float* __restrict__ dbias) {
  pdl_wait();      // may be launched with the PDL attribute: the predecessor can still be running
  pdl_trigger();   // the next kernel in the stream (a GEMM) may start its private prologue
  // One block of 16 warps per SM, (almost) one row per warp: at T = 2400 the kernel is a latency chain per row
  // (load 6 KB -> two reductions -> store), so rows in flight per SM is what matters.  ncu on the 8-warp version that
  // kept the three column sums in register accumulators and prefetched the next row: 255 registers, 1 block x 8 warps
  // per SM, 12 % warps active, 13.8 us.  Here the column sums (dgamma, dbeta, dbias) live in a per-warp shared-memory
  // slot, so a thread holds two rows at most; the 16 slots are combined once per block.
  const uint64_t seed = seed_in.ptr ? *seed_in.ptr : seed_in.value;
  extern __shared__ __align__(16) float lnb_slots[];          // [LNB_WARPS][3][H]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* slot = lnb_slots + warp * 3 * H;
  Row g, b;
  load_row(g, gamma, lane);
  load_row(b, beta, lane);
  bool first = true;
  for (int t = blockIdx.x * LNB_WARPS + warp; t < T; t += gridDim.x * LNB_WARPS) {
    Row dy, xh;
    load_row(dy, dy32 + (int64_t)t * H, lane);
    load_row(xh, y32 + (int64_t)t * H, lane);
    const float rstd = rstd_in[t];
    Row cg;                       // this row's contribution to dgamma; dbeta's is dy itself
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 4 * VEC; ++i) {
      xh.v[i] = fabsf(g.v[i]) > 1e-20f ? (xh.v[i] - b.v[i]) / g.v[i] : 0.f;
      cg.v[i] = dy.v[i] * xh.v[i];
    }
    if (first) {
      store_row(cg, slot, lane);
      store_row(dy, slot + H, lane);
    } else {
      Row a;
      load_row(a, slot, lane);
#pragma unroll
      for (int i = 0; i < 4 * VEC; ++i) a.v[i] += cg.v[i];
      store_row(a, slot, lane);
      load_row(a, slot + H, lane);
#pragma unroll
      for (int i = 0; i < 4 * VEC; ++i) a.v[i] += dy.v[i];
      store_row(a, slot + H, lane);
    }
#pragma unroll
    for (int i = 0; i < 4 * VEC; ++i) {
      dy.v[i] *= g.v[i];
      s1 += dy.v[i];
      s2 += dy.v[i] * xh.v[i];
    }
    const float m1 = warp_sum(s1) * (1.0f / H);
    const float m2 = warp_sum(s2) * (1.0f / H);
#pragma unroll
    for (int i = 0; i < 4 * VEC; ++i) dy.v[i] = rstd * (dy.v[i] - m1 - xh.v[i] * m2);
    store_row(dy, dz32 + (int64_t)t * H, lane);
    row_dropout(dy, seed, site, t, lane, thresh, scale);
    if (first) {
      store_row(dy, slot + 2 * H, lane);
    } else {
      Row a;
      load_row(a, slot + 2 * H, lane);
#pragma unroll
      for (int i = 0; i < 4 * VEC; ++i) a.v[i] += dy.v[i];
      store_row(a, slot + 2 * H, lane);
    }
    store_row_f16(dy, dpre + (int64_t)t * H, lane);
    first = false;
  }
  if (first) {                    // a warp without rows still owns a slot the block reduction reads
    Row z;
    row_zero(z);
    store_row(z, slot, lane);
    store_row(z, slot + H, lane);
    store_row(z, slot + 2 * H, lane);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 3 * H; c += LNB_WARPS * 32) {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < LNB_WARPS; ++w) sum += lnb_slots[w * 3 * H + c];
    float* out = c < H ? dgamma + c : (c < 2 * H ? dbeta + (c - H) : (dbias ? dbias + (c - 2 * H) : nullptr));
    if (out != nullptr) atomicAdd(out, sum * oscale);
  }
}

// --------------------------------------------------------------------------------------------
// out[n] += sum_t in[t][n]   (bias gradients of the QKV and FFN-in projections), fp16 input
__global__ void __launch_bounds__(256) colsum_f16_kernel(const __half* __restrict__ in, int T, int N, float oscale,
                                                         float* __restrict__ out) {
  pdl_wait();      // may be launched with the PDL attribute: the predecessor can still be running
  pdl_trigger();   // the next kernel in the stream (a GEMM) may start its private prologue
  // block = 8 warps; a warp covers 256 columns (16 B per lane); warps and blockIdx.y split the rows
  __shared__ float red[8][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 256 + 8 * lane;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  if (c < N) {
    const int stride = gridDim.y * 8;
    int t = blockIdx.y * 8 + warp;
    for (; t + 3 * stride < T; t += 4 * stride) {
      uint4 u[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) u[q] = *reinterpret_cast<const uint4*>(in + (int64_t)(t + q * stride) * N + c);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t w[4] = {u[q].x, u[q].y, u[q].z, u[q].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
          acc[2 * k] += f.x;
          acc[2 * k + 1] += f.y;
        }
      }
    }
    for (; t < T; t += stride) {
      const uint4 u = *reinterpret_cast<const uint4*>(in + (int64_t)t * N + c);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
        acc[2 * k] += f.x;
        acc[2 * k + 1] += f.y;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[warp][8 * lane + k] = acc[k];
  __syncthreads();
  {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += red[w][threadIdx.x];
    const int cc = blockIdx.x * 256 + threadIdx.x;
    if (cc < N) atomicAdd(out + cc, sum * oscale);
  }
}

// --------------------------------------------------------------------------------------------
__global__ void cast_f32_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(in)[i];
    __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    reinterpret_cast<uint2*>(out)[i] = u;
  }
}

// MAG gate weights -> the block matrix B[4H][KA] (fp16) consumed by the gate GEMM
//   rows [0,H)   : [ W_hv[:, V:]  | W_hv[:, :V] | 0          | 0 ]     (modeling.py:27, cat(visual, text))
//   rows [H,2H)  : [ W_ha[:, A:]  | 0           | W_ha[:, :A]| 0 ]     (modeling.py:28)
//   rows [2H,3H) : [ 0            | W_v         | 0          | 0 ]     (modeling.py:30)
//   rows [3H,4H) : [ 0            | 0           | W_a        | 0 ]
__global__ void mag_pack_weights_kernel(const float* __restrict__ W_hv, const float* __restrict__ W_ha,
                                        const float* __restrict__ W_v, const float* __restrict__ W_a, int V, int A,
                                        int KA, __half* __restrict__ Bm) {
  const int row = blockIdx.x;  // 0 .. 4H
  const int blk = row / H, n = row % H;
  for (int k = threadIdx.x; k < KA; k += blockDim.x) {
    float x = 0.f;
    if (k < H) {
      if (blk == 0) x = W_hv[(int64_t)n * (V + H) + V + k];
      else if (blk == 1) x = W_ha[(int64_t)n * (A + H) + A + k];
    } else if (k < H + V) {
      const int j = k - H;
      if (blk == 0) x = W_hv[(int64_t)n * (V + H) + j];
      else if (blk == 2) x = W_v[(int64_t)n * V + j];
    } else if (k < H + V + A) {
      const int j = k - H - V;
      if (blk == 1) x = W_ha[(int64_t)n * (A + H) + j];
      else if (blk == 3) x = W_a[(int64_t)n * A + j];
    }
    Bm[(int64_t)row * KA + k] = __float2half_rn(x);
  }
}
// inverse scatter of the packed weight gradient dB[4H][KA] (fp32) into the four Linear weight grads (+=)
__global__ void mag_unpack_wgrad_kernel(const float* __restrict__ dB, int V, int A, int KA, float* __restrict__ dW_hv,
                                        float* __restrict__ dW_ha, float* __restrict__ dW_v, float* __restrict__ dW_a) {
  const int row = blockIdx.x;
  const int blk = row / H, n = row % H;
  for (int k = threadIdx.x; k < H + V + A; k += blockDim.x) {
    const float x = dB[(int64_t)row * KA + k];
    if (k < H) {
      if (blk == 0) dW_hv[(int64_t)n * (V + H) + V + k] += x;
      else if (blk == 1) dW_ha[(int64_t)n * (A + H) + A + k] += x;
    } else if (k < H + V) {
      const int j = k - H;
      if (blk == 0) dW_hv[(int64_t)n * (V + H) + j] += x;
      else if (blk == 2) dW_v[(int64_t)n * V + j] += x;
    } else {
      const int j = k - H - V;
      if (blk == 1) dW_ha[(int64_t)n * (A + H) + j] += x;
      else if (blk == 3) dW_a[(int64_t)n * A + j] += x;
    }
  }
}

// --------------------------------------------------------------------------------------------
// Debug/test hook: the keep-mask (1/0 bytes) the kernels use for (seed, site) over n elements.
__global__ void dropout_mask_kernel(uint8_t* __restrict__ out, int64_t n, uint32_t thresh, Seed seed_in, uint32_t site) {
  const uint64_t seed = seed_in.ptr ? *seed_in.ptr : seed_in.value;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (n + 3) / 4; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 u = philox4x32(seed, site, (uint64_t)i);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    for (int k = 0; k < 4; ++k)
      if (4 * i + k < n) out[4 * i + k] = w[k] >= thresh ? 1 : 0;
  }
}

// ============================================================================================ host
static inline void drop_params(float p, uint32_t& thresh, float& scale) {
  if (p <= 0.f) { thresh = 0; scale = 1.f; return; }
  double t = (double)p * 4294967296.0;
  thresh = t >= 4294967295.0 ? 4294967295u : (uint32_t)t;
  scale = 1.0f / (1.0f - p);
}

int embed_ln_fwd(const EmbedArgs& a, cudaStream_t st) {
  MAGB_REQUIRE(a.hidden == H, "embed_ln: hidden must be 768");
  MAGB_REQUIRE(a.amag == nullptr || (a.V + a.A <= a.KA - H && a.KA % 8 == 0), "embed_ln: V+A=%d does not fit KA=%d", a.V + a.A, a.KA);
  uint32_t th; float sc;
  drop_params(a.p, th, sc);
  embed_ln_fwd_kernel<<<(a.T + 3) / 4, 128, 0, st>>>(a.ids, a.tt, a.visual, a.acoustic, a.word, a.pos, a.type, a.gamma,
                                                     a.beta, a.eps, th, sc, a.seed, a.site, a.V, a.A, a.KA, a.T, a.S,
                                                     a.vocab, a.emb32, a.amag);
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int mag_pack_inputs(const float* text, const float* visual, const float* acoustic, int V, int A, int KA, int T,
                    __half* amag, cudaStream_t st) {
  MAGB_REQUIRE(V + A <= KA - H, "mag: V+A=%d does not fit KA=%d", V + A, KA);
  mag_pack_inputs_kernel<<<(T + 3) / 4, 128, 0, st>>>(text, visual, acoustic, V, A, KA, T, amag);
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int mag_pack_split(const float* text, const float* visual, const float* acoustic, const float* W_hv, const float* W_ha,
                   const float* W_v, const float* W_a, int V, int A, int KA, int T, float* a_hi, float* a_lo, float* b_hi,
                   float* b_lo, cudaStream_t st) {
  mag_pack_inputs_split_kernel<<<T, 128, 0, st>>>(text, visual, acoustic, V, A, KA, T, a_hi, a_lo);
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  mag_pack_weights_split_kernel<<<4 * H, 128, 0, st>>>(W_hv, W_ha, W_v, W_a, V, A, KA, b_hi, b_lo);
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

static int grid_rows(int T) {
  int g = (T + 3) / 4;
  return g < 296 ? g : 296;  // 2 x 148 SMs; each warp then loops over rows
}
static int grid_rows_8w(int T) {
  int g = (T + 7) / 8;
  return g < 148 ? g : 148;  // 8 warps per block, one block per SM
}
static int grid_rows_1x(int T) {
  int g = (T + 3) / 4;
  return g < 148 ? g : 148;  // one block per SM: fewer global atomics at the end
}

int embed_ln_bwd(const EmbedArgs& a, const float* dout, float oscale, float* dword, float* dpos, float* dtype,
                 float* dgamma, float* dbeta, cudaStream_t st) {
  uint32_t th; float sc;
  drop_params(a.p, th, sc);
  const int B = a.T / a.S;
  MAGB_REQUIRE(B * a.S == a.T, "embed_ln_bwd: T=%d is not a multiple of S=%d", a.T, a.S);
  int zb = (128 + a.S - 1) / a.S;                 // ~128+ blocks in total
  const int zmax = (B + 7) / 8;
  zb = zb < 1 ? 1 : (zb > zmax ? zmax : zb);
  embed_ln_bwd_kernel<<<dim3(a.S, zb), 256, 0, st>>>(dout, a.ids, a.tt, a.word, a.pos, a.type, a.gamma, a.eps, th, sc,
                                                     a.seed, a.site, B, a.S, a.vocab, a.type_vocab, oscale, dword, dpos, dtype,
                                                     dgamma, dbeta);
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int mag_tail_fwd(const float* G, const float* text32, const float* gamma, const float* beta, float beta_shift, float eps,
                 float p, Seed seed, uint32_t site, int T, float* out32, __half* out16, cudaStream_t st) {
  uint32_t th; float sc;
  drop_params(p, th, sc);
  mag_tail_fwd_kernel<<<(T + 3) / 4, 128, 0, st>>>(G, text32, gamma, beta, beta_shift, eps, th, sc, seed, site, T, out32,
                                                   out16);
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int mag_tail_bwd(const float* dout, const float* G, const float* text32, const float* gamma, float beta_shift, float eps,
                 float p, Seed seed, uint32_t site, int T, float oscale, __half* dG, float* dtext, float* dgamma,
                 float* dbeta, float* dbias4, cudaStream_t st) {
  uint32_t th; float sc;
  drop_params(p, th, sc);
  mag_tail_bwd_kernel<<<grid_rows(T), 128, 0, st>>>(dout, G, text32, gamma, beta_shift, eps, th, sc, seed, site, T, oscale,
                                                    dG, dtext, dgamma, dbeta, dbias4);
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int dropout_residual_ln_fwd(const float* pre, const float* res, const float* gamma, const float* beta, float eps, float p,
                            Seed seed, uint32_t site, int T, float* y32, __half* y16, float* rstd,
                            cudaStream_t st, float* zero_after) {
  uint32_t th; float sc;
  drop_params(p, th, sc);
  MAGB_CHECK_CUDA(launch_pdl(dropout_residual_ln_fwd_kernel, dim3((T + 3) / 4), dim3(128), 0, st, pre, res, gamma, beta, eps,
                             th, sc, seed, site, T, y32, y16, rstd, zero_after));
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int ln_bwd(const float* dy32, const float* y32, const float* gamma, const float* beta, const float* rstd, float p,
           Seed seed, uint32_t site, int T, float oscale, float* dz32, __half* dpre, float* dgamma, float* dbeta,
           float* dbias, cudaStream_t st) {
  uint32_t th; float sc;
  drop_params(p, th, sc);
  constexpr size_t kSlots = (size_t)LNB_WARPS * 3 * H * sizeof(float);     // 147 456 B
  static bool configured[kMaxDevices] = {false};
  const int dev = current_device_slot();
  if (!configured[dev]) {
    MAGB_CHECK_CUDA(cudaFuncSetAttribute(ln_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSlots));
    configured[dev] = true;
  }
  int grid = (T + LNB_WARPS - 1) / LNB_WARPS;
  if (grid > 148) grid = 148;
  MAGB_CHECK_CUDA(launch_pdl(ln_bwd_kernel, dim3(grid), dim3(LNB_WARPS * 32), kSlots, st, dy32, y32, gamma, beta, rstd, th, sc, seed, site, T,
                             oscale, dz32, dpre, dgamma, dbeta, dbias));
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int colsum_f16(const __half* in, int T, int N, float oscale, float* out, cudaStream_t st) {
  MAGB_REQUIRE(N % 8 == 0, "colsum: N must be a multiple of 8");
  const int gx = (N + 255) / 256;
  int gy = T >= 512 ? (296 + gx - 1) / gx : 1;   // ~2 blocks per SM in total
  if (gy > (T + 31) / 32) gy = (T + 31) / 32;
  if (gy < 1) gy = 1;
  dim3 grid(gx, gy);
  MAGB_CHECK_CUDA(launch_pdl(colsum_f16_kernel, grid, dim3(256), 0, st, in, T, N, oscale, out));
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int cast_f32_f16(const float* in, __half* out, int64_t n, cudaStream_t st) {
  MAGB_REQUIRE(n % 4 == 0, "cast: n must be a multiple of 4");
  const int64_t n4 = n / 4;
  int blocks = (int)((n4 + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  cast_f32_f16_kernel<<<blocks, 256, 0, st>>>(in, out, n4);
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int mag_pack_weights(const float* W_hv, const float* W_ha, const float* W_v, const float* W_a, int V, int A, int KA,
                     __half* Bm, cudaStream_t st) {
  MAGB_REQUIRE(V + A <= KA - H, "mag: V+A=%d does not fit KA=%d", V + A, KA);
  mag_pack_weights_kernel<<<4 * H, 128, 0, st>>>(W_hv, W_ha, W_v, W_a, V, A, KA, Bm);
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int mag_unpack_wgrad(const float* dB, int V, int A, int KA, float* dW_hv, float* dW_ha, float* dW_v, float* dW_a,
                     cudaStream_t st) {
  mag_unpack_wgrad_kernel<<<4 * H, 128, 0, st>>>(dB, V, A, KA, dW_hv, dW_ha, dW_v, dW_a);
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int dropout_mask(uint8_t* out, int64_t n, float p, Seed seed, uint32_t site, cudaStream_t st) {
  uint32_t th; float sc;
  drop_params(p, th, sc);
  dropout_mask_kernel<<<256, 256, 0, st>>>(out, n, th, seed, site);
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace magb200

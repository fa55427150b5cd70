// Warp-MMA attention core for short sequences (S <= 64, head_dim 64): BertSelfAttention's
//     P = dropout(softmax(Q K^T / 8 + (1 - mask) * -10000)),  ctx = P V          
// and its backward, one CTA (4 warps) per (head, batch element), everything for that head resident in
// shared memory, scores and probabilities living in mma fragments (never materialised in memory).
//
// Why mma.sync (HMMA) and not tcgen05 here: at S = 50 the attention core is 1.1 % of the layer FLOPs
// and a (b, h) problem is 50x50x64 -- far below one 128-row UMMA tile; the kernel is latency bound, and
// the first CUDA-core version spent 54 us (fwd) / 110 us (bwd) per layer on shared-memory traffic.  The
// tcgen05 budget goes to the GEMMs; this kernel just has to get out of the way.
//
// Forward also emits lse[b][h][i] = max + log(sum) so backward can rebuild P without a cross-warp
// softmax reduction.  Backward works on the TRANSPOSED score tile (keys x queries): each warp owns 16
// keys and accumulates dV, dK for them over all queries (no cross-warp reduction), stores dS to shared
// memory, and after one barrier each warp produces dQ for its 16 queries.
#include <stdlib.h>
#include "common.cuh"
#include "host.h"

namespace magb200 {

namespace {

constexpr int D = 64;     // head dim
constexpr int SP = 64;    // padded sequence tile
constexpr int PITCH = 72; // halfs per smem row: 36 words -> fragment loads (row g, word t) hit 32 distinct banks
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
// softmax runs in the base-2 domain: scores and the additive mask are pre-multiplied by log2(e) so that every
// exponential is one ex2.approx (MUFU) instead of expf's range reduction (ncu: these kernels are issue-bound)

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t ld32(const __half* p) { return *reinterpret_cast<const uint32_t*>(p); }
__device__ __forceinline__ uint32_t packh(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// [S][64] slice (row pitch ld halfs) of a global fp16 matrix -> smem [64][PITCH], rows >= S zeroed
__device__ __forceinline__ void load_tile(const __half* __restrict__ g, int ld, int S, __half* s) {
  for (int i = threadIdx.x; i < SP * 8; i += blockDim.x) {
    const int r = i >> 3, c = (i & 7) * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < S) v = *reinterpret_cast<const uint4*>(g + (int64_t)r * ld + c);
    *reinterpret_cast<uint4*>(s + r * PITCH + c) = v;
  }
}
// same, but stored transposed: s[c][r].  Consecutive lanes take consecutive source rows, so the 2-byte
// transposed stores of a warp fall into consecutive halfs of one smem row (no bank conflicts).
__device__ __forceinline__ void load_tile_t(const __half* __restrict__ g, int ld, int S, __half* s) {
  for (int i = threadIdx.x; i < SP * 8; i += blockDim.x) {
    const int r = i & (SP - 1), c = (i >> 6) * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < S) v = *reinterpret_cast<const uint4*>(g + (int64_t)r * ld + c);
    const __half* h = reinterpret_cast<const __half*>(&v);
#pragma unroll
    for (int k = 0; k < 8; ++k) s[(c + k) * PITCH + r] = h[k];
  }
}

// Asynchronous variant (cp.async, 16 B per request, rows >= S zero-filled): all tiles of a CTA are requested
// back to back and awaited once, instead of paying one global-load latency per tile (the ncu source view showed
// ~40 % of the backward kernel's samples stalled on the first shared-memory store after each tile's loads).
__device__ __forceinline__ void load_tile_async(const __half* __restrict__ g, int ld, int S, __half* s) {
  for (int i = threadIdx.x; i < SP * 8; i += blockDim.x) {
    const int r = i >> 3, c = (i & 7) * 8;
    const bool ok = r < S;
    const __half* src = g + (int64_t)(ok ? r : 0) * ld + c;
    const uint32_t n = ok ? 16u : 0u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(s + r * PITCH + c)), "l"(src), "r"(n)
                 : "memory");
  }
}
__device__ __forceinline__ void async_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}
// dst[c][r] = src[r][c] for a [64][64] tile already in shared memory (consecutive lanes -> consecutive r)
__device__ __forceinline__ void transpose_tile(const __half* src, __half* dst) {
  for (int i = threadIdx.x; i < SP * 8; i += blockDim.x) {
    const int r = i & (SP - 1), c = (i >> 6) * 8;
    const uint4 v = *reinterpret_cast<const uint4*>(src + r * PITCH + c);
    const __half* h = reinterpret_cast<const __half*>(&v);
#pragma unroll
    for (int k = 0; k < 8; ++k) dst[(c + k) * PITCH + r] = h[k];
  }
}

// A fragments (16 rows starting at r0, 16 k-columns starting at k0) of a row-major smem tile
__device__ __forceinline__ void frag_a(const __half* s, int r0, int k0, int g, int t, uint32_t (&a)[4]) {
  a[0] = ld32(s + (r0 + g) * PITCH + k0 + 2 * t);
  a[1] = ld32(s + (r0 + g + 8) * PITCH + k0 + 2 * t);
  a[2] = ld32(s + (r0 + g) * PITCH + k0 + 2 * t + 8);
  a[3] = ld32(s + (r0 + g + 8) * PITCH + k0 + 2 * t + 8);
}
// B fragments for B[k][n] = s[n0 + n][k0 + k]  (an "N x K" row-major smem tile)
__device__ __forceinline__ void frag_b(const __half* s, int n0, int k0, int g, int t, uint32_t (&b)[2]) {
  b[0] = ld32(s + (n0 + g) * PITCH + k0 + 2 * t);
  b[1] = ld32(s + (n0 + g) * PITCH + k0 + 2 * t + 8);
}

// B fragments for B[k][n] = s[k0 + k][n0 + n] (a [k][n] row-major tile: V as [key][d]) for the two n-tiles n0 and n0 + 8,
// read with ldmatrix.trans straight from the row-major tile (r[0], r[1]: n-tile n0; r[2], r[3]: n-tile n0 + 8) -- no
// transposed copy of the tile in shared memory
__device__ __forceinline__ void frag_b_trans2(const __half* s, int k0, int n0, int lane, uint32_t (&r)[4]) {
  const int m = lane >> 3, rr = lane & 7;
  const __half* p = s + (k0 + (m & 1) * 8 + rr) * PITCH + n0 + (m >> 1) * 8;
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}

__device__ __forceinline__ float mask_add(const int64_t* mrow, int j) {
  return mrow == nullptr ? 0.f : (1.0f - (float)mrow[j]) * -10000.0f;   // bert.py:180-182 (3.0.2 arithmetic)
}
// keep * 1/(1-p) (or 0) for P[b][h][i][j]; the counter is the flattened index / 4, as in attention.cu
__device__ __forceinline__ float drop_scale(uint64_t seed, uint32_t site, uint64_t idx, uint32_t thresh, float dscale) {
  if (thresh == 0) return 1.f;
  const uint4 u = philox4x32(seed, site, idx >> 2);
  const uint32_t w = (idx & 3) == 0 ? u.x : (idx & 3) == 1 ? u.y : (idx & 3) == 2 ? u.z : u.w;
  return w >= thresh ? dscale : 0.f;
}

// Keep-flags of one (batch, head) probability tile, generated cooperatively: the S*S elements are a contiguous
// index range [base, base + S*S), i.e. ~S*S/4 Philox counters for the whole CTA (the per-element form costs every
// thread one full Philox-4x32-10 per probability it holds: 32 calls, more instructions than the five tile matmuls
// together).  One bit per element, bit (q * S + key); a thread owns whole 32-bit words so no atomics are needed.
__device__ __forceinline__ void fill_keep_bits(uint32_t* keep, uint64_t seed, uint32_t site, uint64_t base, int n,
                                               uint32_t thresh) {
  const int words = (n + 31) >> 5;
  for (int wd = threadIdx.x; wd < words; wd += blockDim.x) {
    const uint64_t first = base + 32ull * wd;
    const uint64_t last = first + 31;
    uint32_t bits = 0;
    for (uint64_t c = first >> 2; c <= (last >> 2); ++c) {
      const uint4 u = philox4x32(seed, site, c);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int64_t loc = (int64_t)(4 * c + k) - (int64_t)first;     // bit position inside this word
        if (loc >= 0 && loc < 32 && w[k] >= thresh) bits |= 1u << loc;
      }
    }
    keep[wd] = bits;
  }
}
__device__ __forceinline__ bool keep_bit(const uint32_t* keep, int idx) { return (keep[idx >> 5] >> (idx & 31)) & 1u; }

// the same for two consecutive elements idx, idx+1 (one Philox call when they share a counter)
__device__ __forceinline__ void drop_scale2(uint64_t seed, uint32_t site, uint64_t idx, uint32_t thresh, float dscale,
                                            float& d0, float& d1) {
  const uint4 u = philox4x32(seed, site, idx >> 2);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
  const int c = (int)(idx & 3);
  d0 = w[c] >= thresh ? dscale : 0.f;
  if (c < 3) d1 = w[c + 1] >= thresh ? dscale : 0.f;
  else d1 = philox4x32(seed, site, (idx + 1) >> 2).x >= thresh ? dscale : 0.f;
}

// ------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(128) sdpa_fwd_mma_kernel(const __half* __restrict__ qkv,
                                                           const int64_t* __restrict__ mask, int S, int heads,
                                                           uint32_t thresh, float dscale, Seed seed_in, uint32_t site,
                                                           __half* __restrict__ ctx, float* __restrict__ lse) {
  pdl_wait();      // may be launched with the PDL attribute: the predecessor can still be running
  pdl_trigger();   // the next kernel in the stream (a GEMM) may start its private prologue
  const uint64_t seed = seed_in.ptr ? *seed_in.ptr : seed_in.value;
  __shared__ __align__(16) __half sQ[SP * PITCH];
  __shared__ __align__(16) __half sK[SP * PITCH];
  __shared__ __align__(16) __half sV[SP * PITCH];
  __shared__ uint32_t sKeep[SP * SP / 32];
  __shared__ float sMask[SP];        // additive key mask * log2(e); -inf beyond S
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int Hd = heads * D, ld = 3 * Hd;
  const __half* base = qkv + (int64_t)b * S * ld + h * D;
  if (threadIdx.x < SP) {
    const int j = threadIdx.x;
    sMask[j] = j < S ? mask_add(mask ? mask + (int64_t)b * S : nullptr, j) * LOG2E : -INFINITY;
  }
  load_tile_async(base, ld, S, sQ);
  load_tile_async(base + Hd, ld, S, sK);
  load_tile_async(base + 2 * Hd, ld, S, sV);
  if (thresh != 0) fill_keep_bits(sKeep, seed, site, ((uint64_t)b * heads + h) * S * S, S * S, thresh);
  async_wait_all();
  __syncthreads();
  const int r0 = warp * 16;
  if (r0 >= S) return;

  float sc[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t a[4];
    frag_a(sQ, r0, kk * 16, g, t, a);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      uint32_t bb[2];
      frag_b(sK, nt * 8, kk * 16, g, t, bb);
      mma16816(sc[nt], a[0], a[1], a[2], a[3], bb[0], bb[1]);
    }
  }
  // softmax over keys for rows r0+g (elements [.][0..1]) and r0+g+8 (elements [.][2..3])
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float m = sMask[nt * 8 + 2 * t + e];
      sc[nt][e] = fmaf(sc[nt][e], 0.125f * LOG2E, m);
      sc[nt][2 + e] = fmaf(sc[nt][2 + e], 0.125f * LOG2E, m);
      mx0 = fmaxf(mx0, sc[nt][e]);
      mx1 = fmaxf(mx1, sc[nt][2 + e]);
    }
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      sc[nt][e] = ex2_approx(sc[nt][e] - mx0);          // 2^(-inf) = 0 for padded keys
      sc[nt][2 + e] = ex2_approx(sc[nt][2 + e] - mx1);
      sum0 += sc[nt][e];
      sum1 += sc[nt][2 + e];
    }
  }
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
  const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
  const int i0 = r0 + g, i1 = r0 + g + 8;
  if (lse != nullptr && t == 0) {
    if (i0 < S) lse[((int64_t)b * heads + h) * S + i0] = mx0 * LN2 + logf(sum0);      // natural-log lse (the ABI's unit)
    if (i1 < S) lse[((int64_t)b * heads + h) * S + i1] = mx1 * LN2 + logf(sum1);
  }
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int j = nt * 8 + 2 * t;
    float d00 = 1.f, d01 = 1.f, d10 = 1.f, d11 = 1.f;
    if (thresh != 0) {
      d00 = keep_bit(sKeep, i0 * S + j) ? dscale : 0.f;
      d01 = keep_bit(sKeep, i0 * S + j + 1) ? dscale : 0.f;
      d10 = keep_bit(sKeep, i1 * S + j) ? dscale : 0.f;
      d11 = keep_bit(sKeep, i1 * S + j + 1) ? dscale : 0.f;
    }
    sc[nt][0] *= inv0 * d00; sc[nt][1] *= inv0 * d01;
    sc[nt][2] *= inv1 * d10; sc[nt][3] *= inv1 * d11;
  }
  // O = P V
  float o[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const uint32_t a0 = packh(sc[2 * kk][0], sc[2 * kk][1]), a1 = packh(sc[2 * kk][2], sc[2 * kk][3]);
    const uint32_t a2 = packh(sc[2 * kk + 1][0], sc[2 * kk + 1][1]), a3 = packh(sc[2 * kk + 1][2], sc[2 * kk + 1][3]);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t bb[4];
      frag_b_trans2(sV, kk * 16, np * 16, lane, bb);      // B[k = key][n = d] = V[key][d]
      mma16816(o[2 * np], a0, a1, a2, a3, bb[0], bb[1]);
      mma16816(o[2 * np + 1], a0, a1, a2, a3, bb[2], bb[3]);
    }
  }
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int d = nt * 8 + 2 * t;
    if (i0 < S) *reinterpret_cast<uint32_t*>(ctx + ((int64_t)b * S + i0) * Hd + h * D + d) = packh(o[nt][0], o[nt][1]);
    if (i1 < S) *reinterpret_cast<uint32_t*>(ctx + ((int64_t)b * S + i1) * Hd + h * D + d) = packh(o[nt][2], o[nt][3]);
  }
}

// ------------------------------------------------------------------------------------------ backward
struct BwdSmem {
  __half q[SP * PITCH], k[SP * PITCH], v[SP * PITCH], dO[SP * PITCH];
  __half kt[D * PITCH], qt[D * PITCH], dOt[D * PITCH], dS[SP * PITCH];
  float Dq[SP];        // rowsum(dO * O) per query
  float lse[SP];
  float bias[3 * D];   // column sums of dQ | dK | dV for this head
  uint32_t keep[SP * SP / 32];   // dropout keep bits of the probability tile, bit (q * S + key)
};

// column sums (over the 16 rows of this warp's tile) of a C-layout accumulator -> smem accumulator[64]
__device__ __forceinline__ void colsum_tile(const float (&acc)[8][4], float* sacc, int g, int t, bool row0_ok, bool row1_ok) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    float s0 = (row0_ok ? acc[nt][0] : 0.f) + (row1_ok ? acc[nt][2] : 0.f);
    float s1 = (row0_ok ? acc[nt][1] : 0.f) + (row1_ok ? acc[nt][3] : 0.f);
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if (g == 0) {
      atomicAdd(sacc + nt * 8 + 2 * t, s0);
      atomicAdd(sacc + nt * 8 + 2 * t + 1, s1);
    }
  }
}

__global__ void __launch_bounds__(128) sdpa_bwd_mma_kernel(const __half* __restrict__ qkv, const __half* __restrict__ dctx,
                                                           const __half* __restrict__ ctx, const float* __restrict__ lse,
                                                           const int64_t* __restrict__ mask, int S, int heads,
                                                           uint32_t thresh, float dscale, Seed seed_in, uint32_t site,
                                                           float oscale, __half* __restrict__ dqkv,
                                                           float* __restrict__ dbias) {
  pdl_trigger();   // the next kernel in the stream (a GEMM) may start its private prologue
  const uint64_t seed = seed_in.ptr ? *seed_in.ptr : seed_in.value;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  BwdSmem& sm = *reinterpret_cast<BwdSmem*>(smem_raw);
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int Hd = heads * D, ld = 3 * Hd;
  const __half* base = qkv + (int64_t)b * S * ld + h * D;
  const __half* dObase = dctx + (int64_t)b * S * Hd + h * D;
  load_tile_async(base, ld, S, sm.q);
  load_tile_async(base + Hd, ld, S, sm.k);
  load_tile_async(base + 2 * Hd, ld, S, sm.v);
  load_tile_async(dObase, Hd, S, sm.dO);
  for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) sm.bias[i] = 0.f;
  if (thresh != 0) fill_keep_bits(sm.keep, seed, site, ((uint64_t)b * heads + h) * S * S, S * S, thresh);
  {
    // D[q] = sum_d dO[q][d] * O[q][d]: two threads per query (32 head dims each), 16-byte loads
    const int i = threadIdx.x >> 1, half = threadIdx.x & 1;
    float acc = 0.f;
    if (i < S) {
      const __half* o = ctx + ((int64_t)b * S + i) * Hd + h * D + half * 32;
      const __half* d = dObase + (int64_t)i * Hd + half * 32;
      uint4 ov[4], dv[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        ov[c] = *reinterpret_cast<const uint4*>(o + c * 8);
        dv[c] = *reinterpret_cast<const uint4*>(d + c * 8);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const __half2* oh = reinterpret_cast<const __half2*>(&ov[c]);
        const __half2* dh = reinterpret_cast<const __half2*>(&dv[c]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 a = __half22float2(oh[k]), e = __half22float2(dh[k]);
          acc += a.x * e.x + a.y * e.y;
        }
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (half == 0) {
      sm.Dq[i] = acc;
      sm.lse[i] = i < S ? lse[((int64_t)b * heads + h) * S + i] : 0.f;
    }
  }
  async_wait_all();
  __syncthreads();
  transpose_tile(sm.q, sm.qt);
  transpose_tile(sm.k, sm.kt);
  transpose_tile(sm.dO, sm.dOt);
  __syncthreads();
  const int64_t* mrow = mask ? mask + (int64_t)b * S : nullptr;
  const int r0 = warp * 16;
  const int k0 = r0 + g, k1 = r0 + g + 8;            // this thread's two key rows
  const uint64_t bh = (uint64_t)b * heads + h;
  __half* obase = dqkv + (int64_t)b * S * ld + h * D;

  if (r0 < S) {
    // ---- S^T = K Q^T (keys x queries) and dPd^T = V dO^T
    float st[8][4], dp[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      st[nt][0] = st[nt][1] = st[nt][2] = st[nt][3] = 0.f;
      dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t ak[4], av[4];
      frag_a(sm.k, r0, kk * 16, g, t, ak);
      frag_a(sm.v, r0, kk * 16, g, t, av);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        uint32_t bq[2], bo[2];
        frag_b(sm.q, nt * 8, kk * 16, g, t, bq);       // B[k = d][n = query] = Q[query][d]
        frag_b(sm.dO, nt * 8, kk * 16, g, t, bo);
        mma16816(st[nt], ak[0], ak[1], ak[2], ak[3], bq[0], bq[1]);
        mma16816(dp[nt], av[0], av[1], av[2], av[3], bo[0], bo[1]);
      }
    }
    // ---- P^T, Pd^T, dS^T in place
    const float m0 = k0 < S ? mask_add(mrow, k0) : 0.f, m1 = k1 < S ? mask_add(mrow, k1) : 0.f;
    float pd[8][4];                                  // dropped probabilities (operand of dV)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = (e < 2) ? k0 : k1;
        const int qi = nt * 8 + 2 * t + (e & 1);
        float p = 0.f, dsv = 0.f, pdv = 0.f;
        if (key < S && qi < S) {
          p = expf(st[nt][e] * 0.125f + ((e < 2) ? m0 : m1) - sm.lse[qi]);
          const float ds = thresh == 0 ? 1.f : (keep_bit(sm.keep, qi * S + key) ? dscale : 0.f);
          pdv = p * ds;
          dsv = p * (dp[nt][e] * ds - sm.Dq[qi]) * 0.125f;      // includes d(score)/d(qk) = 1/sqrt(64)
        }
        pd[nt][e] = pdv;
        st[nt][e] = dsv;                              // st now holds dS^T
      }
      // dS[query][key] for the dQ phase
      const int q0 = nt * 8 + 2 * t;
      sm.dS[q0 * PITCH + k0] = __float2half_rn(st[nt][0]);
      sm.dS[(q0 + 1) * PITCH + k0] = __float2half_rn(st[nt][1]);
      sm.dS[q0 * PITCH + k1] = __float2half_rn(st[nt][2]);
      sm.dS[(q0 + 1) * PITCH + k1] = __float2half_rn(st[nt][3]);
    }
    // ---- dV = Pd^T dO,  dK = dS^T Q   (reduction over queries)
    float dv[8][4], dk[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      dv[nt][0] = dv[nt][1] = dv[nt][2] = dv[nt][3] = 0.f;
      dk[nt][0] = dk[nt][1] = dk[nt][2] = dk[nt][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const uint32_t p0 = packh(pd[2 * kk][0], pd[2 * kk][1]), p1 = packh(pd[2 * kk][2], pd[2 * kk][3]);
      const uint32_t p2 = packh(pd[2 * kk + 1][0], pd[2 * kk + 1][1]), p3 = packh(pd[2 * kk + 1][2], pd[2 * kk + 1][3]);
      const uint32_t s0 = packh(st[2 * kk][0], st[2 * kk][1]), s1 = packh(st[2 * kk][2], st[2 * kk][3]);
      const uint32_t s2 = packh(st[2 * kk + 1][0], st[2 * kk + 1][1]), s3 = packh(st[2 * kk + 1][2], st[2 * kk + 1][3]);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        uint32_t bo[2], bq[2];
        frag_b(sm.dOt, nt * 8, kk * 16, g, t, bo);     // B[k = query][n = d] = dOt[d][query]
        frag_b(sm.qt, nt * 8, kk * 16, g, t, bq);
        mma16816(dv[nt], p0, p1, p2, p3, bo[0], bo[1]);
        mma16816(dk[nt], s0, s1, s2, s3, bq[0], bq[1]);
      }
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int d = nt * 8 + 2 * t;
      if (k0 < S) {
        *reinterpret_cast<uint32_t*>(obase + (int64_t)k0 * ld + Hd + d) = packh(dk[nt][0], dk[nt][1]);
        *reinterpret_cast<uint32_t*>(obase + (int64_t)k0 * ld + 2 * Hd + d) = packh(dv[nt][0], dv[nt][1]);
      }
      if (k1 < S) {
        *reinterpret_cast<uint32_t*>(obase + (int64_t)k1 * ld + Hd + d) = packh(dk[nt][2], dk[nt][3]);
        *reinterpret_cast<uint32_t*>(obase + (int64_t)k1 * ld + 2 * Hd + d) = packh(dv[nt][2], dv[nt][3]);
      }
    }
    if (dbias != nullptr) {
      colsum_tile(dk, sm.bias + D, g, t, k0 < S, k1 < S);
      colsum_tile(dv, sm.bias + 2 * D, g, t, k0 < S, k1 < S);
    }
  } else {
    // keys >= S: dS rows of this warp's key range are zero
    for (int i = lane; i < SP * 16; i += 32) sm.dS[(i >> 4) * PITCH + r0 + (i & 15)] = __float2half_rn(0.f);
  }
  __syncthreads();
  // ---- dQ = dS K for query rows r0 .. r0+15
  if (r0 < S) {
    float dq[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) dq[nt][0] = dq[nt][1] = dq[nt][2] = dq[nt][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a[4];
      frag_a(sm.dS, r0, kk * 16, g, t, a);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        uint32_t bb[2];
        frag_b(sm.kt, nt * 8, kk * 16, g, t, bb);      // B[k = key][n = d] = Kt[d][key]
        mma16816(dq[nt], a[0], a[1], a[2], a[3], bb[0], bb[1]);
      }
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int d = nt * 8 + 2 * t;
      if (k0 < S) *reinterpret_cast<uint32_t*>(obase + (int64_t)k0 * ld + d) = packh(dq[nt][0], dq[nt][1]);
      if (k1 < S) *reinterpret_cast<uint32_t*>(obase + (int64_t)k1 * ld + d) = packh(dq[nt][2], dq[nt][3]);
    }
    if (dbias != nullptr) colsum_tile(dq, sm.bias, g, t, k0 < S, k1 < S);
  }
  if (dbias != nullptr) {
    __syncthreads();
    // bias gradients of the fused query|key|value projection: dbias[{0,Hd,2Hd} + h*64 + d]
    for (int i = threadIdx.x; i < 3 * D; i += blockDim.x)
      atomicAdd(dbias + (i / D) * Hd + h * D + (i % D), sm.bias[i] * oscale);
  }
}

// ---------------------------------------------------------------------------------------------
// Backward with 8 warps per (head, batch element).  The 4-warp kernel above is latency bound (ncu: issue active
// 30 %, ~3 warps per scheduler): here warp w = (key block kb = w & 3, half hq = w >> 2) owns a quarter of the work
//   phase 1: S^T, dP^T for 16 keys x 32 queries (query half hq)  -> Pd^T, dS^T (fp16, [key][query]) and dS ([query][key])
//            go to shared memory; Pd^T / dS^T reuse the V / dO tiles, which are dead once every warp has left phase 1
//   phase 2: dV, dK for 16 keys x 32 head dims (dim half hq), reduction over all 64 queries from shared memory
//   phase 3: dQ for 16 queries (block kb) x 32 head dims
// shared memory of the 8-warp kernel: transposed operands are read with ldmatrix.trans from the row-major tiles, so no
// transposed copies exist; Pd^T / dS^T ([key][query]) and dS ([query][key]) have their own tiles
struct Bwd8Smem {
  __half q[SP * PITCH], k[SP * PITCH], v[SP * PITCH], dO[SP * PITCH];
  __half pdT[SP * PITCH], dsT[SP * PITCH], dS[SP * PITCH];
  float Dq[SP];
  float lse[SP];
  float bias[3 * D];
  uint32_t keep[SP * SP / 32];
};

template <int NT>
__device__ __forceinline__ void colsum_part(const float (&acc)[NT][4], float* sacc, int g, int t, bool row0_ok, bool row1_ok) {
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    float s0 = (row0_ok ? acc[nt][0] : 0.f) + (row1_ok ? acc[nt][2] : 0.f);
    float s1 = (row0_ok ? acc[nt][1] : 0.f) + (row1_ok ? acc[nt][3] : 0.f);
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if (g == 0) {
      atomicAdd(sacc + nt * 8 + 2 * t, s0);
      atomicAdd(sacc + nt * 8 + 2 * t + 1, s1);
    }
  }
}

__global__ void __launch_bounds__(256) sdpa_bwd_mma8_kernel(const __half* __restrict__ qkv, const __half* __restrict__ dctx,
                                                            const __half* __restrict__ ctx, const float* __restrict__ lse,
                                                            const int64_t* __restrict__ mask, int S, int heads,
                                                            uint32_t thresh, float dscale, Seed seed_in, uint32_t site,
                                                            float oscale, __half* __restrict__ dqkv,
                                                            float* __restrict__ dbias) {
  pdl_wait();      // may be launched with the PDL attribute: the predecessor can still be running
  pdl_trigger();   // the next kernel in the stream (a GEMM) may start its private prologue
  const uint64_t seed = seed_in.ptr ? *seed_in.ptr : seed_in.value;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Bwd8Smem& sm = *reinterpret_cast<Bwd8Smem*>(smem_raw);
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int kb = warp & 3, hq = warp >> 2;
  const int Hd = heads * D, ld = 3 * Hd;
  const __half* base = qkv + (int64_t)b * S * ld + h * D;
  const __half* dObase = dctx + (int64_t)b * S * Hd + h * D;
  load_tile_async(base, ld, S, sm.q);
  load_tile_async(base + Hd, ld, S, sm.k);
  load_tile_async(base + 2 * Hd, ld, S, sm.v);
  load_tile_async(dObase, Hd, S, sm.dO);
  for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) sm.bias[i] = 0.f;
  if (thresh != 0) fill_keep_bits(sm.keep, seed, site, ((uint64_t)b * heads + h) * S * S, S * S, thresh);
  if (threadIdx.x < 2 * SP) {
    // D[q] = sum_d dO[q][d] * O[q][d]: two threads per query (32 head dims each), 16-byte loads
    const int i = threadIdx.x >> 1, half = threadIdx.x & 1;
    float acc = 0.f;
    if (i < S) {
      const __half* o = ctx + ((int64_t)b * S + i) * Hd + h * D + half * 32;
      const __half* d = dObase + (int64_t)i * Hd + half * 32;
      uint4 ov[4], dv[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        ov[c] = *reinterpret_cast<const uint4*>(o + c * 8);
        dv[c] = *reinterpret_cast<const uint4*>(d + c * 8);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const __half2* oh = reinterpret_cast<const __half2*>(&ov[c]);
        const __half2* dh = reinterpret_cast<const __half2*>(&dv[c]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 a = __half22float2(oh[k]), e = __half22float2(dh[k]);
          acc += a.x * e.x + a.y * e.y;
        }
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (half == 0) {
      sm.Dq[i] = acc;
      sm.lse[i] = i < S ? lse[((int64_t)b * heads + h) * S + i] * LOG2E : 0.f;          // base-2 domain
    }
  }
  async_wait_all();
  __syncthreads();
  const int64_t* mrow = mask ? mask + (int64_t)b * S : nullptr;
  const int r0 = kb * 16;
  const int k0 = r0 + g, k1 = r0 + g + 8;            // this thread's two key rows (phase 1/2) or query rows (phase 3)
  const int nb = hq * 4;                             // first 8-wide n-tile of this warp's half
  __half* obase = dqkv + (int64_t)b * S * ld + h * D;
  const bool active = r0 < S;

  // ---- phase 1: S^T = K Q^T and dPd^T = V dO^T for (16 keys) x (32 queries)
  float st[4][4], pd[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    st[j][0] = st[j][1] = st[j][2] = st[j][3] = 0.f;
    pd[j][0] = pd[j][1] = pd[j][2] = pd[j][3] = 0.f;
  }
  if (active) {
    float dp[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t ak[4], av[4];
      frag_a(sm.k, r0, kk * 16, g, t, ak);
      frag_a(sm.v, r0, kk * 16, g, t, av);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t bq[2], bo[2];
        frag_b(sm.q, (nb + j) * 8, kk * 16, g, t, bq);       // B[k = d][n = query] = Q[query][d]
        frag_b(sm.dO, (nb + j) * 8, kk * 16, g, t, bo);
        mma16816(st[j], ak[0], ak[1], ak[2], ak[3], bq[0], bq[1]);
        mma16816(dp[j], av[0], av[1], av[2], av[3], bo[0], bo[1]);
      }
    }
    const float m0 = k0 < S ? mask_add(mrow, k0) * LOG2E : 0.f, m1 = k1 < S ? mask_add(mrow, k1) * LOG2E : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = (e < 2) ? k0 : k1;
        const int qi = (nb + j) * 8 + 2 * t + (e & 1);
        float dsv = 0.f, pdv = 0.f;
        if (key < S && qi < S) {
          const float p = ex2_approx(fmaf(st[j][e], 0.125f * LOG2E, ((e < 2) ? m0 : m1) - sm.lse[qi]));
          const float ds = thresh == 0 ? 1.f : (keep_bit(sm.keep, qi * S + key) ? dscale : 0.f);
          pdv = p * ds;
          dsv = p * (dp[j][e] * ds - sm.Dq[qi]) * 0.125f;      // includes d(score)/d(qk) = 1/sqrt(64)
        }
        pd[j][e] = pdv;
        st[j][e] = dsv;                                          // st now holds dS^T
      }
    }
  }
  __half* pdT = sm.pdT;                              // [key][query]
  __half* dsT = sm.dsT;                              // [key][query]
  if (active) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int q0 = (nb + j) * 8 + 2 * t;
      *reinterpret_cast<uint32_t*>(pdT + k0 * PITCH + q0) = packh(pd[j][0], pd[j][1]);
      *reinterpret_cast<uint32_t*>(pdT + k1 * PITCH + q0) = packh(pd[j][2], pd[j][3]);
      *reinterpret_cast<uint32_t*>(dsT + k0 * PITCH + q0) = packh(st[j][0], st[j][1]);
      *reinterpret_cast<uint32_t*>(dsT + k1 * PITCH + q0) = packh(st[j][2], st[j][3]);
      sm.dS[q0 * PITCH + k0] = __float2half_rn(st[j][0]);
      sm.dS[(q0 + 1) * PITCH + k0] = __float2half_rn(st[j][1]);
      sm.dS[q0 * PITCH + k1] = __float2half_rn(st[j][2]);
      sm.dS[(q0 + 1) * PITCH + k1] = __float2half_rn(st[j][3]);
    }
  } else {
    // keys >= S: the dS columns of this key block are zero (for this warp's 32 queries)
    for (int i = lane; i < 32 * 16; i += 32) sm.dS[(hq * 32 + (i >> 4)) * PITCH + r0 + (i & 15)] = __float2half_rn(0.f);
  }
  __syncthreads();

  if (active) {
    // ---- phase 2: dV = Pd^T dO, dK = dS^T Q for (16 keys) x (32 head dims), reduction over the 64 queries
    float dv[4][4], dk[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.f;
      dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t ap[4], as[4];
      frag_a(pdT, r0, kk * 16, g, t, ap);
      frag_a(dsT, r0, kk * 16, g, t, as);
#pragma unroll
      for (int jp = 0; jp < 2; ++jp) {
        uint32_t bo[4], bq[4];
        frag_b_trans2(sm.dO, kk * 16, (nb + 2 * jp) * 8, lane, bo);     // B[k = query][n = d] = dO[query][d]
        frag_b_trans2(sm.q, kk * 16, (nb + 2 * jp) * 8, lane, bq);
        mma16816(dv[2 * jp], ap[0], ap[1], ap[2], ap[3], bo[0], bo[1]);
        mma16816(dv[2 * jp + 1], ap[0], ap[1], ap[2], ap[3], bo[2], bo[3]);
        mma16816(dk[2 * jp], as[0], as[1], as[2], as[3], bq[0], bq[1]);
        mma16816(dk[2 * jp + 1], as[0], as[1], as[2], as[3], bq[2], bq[3]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int d = (nb + j) * 8 + 2 * t;
      if (k0 < S) {
        *reinterpret_cast<uint32_t*>(obase + (int64_t)k0 * ld + Hd + d) = packh(dk[j][0], dk[j][1]);
        *reinterpret_cast<uint32_t*>(obase + (int64_t)k0 * ld + 2 * Hd + d) = packh(dv[j][0], dv[j][1]);
      }
      if (k1 < S) {
        *reinterpret_cast<uint32_t*>(obase + (int64_t)k1 * ld + Hd + d) = packh(dk[j][2], dk[j][3]);
        *reinterpret_cast<uint32_t*>(obase + (int64_t)k1 * ld + 2 * Hd + d) = packh(dv[j][2], dv[j][3]);
      }
    }
    if (dbias != nullptr) {
      colsum_part<4>(dk, sm.bias + D + nb * 8, g, t, k0 < S, k1 < S);
      colsum_part<4>(dv, sm.bias + 2 * D + nb * 8, g, t, k0 < S, k1 < S);
    }
    // ---- phase 3: dQ = dS K for (16 queries) x (32 head dims)
    float dq[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a[4];
      frag_a(sm.dS, r0, kk * 16, g, t, a);
#pragma unroll
      for (int jp = 0; jp < 2; ++jp) {
        uint32_t bb[4];
        frag_b_trans2(sm.k, kk * 16, (nb + 2 * jp) * 8, lane, bb);      // B[k = key][n = d] = K[key][d]
        mma16816(dq[2 * jp], a[0], a[1], a[2], a[3], bb[0], bb[1]);
        mma16816(dq[2 * jp + 1], a[0], a[1], a[2], a[3], bb[2], bb[3]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int d = (nb + j) * 8 + 2 * t;
      if (k0 < S) *reinterpret_cast<uint32_t*>(obase + (int64_t)k0 * ld + d) = packh(dq[j][0], dq[j][1]);
      if (k1 < S) *reinterpret_cast<uint32_t*>(obase + (int64_t)k1 * ld + d) = packh(dq[j][2], dq[j][3]);
    }
    if (dbias != nullptr) colsum_part<4>(dq, sm.bias + nb * 8, g, t, k0 < S, k1 < S);
  }
  if (dbias != nullptr) {
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * D; i += blockDim.x)
      atomicAdd(dbias + (i / D) * Hd + h * D + (i % D), sm.bias[i] * oscale);
  }
}

inline void drop_params(float p, uint32_t& thresh, float& scale) {
  if (p <= 0.f) { thresh = 0; scale = 1.f; return; }
  double t = (double)p * 4294967296.0;
  thresh = t >= 4294967295.0 ? 4294967295u : (uint32_t)t;
  scale = 1.0f / (1.0f - p);
}

}  // namespace

int sdpa_fwd_mma(const __half* qkv, const int64_t* mask, int B, int S, int heads, float p, Seed seed, uint32_t site,
                 __half* ctx, float* lse, cudaStream_t st) {
  MAGB_REQUIRE(S >= 1 && S <= SP, "sdpa_fwd_mma: S=%d out of range", S);
  uint32_t th; float ds;
  drop_params(p, th, ds);
  MAGB_CHECK_CUDA(launch_pdl(sdpa_fwd_mma_kernel, dim3(heads, B), dim3(128), 0, st, qkv, mask, S, heads, th, ds, seed, site, ctx,
                             lse));
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int sdpa_bwd_mma(const __half* qkv, const __half* dctx, const __half* ctx, const float* lse, const int64_t* mask, int B,
                 int S, int heads, float p, Seed seed, uint32_t site, float oscale, __half* dqkv, float* dbias,
                 cudaStream_t st) {
  MAGB_REQUIRE(S >= 1 && S <= SP, "sdpa_bwd_mma: S=%d out of range", S);
  uint32_t th; float ds;
  drop_params(p, th, ds);
  static bool configured_by_dev[kMaxDevices] = {false};
  static bool four_warps = false;
  bool& configured = configured_by_dev[current_device_slot()];
  if (!configured) {
    MAGB_CHECK_CUDA(cudaFuncSetAttribute(sdpa_bwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(BwdSmem)));
    MAGB_CHECK_CUDA(cudaFuncSetAttribute(sdpa_bwd_mma8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(Bwd8Smem)));
    const char* e = getenv("MAGB200_SDPA_BWD_WARPS");      // 4 = the original one-key-block-per-warp kernel
    four_warps = e && e[0] == '4';
    configured = true;
  }
  if (four_warps)
    sdpa_bwd_mma_kernel<<<dim3(heads, B), 128, sizeof(BwdSmem), st>>>(qkv, dctx, ctx, lse, mask, S, heads, th, ds, seed,
                                                                     site, oscale, dqkv, dbias);
  else
    MAGB_CHECK_CUDA(launch_pdl(sdpa_bwd_mma8_kernel, dim3(heads, B), dim3(256), sizeof(Bwd8Smem), st, qkv, dctx, ctx, lse, mask, S, heads, th, ds, seed,
                                                                      site, oscale, dqkv, dbias));
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace magb200

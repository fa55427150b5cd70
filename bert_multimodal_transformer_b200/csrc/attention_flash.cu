// Non-materialising (flash-style) attention core for any sequence length (used for S > 64; S <= 64 stays on the
// single-tile kernels of attention_mma.cu):  BertSelfAttention's
//     P = dropout(softmax(Q K^T / 8 + (1 - mask) * -10000)),  ctx = P V          
// and its backward, with scores and probabilities living only in mma fragments: HBM sees q|k|v, ctx, lse and the
// gradients, never an S x S tensor (BASELINE.json configs[4]: S = 512, where the materialising kernels of
// attention.cu would move 12.6 MB of fp32 scores per (batch, layer) and run out of shared memory).
//
//   forward  : CTA = 64 queries of one (batch, head), 4 warps x 16 queries; loop over 64-key blocks with online
//              softmax (running max / sum per query row), K/V blocks double-buffered with cp.async; emits lse.
//   backward : two passes that each own their outputs (no atomics, deterministic):
//                dK,dV pass: CTA = 64 keys, loop over query blocks  (S^T, dP^T -> Pd^T, dS^T -> dV, dK)
//                dQ    pass: CTA = 64 queries, loop over key blocks (S, dP -> dS -> dQ)
//              P is rebuilt from lse, D[q] = rowsum(dO * O) comes from a small pre-pass.
// Tensor cores: mma.sync.m16n8k16 (fp16 in, fp32 accumulate); transposed B operands (V, K, Q, dO as [k][n] tiles)
// are read with ldmatrix.trans straight from the row-major tiles, so nothing is transposed through shared memory.
// Dropout uses the same counter convention as the other attention kernels (element index of P[b][h][i][j]).
#include <stdlib.h>
#include "common.cuh"
#include "host.h"

namespace magb200 {

namespace {

constexpr int D = 64;      // head dim
constexpr int BLK = 64;    // queries / keys per block
constexpr int PITCH = 72;  // halfs per smem row (144 B: 16-byte aligned rows, conflict-free fragment and ldmatrix reads)
constexpr int TILE = BLK * PITCH;

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
// A fragments (16 rows from r0, 16 k-columns from k0) of a row-major [row][k] tile
__device__ __forceinline__ void frag_a(const __half* s, int r0, int k0, int g, int t, uint32_t (&a)[4]) {
  a[0] = ld32(s + (r0 + g) * PITCH + k0 + 2 * t);
  a[1] = ld32(s + (r0 + g + 8) * PITCH + k0 + 2 * t);
  a[2] = ld32(s + (r0 + g) * PITCH + k0 + 2 * t + 8);
  a[3] = ld32(s + (r0 + g + 8) * PITCH + k0 + 2 * t + 8);
}
// B fragments for B[k][n] = s[n0 + n][k0 + k]  (an [n][k] row-major tile: K^T from K, Q^T from Q, ...)
__device__ __forceinline__ void frag_b(const __half* s, int n0, int k0, int g, int t, uint32_t (&b)[2]) {
  b[0] = ld32(s + (n0 + g) * PITCH + k0 + 2 * t);
  b[1] = ld32(s + (n0 + g) * PITCH + k0 + 2 * t + 8);
}
// B fragments for B[k][n] = s[k0 + k][n0 + n]  (a [k][n] row-major tile) for the two n-tiles n0 and n0 + 8:
// r[0], r[1] = (b0, b1) of n-tile n0; r[2], r[3] = (b0, b1) of n-tile n0 + 8.  ldmatrix.trans hands thread (g, t) the
// elements (row 2t, 2t+1; column g) of each 8x8 block, which is exactly the col-major B fragment.
__device__ __forceinline__ void frag_b_trans2(const __half* s, int k0, int n0, int lane, uint32_t (&r)[4]) {
  const int m = lane >> 3, rr = lane & 7;
  const __half* p = s + (k0 + (m & 1) * 8 + rr) * PITCH + n0 + (m >> 1) * 8;
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}

// rows [row0, row0 + 64) x 64 halfs of a global fp16 matrix (row pitch ld) -> smem tile, rows >= S zero-filled
__device__ __forceinline__ void load_tile_async(const __half* __restrict__ g, int ld, int row0, int S, __half* s) {
  for (int i = threadIdx.x; i < BLK * 8; i += blockDim.x) {
    const int r = i >> 3, c = (i & 7) * 8;
    const bool ok = row0 + r < S;
    const __half* src = g + (int64_t)(ok ? row0 + r : 0) * ld + c;
    const uint32_t n = ok ? 16u : 0u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(s + r * PITCH + c)), "l"(src), "r"(n)
                 : "memory");
  }
}
__device__ __forceinline__ void async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// keep flags of the 64 x 64 tile (queries q0.., keys k0..) of P[b][h]: bit (ql * 64 + kl).  Thread `tid` (< 128) owns
// one 32-bit word = 32 consecutive keys of one query row: the element indices are contiguous, i.e. <= 9 Philox
// counters per word.
__device__ __forceinline__ void fill_keep_tile(uint32_t* keep, uint64_t seed, uint32_t site, uint64_t bh, int S, int q0,
                                               int k0, uint32_t thresh) {
  for (int wd = threadIdx.x; wd < BLK * 2; wd += blockDim.x) {
    const int ql = wd >> 1, kl0 = (wd & 1) * 32;
    const int q = q0 + ql;
    uint32_t bits = 0;
    if (q < S && k0 + kl0 < S) {
      const uint64_t first = (bh * S + q) * (uint64_t)S + k0 + kl0;
      const int n = min(32, S - (k0 + kl0));
      const uint64_t last = first + n - 1;
      for (uint64_t c = first >> 2; c <= (last >> 2); ++c) {
        const uint4 u = philox4x32(seed, site, c);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int64_t loc = (int64_t)(4 * c + k) - (int64_t)first;
          if (loc >= 0 && loc < n && w[k] >= thresh) bits |= 1u << loc;
        }
      }
    }
    keep[wd] = bits;
  }
}
__device__ __forceinline__ bool keep_bit(const uint32_t* keep, int ql, int kl) {
  return (keep[ql * 2 + (kl >> 5)] >> (kl & 31)) & 1u;
}

// additive key mask of a 64-key block: (1 - m) * -10000 (bert.py:180-182, transformers 3.0.2), -inf beyond S
__device__ __forceinline__ void fill_mask_block(float* sm, const int64_t* mrow, int k0, int S) {
  if (threadIdx.x < BLK) {
    const int j = k0 + threadIdx.x;
    float v = -INFINITY;
    if (j < S) v = mrow == nullptr ? 0.f : (1.0f - (float)mrow[j]) * -10000.0f;
    sm[threadIdx.x] = v;
  }
}

// ------------------------------------------------------------------------------------------ forward
struct FwdSmem {
  __half q[TILE];
  __half k[2][TILE];
  __half v[2][TILE];
  float madd[2][BLK];
  uint32_t keep[BLK * 2];
};

__global__ void __launch_bounds__(128) flash_fwd_kernel(const __half* __restrict__ qkv, const int64_t* __restrict__ mask,
                                                        int S, int heads, uint32_t thresh, float dscale, Seed seed_in,
                                                        uint32_t site, __half* __restrict__ ctx, float* __restrict__ lse) {
  pdl_wait();
  pdl_trigger();
  const uint64_t seed = seed_in.ptr ? *seed_in.ptr : seed_in.value;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  FwdSmem& sm = *reinterpret_cast<FwdSmem*>(smem_raw);
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int Hd = heads * D, ld = 3 * Hd;
  const __half* base = qkv + (int64_t)b * S * ld + h * D;
  const int64_t* mrow = mask ? mask + (int64_t)b * S : nullptr;
  const uint64_t bh = (uint64_t)b * heads + h;
  const int q0 = qb * BLK;
  const int nkb = (S + BLK - 1) / BLK;

  load_tile_async(base, ld, q0, S, sm.q);
  load_tile_async(base + Hd, ld, 0, S, sm.k[0]);
  load_tile_async(base + 2 * Hd, ld, 0, S, sm.v[0]);
  async_commit();

  const int r0 = warp * 16;
  float o[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  for (int kb = 0; kb < nkb; ++kb) {
    const int buf = kb & 1;
    const int k0 = kb * BLK;
    if (kb + 1 < nkb) {      // the other buffer was last read in iteration kb-1, which ended with a barrier
      load_tile_async(base + Hd, ld, k0 + BLK, S, sm.k[buf ^ 1]);
      load_tile_async(base + 2 * Hd, ld, k0 + BLK, S, sm.v[buf ^ 1]);
    }
    async_commit();
    fill_mask_block(sm.madd[buf], mrow, k0, S);
    if (thresh != 0) fill_keep_tile(sm.keep, seed, site, bh, S, q0, k0, thresh);
    async_wait<1>();           // everything but the prefetch just issued has landed
    __syncthreads();
    if (q0 + r0 < S) {
      const __half* sK = sm.k[buf];
      const __half* sV = sm.v[buf];
      float sc[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t a[4];
        frag_a(sm.q, r0, kk * 16, g, t, a);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          uint32_t bb[2];
          frag_b(sK, nt * 8, kk * 16, g, t, bb);
          mma16816(sc[nt], a[0], a[1], a[2], a[3], bb[0], bb[1]);
        }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float ma = sm.madd[buf][nt * 8 + 2 * t + e];
          sc[nt][e] = sc[nt][e] * 0.125f + ma;
          sc[nt][2 + e] = sc[nt][2 + e] * 0.125f + ma;
          mx0 = fmaxf(mx0, sc[nt][e]);
          mx1 = fmaxf(mx1, sc[nt][2 + e]);
        }
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);     // finite: every block holds at least one key < S
      const float c0 = expf(m0 - mn0), c1 = expf(m1 - mn1);       // exp(-inf) = 0 on the first block
      m0 = mn0; m1 = mn1;
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          sc[nt][e] = expf(sc[nt][e] - mn0);
          sc[nt][2 + e] = expf(sc[nt][2 + e] - mn1);
          s0 += sc[nt][e];
          s1 += sc[nt][2 + e];
        }
      }
      s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
      s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
      s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
      s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
      l0 = l0 * c0 + s0;
      l1 = l1 * c1 + s1;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        o[nt][0] *= c0; o[nt][1] *= c0; o[nt][2] *= c1; o[nt][3] *= c1;
      }
      if (thresh != 0) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const int j = nt * 8 + 2 * t;
          sc[nt][0] = keep_bit(sm.keep, r0 + g, j) ? sc[nt][0] * dscale : 0.f;
          sc[nt][1] = keep_bit(sm.keep, r0 + g, j + 1) ? sc[nt][1] * dscale : 0.f;
          sc[nt][2] = keep_bit(sm.keep, r0 + g + 8, j) ? sc[nt][2] * dscale : 0.f;
          sc[nt][3] = keep_bit(sm.keep, r0 + g + 8, j + 1) ? sc[nt][3] * dscale : 0.f;
        }
      }
      // O += P V     (B[k = key][n = d] = V[key][d])
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint32_t a0 = packh(sc[2 * kk][0], sc[2 * kk][1]), a1 = packh(sc[2 * kk][2], sc[2 * kk][3]);
        const uint32_t a2 = packh(sc[2 * kk + 1][0], sc[2 * kk + 1][1]), a3 = packh(sc[2 * kk + 1][2], sc[2 * kk + 1][3]);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t bb[4];
          frag_b_trans2(sV, kk * 16, np * 16, lane, bb);
          mma16816(o[2 * np], a0, a1, a2, a3, bb[0], bb[1]);
          mma16816(o[2 * np + 1], a0, a1, a2, a3, bb[2], bb[3]);
        }
      }
    }
    __syncthreads();           // this block's K/V buffer, mask and keep bits may be overwritten next iteration
  }
  if (q0 + r0 < S) {
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
    const int i0 = q0 + r0 + g, i1 = i0 + 8;
    if (lse != nullptr && t == 0) {
      if (i0 < S) lse[bh * S + i0] = m0 + logf(l0);
      if (i1 < S) lse[bh * S + i1] = m1 + logf(l1);
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int d = nt * 8 + 2 * t;
      if (i0 < S) *reinterpret_cast<uint32_t*>(ctx + ((int64_t)b * S + i0) * Hd + h * D + d) = packh(o[nt][0] * inv0, o[nt][1] * inv0);
      if (i1 < S) *reinterpret_cast<uint32_t*>(ctx + ((int64_t)b * S + i1) * Hd + h * D + d) = packh(o[nt][2] * inv1, o[nt][3] * inv1);
    }
  }
}

// ------------------------------------------------------------------------------------------ backward pre-pass
// Dq[b][h][q] = sum_d dO[b, q, h, d] * O[b, q, h, d]: two threads per (token, head)
__global__ void __launch_bounds__(256) flash_bwd_prep_kernel(const __half* __restrict__ dctx, const __half* __restrict__ ctx,
                                                             int B, int S, int heads, float* __restrict__ Dq) {
  pdl_wait();
  pdl_trigger();
  const int64_t idx = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 1;
  const int half = threadIdx.x & 1;
  const int64_t total = (int64_t)B * S * heads;
  float acc = 0.f;
  const bool ok = idx < total;
  int64_t tok = 0; int h = 0;
  if (ok) {
    tok = idx / heads; h = (int)(idx % heads);
    const __half* o = ctx + tok * heads * D + h * D + half * 32;
    const __half* d = dctx + tok * heads * D + h * D + half * 32;
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
  if (ok && half == 0) {
    const int64_t b = tok / S, q = tok % S;
    Dq[(b * heads + h) * S + q] = acc;
  }
}

// ------------------------------------------------------------------------------------------ backward: dK, dV
struct BwdKvSmem {
  __half k[TILE], v[TILE];
  __half q[2][TILE], dO[2][TILE];
  float lse[2][BLK], Dq[2][BLK];
  uint32_t keep[BLK * 2];
};

__global__ void __launch_bounds__(128) flash_bwd_dkv_kernel(const __half* __restrict__ qkv, const __half* __restrict__ dctx,
                                                            const float* __restrict__ lse, const float* __restrict__ Dq,
                                                            const int64_t* __restrict__ mask, int S, int heads,
                                                            uint32_t thresh, float dscale, Seed seed_in, uint32_t site,
                                                            __half* __restrict__ dqkv) {
  pdl_wait();
  pdl_trigger();
  const uint64_t seed = seed_in.ptr ? *seed_in.ptr : seed_in.value;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  BwdKvSmem& sm = *reinterpret_cast<BwdKvSmem*>(smem_raw);
  const int kb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int Hd = heads * D, ld = 3 * Hd;
  const __half* base = qkv + (int64_t)b * S * ld + h * D;
  const __half* dObase = dctx + (int64_t)b * S * Hd + h * D;
  const uint64_t bh = (uint64_t)b * heads + h;
  const int k0 = kb * BLK;
  const int nqb = (S + BLK - 1) / BLK;

  load_tile_async(base + Hd, ld, k0, S, sm.k);
  load_tile_async(base + 2 * Hd, ld, k0, S, sm.v);
  load_tile_async(base, ld, 0, S, sm.q[0]);
  load_tile_async(dObase, Hd, 0, S, sm.dO[0]);
  async_commit();

  const int r0 = warp * 16;
  const int key0 = k0 + r0 + g, key1 = key0 + 8;
  float ma0 = 0.f, ma1 = 0.f;
  if (mask != nullptr) {
    if (key0 < S) ma0 = (1.0f - (float)mask[(int64_t)b * S + key0]) * -10000.0f;
    if (key1 < S) ma1 = (1.0f - (float)mask[(int64_t)b * S + key1]) * -10000.0f;
  }
  float dv[8][4], dk[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    dv[nt][0] = dv[nt][1] = dv[nt][2] = dv[nt][3] = 0.f;
    dk[nt][0] = dk[nt][1] = dk[nt][2] = dk[nt][3] = 0.f;
  }

  for (int qb = 0; qb < nqb; ++qb) {
    const int buf = qb & 1;
    const int q0 = qb * BLK;
    if (qb + 1 < nqb) {
      load_tile_async(base, ld, q0 + BLK, S, sm.q[buf ^ 1]);
      load_tile_async(dObase, Hd, q0 + BLK, S, sm.dO[buf ^ 1]);
    }
    async_commit();
    if (threadIdx.x < BLK) {
      const int q = q0 + threadIdx.x;
      sm.lse[buf][threadIdx.x] = q < S ? lse[bh * S + q] : 0.f;
      sm.Dq[buf][threadIdx.x] = q < S ? Dq[bh * S + q] : 0.f;
    }
    if (thresh != 0) fill_keep_tile(sm.keep, seed, site, bh, S, q0, k0, thresh);
    async_wait<1>();
    __syncthreads();
    if (k0 + r0 < S) {
      const __half* sQ = sm.q[buf];
      const __half* sdO = sm.dO[buf];
      // S^T = K Q^T (keys x queries), dP^T = V dO^T
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
          frag_b(sQ, nt * 8, kk * 16, g, t, bq);        // B[k = d][n = query] = Q[query][d]
          frag_b(sdO, nt * 8, kk * 16, g, t, bo);
          mma16816(st[nt], ak[0], ak[1], ak[2], ak[3], bq[0], bq[1]);
          mma16816(dp[nt], av[0], av[1], av[2], av[3], bo[0], bo[1]);
        }
      }
      // P^T -> Pd^T (dropped probabilities) in dp, dS^T in st
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = (e < 2) ? key0 : key1;
          const int ql = nt * 8 + 2 * t + (e & 1);
          float dsv = 0.f, pdv = 0.f;
          if (key < S && q0 + ql < S) {
            const float p = expf(st[nt][e] * 0.125f + ((e < 2) ? ma0 : ma1) - sm.lse[buf][ql]);
            const float ds = thresh == 0 ? 1.f : (keep_bit(sm.keep, ql, key - k0) ? dscale : 0.f);
            pdv = p * ds;
            dsv = p * (dp[nt][e] * ds - sm.Dq[buf][ql]) * 0.125f;      // includes d(score)/d(qk) = 1/sqrt(64)
          }
          dp[nt][e] = pdv;
          st[nt][e] = dsv;
        }
      }
      // dV += Pd^T dO, dK += dS^T Q   (B[k = query][n = d] = dO[query][d] / Q[query][d])
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint32_t p0 = packh(dp[2 * kk][0], dp[2 * kk][1]), p1 = packh(dp[2 * kk][2], dp[2 * kk][3]);
        const uint32_t p2 = packh(dp[2 * kk + 1][0], dp[2 * kk + 1][1]), p3 = packh(dp[2 * kk + 1][2], dp[2 * kk + 1][3]);
        const uint32_t s0 = packh(st[2 * kk][0], st[2 * kk][1]), s1 = packh(st[2 * kk][2], st[2 * kk][3]);
        const uint32_t s2 = packh(st[2 * kk + 1][0], st[2 * kk + 1][1]), s3 = packh(st[2 * kk + 1][2], st[2 * kk + 1][3]);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t bo[4], bq[4];
          frag_b_trans2(sdO, kk * 16, np * 16, lane, bo);
          frag_b_trans2(sQ, kk * 16, np * 16, lane, bq);
          mma16816(dv[2 * np], p0, p1, p2, p3, bo[0], bo[1]);
          mma16816(dv[2 * np + 1], p0, p1, p2, p3, bo[2], bo[3]);
          mma16816(dk[2 * np], s0, s1, s2, s3, bq[0], bq[1]);
          mma16816(dk[2 * np + 1], s0, s1, s2, s3, bq[2], bq[3]);
        }
      }
    }
    __syncthreads();
  }
  if (k0 + r0 < S) {
    __half* obase = dqkv + (int64_t)b * S * ld + h * D;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int d = nt * 8 + 2 * t;
      if (key0 < S) {
        *reinterpret_cast<uint32_t*>(obase + (int64_t)key0 * ld + Hd + d) = packh(dk[nt][0], dk[nt][1]);
        *reinterpret_cast<uint32_t*>(obase + (int64_t)key0 * ld + 2 * Hd + d) = packh(dv[nt][0], dv[nt][1]);
      }
      if (key1 < S) {
        *reinterpret_cast<uint32_t*>(obase + (int64_t)key1 * ld + Hd + d) = packh(dk[nt][2], dk[nt][3]);
        *reinterpret_cast<uint32_t*>(obase + (int64_t)key1 * ld + 2 * Hd + d) = packh(dv[nt][2], dv[nt][3]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ backward: dQ
struct BwdQSmem {
  __half q[TILE], dO[TILE];
  __half k[2][TILE], v[2][TILE];
  float lse[BLK], Dq[BLK];
  float madd[2][BLK];
  uint32_t keep[BLK * 2];
};

__global__ void __launch_bounds__(128) flash_bwd_dq_kernel(const __half* __restrict__ qkv, const __half* __restrict__ dctx,
                                                           const float* __restrict__ lse, const float* __restrict__ Dq,
                                                           const int64_t* __restrict__ mask, int S, int heads,
                                                           uint32_t thresh, float dscale, Seed seed_in, uint32_t site,
                                                           __half* __restrict__ dqkv) {
  pdl_wait();
  pdl_trigger();
  const uint64_t seed = seed_in.ptr ? *seed_in.ptr : seed_in.value;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  BwdQSmem& sm = *reinterpret_cast<BwdQSmem*>(smem_raw);
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int Hd = heads * D, ld = 3 * Hd;
  const __half* base = qkv + (int64_t)b * S * ld + h * D;
  const __half* dObase = dctx + (int64_t)b * S * Hd + h * D;
  const int64_t* mrow = mask ? mask + (int64_t)b * S : nullptr;
  const uint64_t bh = (uint64_t)b * heads + h;
  const int q0 = qb * BLK;
  const int nkb = (S + BLK - 1) / BLK;

  load_tile_async(base, ld, q0, S, sm.q);
  load_tile_async(dObase, Hd, q0, S, sm.dO);
  load_tile_async(base + Hd, ld, 0, S, sm.k[0]);
  load_tile_async(base + 2 * Hd, ld, 0, S, sm.v[0]);
  async_commit();
  if (threadIdx.x < BLK) {
    const int q = q0 + threadIdx.x;
    sm.lse[threadIdx.x] = q < S ? lse[bh * S + q] : 0.f;
    sm.Dq[threadIdx.x] = q < S ? Dq[bh * S + q] : 0.f;
  }
  const int r0 = warp * 16;
  float dq[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) dq[nt][0] = dq[nt][1] = dq[nt][2] = dq[nt][3] = 0.f;

  for (int kb = 0; kb < nkb; ++kb) {
    const int buf = kb & 1;
    const int k0 = kb * BLK;
    if (kb + 1 < nkb) {
      load_tile_async(base + Hd, ld, k0 + BLK, S, sm.k[buf ^ 1]);
      load_tile_async(base + 2 * Hd, ld, k0 + BLK, S, sm.v[buf ^ 1]);
    }
    async_commit();
    fill_mask_block(sm.madd[buf], mrow, k0, S);
    if (thresh != 0) fill_keep_tile(sm.keep, seed, site, bh, S, q0, k0, thresh);
    async_wait<1>();
    __syncthreads();
    if (q0 + r0 < S) {
      const __half* sK = sm.k[buf];
      const __half* sV = sm.v[buf];
      const float l0 = sm.lse[r0 + g], l1 = sm.lse[r0 + g + 8];
      const float D0 = sm.Dq[r0 + g], D1 = sm.Dq[r0 + g + 8];
      // S = Q K^T, dP = dO V^T   (queries x keys)
      float sc[8][4], dp[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
        dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t aq[4], ao[4];
        frag_a(sm.q, r0, kk * 16, g, t, aq);
        frag_a(sm.dO, r0, kk * 16, g, t, ao);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          uint32_t bk[2], bv[2];
          frag_b(sK, nt * 8, kk * 16, g, t, bk);        // B[k = d][n = key] = K[key][d]
          frag_b(sV, nt * 8, kk * 16, g, t, bv);
          mma16816(sc[nt], aq[0], aq[1], aq[2], aq[3], bk[0], bk[1]);
          mma16816(dp[nt], ao[0], ao[1], ao[2], ao[3], bv[0], bv[1]);
        }
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int kl = nt * 8 + 2 * t + (e & 1);
          const int ql = r0 + g + ((e < 2) ? 0 : 8);
          float dsv = 0.f;
          if (k0 + kl < S && q0 + ql < S) {
            const float p = expf(sc[nt][e] * 0.125f + sm.madd[buf][kl] - ((e < 2) ? l0 : l1));
            const float ds = thresh == 0 ? 1.f : (keep_bit(sm.keep, ql, kl) ? dscale : 0.f);
            dsv = p * (dp[nt][e] * ds - ((e < 2) ? D0 : D1)) * 0.125f;
          }
          sc[nt][e] = dsv;
        }
      }
      // dQ += dS K   (B[k = key][n = d] = K[key][d])
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint32_t a0 = packh(sc[2 * kk][0], sc[2 * kk][1]), a1 = packh(sc[2 * kk][2], sc[2 * kk][3]);
        const uint32_t a2 = packh(sc[2 * kk + 1][0], sc[2 * kk + 1][1]), a3 = packh(sc[2 * kk + 1][2], sc[2 * kk + 1][3]);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t bb[4];
          frag_b_trans2(sK, kk * 16, np * 16, lane, bb);
          mma16816(dq[2 * np], a0, a1, a2, a3, bb[0], bb[1]);
          mma16816(dq[2 * np + 1], a0, a1, a2, a3, bb[2], bb[3]);
        }
      }
    }
    __syncthreads();
  }
  if (q0 + r0 < S) {
    __half* obase = dqkv + (int64_t)b * S * ld + h * D;
    const int i0 = q0 + r0 + g, i1 = i0 + 8;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int d = nt * 8 + 2 * t;
      if (i0 < S) *reinterpret_cast<uint32_t*>(obase + (int64_t)i0 * ld + d) = packh(dq[nt][0], dq[nt][1]);
      if (i1 < S) *reinterpret_cast<uint32_t*>(obase + (int64_t)i1 * ld + d) = packh(dq[nt][2], dq[nt][3]);
    }
  }
}

inline void drop_params(float p, uint32_t& thresh, float& scale) {
  if (p <= 0.f) { thresh = 0; scale = 1.f; return; }
  double t = (double)p * 4294967296.0;
  thresh = t >= 4294967295.0 ? 4294967295u : (uint32_t)t;
  scale = 1.0f / (1.0f - p);
}

struct FlashConfig {
  bool done[16] = {false};
};
FlashConfig g_cfg;

int configure(int dev) {
  if (g_cfg.done[dev]) return 0;
  MAGB_CHECK_CUDA(cudaFuncSetAttribute(flash_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FwdSmem)));
  MAGB_CHECK_CUDA(cudaFuncSetAttribute(flash_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdKvSmem)));
  MAGB_CHECK_CUDA(cudaFuncSetAttribute(flash_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdQSmem)));
  g_cfg.done[dev] = true;
  return 0;
}

}  // namespace

size_t sdpa_flash_scratch_bytes(int B, int S, int heads) { return (size_t)B * heads * S * sizeof(float); }

int sdpa_fwd_flash(const __half* qkv, const int64_t* mask, int B, int S, int heads, float p, Seed seed, uint32_t site,
                   __half* ctx, float* lse, cudaStream_t st) {
  MAGB_REQUIRE(S >= 1 && B >= 1, "sdpa_fwd_flash: bad shape B=%d S=%d", B, S);
  int dev = 0;
  MAGB_CHECK_CUDA(cudaGetDevice(&dev));
  MAGB_REQUIRE(dev >= 0 && dev < 16, "sdpa_fwd_flash: device index %d out of range", dev);
  int rc = configure(dev);
  if (rc) return rc;
  uint32_t th; float ds;
  drop_params(p, th, ds);
  MAGB_CHECK_CUDA(launch_pdl(flash_fwd_kernel, dim3((S + BLK - 1) / BLK, heads, B), dim3(128), sizeof(FwdSmem), st, qkv, mask,
                             S, heads, th, ds, seed, site, ctx, lse));
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// Dq_scratch: [B][heads][S] fp32 (sdpa_flash_scratch_bytes)
int sdpa_bwd_flash(const __half* qkv, const __half* dctx, const __half* ctx, const float* lse, const int64_t* mask, int B,
                   int S, int heads, float p, Seed seed, uint32_t site, float* Dq_scratch, __half* dqkv, cudaStream_t st) {
  MAGB_REQUIRE(S >= 1 && B >= 1, "sdpa_bwd_flash: bad shape B=%d S=%d", B, S);
  MAGB_REQUIRE(lse != nullptr && ctx != nullptr && Dq_scratch != nullptr, "sdpa_bwd_flash: lse, ctx and the D scratch are required");
  int dev = 0;
  MAGB_CHECK_CUDA(cudaGetDevice(&dev));
  MAGB_REQUIRE(dev >= 0 && dev < 16, "sdpa_bwd_flash: device index %d out of range", dev);
  int rc = configure(dev);
  if (rc) return rc;
  uint32_t th; float ds;
  drop_params(p, th, ds);
  const int64_t pairs = (int64_t)B * S * heads * 2;
  MAGB_CHECK_CUDA(launch_pdl(flash_bwd_prep_kernel, dim3((unsigned)((pairs + 255) / 256)), dim3(256), 0, st, dctx, ctx, B, S,
                             heads, Dq_scratch));
  count_launch();
  const dim3 grid((S + BLK - 1) / BLK, heads, B);
  MAGB_CHECK_CUDA(launch_pdl(flash_bwd_dkv_kernel, grid, dim3(128), sizeof(BwdKvSmem), st, qkv, dctx, lse,
                             (const float*)Dq_scratch, mask, S, heads, th, ds, seed, site, dqkv));
  count_launch();
  MAGB_CHECK_CUDA(launch_pdl(flash_bwd_dq_kernel, grid, dim3(128), sizeof(BwdQSmem), st, qkv, dctx, lse,
                             (const float*)Dq_scratch, mask, S, heads, th, ds, seed, site, dqkv));
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace magb200

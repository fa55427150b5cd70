// Linear + bias + dropout + residual + LayerNorm in ONE tcgen05 kernel (BertSelfOutput / BertOutput):
//
//     y = LayerNorm(dropout(A W^T + b) + res) ,   hidden size N = 768 = 3 x 256 columns
//
// LayerNorm needs statistics over a whole row, i.e. over three 256-column accumulator tiles -- more than the 512 TMEM
// columns of one SM.  So a CLUSTER of 2 x 3 CTAs owns 256 token rows x 768 columns: the two CTAs of a pair (cluster x)
// run tcgen05.mma.cta_group::2 on a 256 x 256 tile exactly like gemm2_tcgen05_kernel, the three pairs (cluster y) own
// the three column tiles, and the row sums travel through distributed shared memory:
//
//   (while the last MMAs drain, the TMA producer refills the now idle pipeline stages with the CTA's 128 x 256 fp32
//    residual tile as SWIZZLE_128B boxes: one load latency in total instead of one per 32-column chunk)
//   pass 1   v = acc + bias -> dropout -> + residual;  s1 = sum(v) over the CTA's 256 columns;  v goes BACK to TMEM
//            (tcgen05.st, 256 B/clk) so nothing is recomputed; s1 -> the 3 CTAs of the row group (st.shared::cluster)
//   barrier  mean = (s1_0 + s1_1 + s1_2) / 768
//   pass 1b  s2 = sum((v - mean)^2)  (two-pass variance, like torch's layer_norm) -> exchanged the same way
//   barrier  rstd = 1 / sqrt(s2 / 768 + eps)
//   pass 2   y = (v - mean) * rstd * gamma + beta -> fp32 residual stream + fp16 operand copy, coalesced through a
//            staging tile; rstd is saved for the backward pass
//
// This removes the dropout_residual_ln_fwd launch and the fp32 `pre` round trip (write 3 KB + read 3 KB per token)
// behind each of the two N = 768 Linears of a layer.
#pragma once

#include "gemm_sm100.cuh"

namespace magb200 {

struct GemmLnArgs {
  int M, K;                 // N is 768
  int kb_total;
  const float* bias;        // [768]
  const float* res;         // [M][768] fp32 residual
  const float* gamma; const float* beta; float eps;
  float* y32; __half* y16; float* rstd;
  uint32_t idesc;
  // dropout on the Linear output (element index m * 768 + n, as dropout_residual_ln_fwd)
  uint32_t drop_thresh; float drop_scale; uint32_t drop_site;
  unsigned long long drop_seed; const unsigned long long* drop_seed_ptr;
};

#ifdef __CUDACC__

constexpr int LN_N = 768, LN_BN = 256, LN_NT = 3;

template <int STAGES>
struct GemmLnSmem {
  static constexpr int A_BYTES = 128 * 128;
  static constexpr int B_BYTES = (LN_BN / 2) * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int PAR_OFFSET = BAR_OFFSET + 128;             // bias | gamma | beta of this CTA's 256 columns
  static constexpr int RED_OFFSET = PAR_OFFSET + 3 * LN_BN * 4;   // [2 passes][3 column tiles][128 rows] partial sums
  static constexpr int TOTAL = RED_OFFSET + 2 * LN_NT * 128 * 4 + 1024;
  static constexpr int STG_PITCH = 32 * 4 + 16;                   // staging row: one 32-column fp32 chunk + bank skew
  static constexpr int STG_WARP = 32 * STG_PITCH;
  static constexpr int RES_BYTES = 4 * (LN_BN / 32) * 4096;       // residual tile: [4 warps][8 chunks] boxes of 32 x 32 fp32
  static_assert(4 * STG_WARP <= STAGES * STAGE_BYTES, "staging does not fit in the pipeline smem");
  static_assert(RES_BYTES <= STAGES * STAGE_BYTES, "the residual tile is staged in the pipeline smem");
  static_assert((2 * STAGES + 2) * 8 + 4 <= 128, "barrier area");
};

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_shared_cluster_f32(uint32_t cluster_addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(cluster_addr), "f"(v) : "memory");
}
// completion of the leader's prior tcgen05.mma.cta_group::2 arrives on the barrier at this offset in the CTAs of `mask`
__device__ __forceinline__ void tc_commit_2sm_mask(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}

template <int STAGES>
__global__ void __launch_bounds__(192, 1)
gemm2_ln_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const __grid_constant__ CUtensorMap tmR, const GemmLnArgs g) {
  using T = GemmTile<2>;
  using SM = GemmLnSmem<STAGES>;
  constexpr int BN = LN_BN;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* res_bar = tmem_full_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 1);
  float* s_bias = reinterpret_cast<float*>(smem + SM::PAR_OFFSET);
  float* s_gamma = s_bias + BN;
  float* s_beta = s_gamma + BN;
  float* s_red = reinterpret_cast<float*>(smem + SM::RED_OFFSET);      // [2][3][128]

  const int warp = warp_idx_uniform();
  const int lane = threadIdx.x & 31;
  // cluster (2, 3, 1): rank = x + 2 y.  x = which 128 rows of the 256-row pair tile, y = which 256-column tile.
  const uint32_t crank = __shfl_sync(0xffffffffu, cluster_ctarank(), 0);
  const uint32_t cx = crank & 1, cy = crank >> 1;
  const bool leader = cx == 0;
  const uint32_t leader_rank = crank & ~1u;
  const uint16_t pair_mask = (uint16_t)(3u << leader_rank);
  const int m0 = blockIdx.x * T::BM;
  const int n0 = (int)cy * BN;
  const int num_kb = g.kb_total;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmR);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      mbar_init(tmem_full_bar, 1);
      mbar_init(res_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_2sm(tmem_slot, BN);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (every CTA; one lane issues)
    const int nb0 = n0 + (int)cx * (BN / 2);          // this CTA's half of the pair's B tile
    for (int base = 0; base < num_kb; base += STAGES) {
      const uint32_t ph = (base / STAGES) & 1;
#pragma unroll
      for (int s = 0; s < STAGES; ++s) {
        if (base + s < num_kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (elect_one_sync()) {
            if (leader) mbar_arrive_expect_tx(&full_bar[s], 2 * SM::STAGE_BYTES);
            const uint32_t bar = mapa_shared(&full_bar[s], leader_rank);
            uint8_t* a_dst = smem + s * SM::STAGE_BYTES;
            const int k0 = (base + s) * T::BK;
            tma_load_2d_2sm(a_dst, &tmA, bar, k0, m0);
            tma_load_2d_2sm(a_dst + SM::A_BYTES, &tmB, bar, k0, nb0);
          }
          __syncwarp();
        }
      }
    }
    // every stage has been read by its last MMA -> the pipeline smem is idle: fetch the residual tile into it,
    // box (q, c) = rows m0 + 32 q .., columns n0 + 32 c .. at offset (8 q + c) * 4096 (rows beyond M are zero-filled)
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      const int uses = (num_kb - s + STAGES - 1) / STAGES;          // how often stage s was filled
      if (uses > 0) mbar_wait(&empty_bar[s], (uint32_t)((uses - 1) & 1));
    }
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(res_bar, SM::RES_BYTES);
      for (int b = 0; b < 4 * (BN / 32); ++b)
        tma_load_2d(smem + b * 4096, &tmR, res_bar, n0 + (b & 7) * 32, m0 + (b >> 3) * 32);
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (pair leaders; one lane issues)
    if (leader) {
      const uint32_t sb16 = smem_u32(smem) >> 4;
      const uint32_t dhi = smem_desc_hi_sw128(1024);
      for (int base = 0; base < num_kb; base += STAGES) {
        const uint32_t ph = (base / STAGES) & 1;
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
          if (base + s < num_kb) {
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            if (elect_one_sync()) {
#pragma unroll
              for (int k = 0; k < T::BK / T::UK; ++k) {
                const uint64_t adesc = smem_desc_from(dhi, sb16 + ((s * SM::STAGE_BYTES + k * 32) >> 4), 0);
                const uint64_t bdesc = smem_desc_from(dhi, sb16 + ((s * SM::STAGE_BYTES + SM::A_BYTES + k * 32) >> 4), 0);
                tc_mma_f16_2sm(tmem_base, adesc, bdesc, g.idesc, (s | k) != 0 ? 1u : (uint32_t)(base != 0));
              }
              tc_commit_2sm_mask(&empty_bar[s], pair_mask);
            }
            __syncwarp();
          }
        }
      }
      if (elect_one_sync()) tc_commit_2sm_mask(tmem_full_bar, pair_mask);
      __syncwarp();
    }
  }

  // ------------------------------------------------------------------ epilogue (warps 2..5), three passes
  const bool epi = warp >= 2;
  const int q = warp & 3;                              // TMEM lane quarter
  const int m_base = m0 + q * 32;
  const int rows = min(32, g.M - m_base);              // <= 0 for rows beyond M (still takes part in the barriers)
  const int row = q * 32 + lane;                       // row of this thread inside the CTA's 128
  const uint32_t tw = tmem_base + ((uint32_t)(q * 32) << 16);
  uint8_t* stg = smem + q * SM::STG_WARP;
  uint8_t* myrow = stg + lane * SM::STG_PITCH;
  const uint64_t seed = g.drop_seed_ptr ? *g.drop_seed_ptr : g.drop_seed;
  float mean = 0.f, rstd = 0.f;

  if (epi) {
    for (int i = threadIdx.x - 64; i < BN; i += 128) {       // parameters of this CTA's columns
      s_bias[i] = __ldg(g.bias + n0 + i);
      s_gamma[i] = __ldg(g.gamma + n0 + i);
      s_beta[i] = __ldg(g.beta + n0 + i);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");           // the four epilogue warps
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    mbar_wait(res_bar, 0);                                    // the residual tile has landed in the pipeline smem
    // ---- pass 1: v = dropout(acc + bias) + residual, row sum; v back to TMEM
    float s1 = 0.f;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tw + (uint32_t)(c * 32), r);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + s_bias[c * 32 + j];
      if (g.drop_thresh != 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint4 u = philox4x32(seed, g.drop_site, ((unsigned long long)(m_base + lane) * LN_N + n0 + c * 32 + 4 * j) >> 2);
          v[4 * j] = u.x >= g.drop_thresh ? v[4 * j] * g.drop_scale : 0.f;
          v[4 * j + 1] = u.y >= g.drop_thresh ? v[4 * j + 1] * g.drop_scale : 0.f;
          v[4 * j + 2] = u.z >= g.drop_thresh ? v[4 * j + 2] * g.drop_scale : 0.f;
          v[4 * j + 3] = u.w >= g.drop_thresh ? v[4 * j + 3] * g.drop_scale : 0.f;
        }
      }
      {
        // residual row of this thread in box (q, c): 16-byte chunk j lives at chunk j ^ (row & 7) (SWIZZLE_128B)
        const uint8_t* rrow = smem + (q * (BN / 32) + c) * 4096 + lane * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 x = *reinterpret_cast<const float4*>(rrow + ((j ^ (lane & 7)) << 4));
          v[4 * j] += x.x; v[4 * j + 1] += x.y; v[4 * j + 2] += x.z; v[4 * j + 3] += x.w;
        }
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        s1 += v[j];
        r[j] = __float_as_uint(v[j]);
      }
      tmem_st_32x32(tw + (uint32_t)(c * 32), r);
    }
    tmem_st_wait();
    // partial row sum -> the three CTAs that share these rows (same x, y = 0..2)
#pragma unroll
    for (uint32_t y = 0; y < LN_NT; ++y)
      st_shared_cluster_f32(mapa_shared(s_red + cy * 128 + row, cx + 2 * y), s1);
  }
  cluster_sync_all();                                   // every thread of the six CTAs
  if (epi) {
    mean = (s_red[row] + s_red[128 + row] + s_red[256 + row]) * (1.0f / LN_N);
    // ---- pass 1b: centred sum of squares
    float s2 = 0.f;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tw + (uint32_t)(c * 32), r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float d = __uint_as_float(r[j]) - mean;
        s2 = fmaf(d, d, s2);
      }
    }
#pragma unroll
    for (uint32_t y = 0; y < LN_NT; ++y)
      st_shared_cluster_f32(mapa_shared(s_red + 3 * 128 + cy * 128 + row, cx + 2 * y), s2);
  }
  cluster_sync_all();
  if (epi) {
    const float var = (s_red[3 * 128 + row] + s_red[4 * 128 + row] + s_red[5 * 128 + row]) * (1.0f / LN_N);
    rstd = 1.0f / sqrtf(var + g.eps);
    if (g.rstd != nullptr && cy == 0 && lane < rows) g.rstd[m_base + lane] = rstd;
    // ---- pass 2: normalise, fp32 + fp16 outputs through the staging tile (coalesced 128 B / 64 B row pieces)
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tw + (uint32_t)(c * 32), r);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j)
        v[j] = (__uint_as_float(r[j]) - mean) * rstd * s_gamma[c * 32 + j] + s_beta[c * 32 + j];
      row_write_f32(myrow, 0, v);
      __syncwarp();
      if (rows > 0)
        stage_out(stg, SM::STG_PITCH, reinterpret_cast<uint8_t*>(g.y32 + (size_t)m_base * LN_N + n0 + c * 32), (size_t)LN_N * 4,
                  rows, 128, lane, false);
      __syncwarp();
      if (g.y16 != nullptr) {
        row_write_16<false>(myrow, 0, v);
        __syncwarp();
        if (rows > 0)
          stage_out(stg, SM::STG_PITCH, reinterpret_cast<uint8_t*>(g.y16 + (size_t)m_base * LN_N + n0 + c * 32), (size_t)LN_N * 2,
                    rows, 64, lane, false);
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // nobody exits while a peer can still signal its barriers / write its smem
  if (warp == 1) tmem_dealloc_2sm(tmem_base, BN);
}

#endif  // __CUDACC__

}  // namespace magb200

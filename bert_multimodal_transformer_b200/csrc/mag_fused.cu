// K1: BertEmbeddings + MAG as ONE warp-specialised kernel (bert.py:211-219 -> modeling.py:25-51):
//
//   word/position/type gather + LayerNorm(1e-12) + dropout  ->  the four gate Linears on tcgen05 (TMA-fed, TMEM
//   accumulators)  ->  ReLU gates, h_m = relu(W_hv[v;h]) * W_v v + relu(W_ha[a;h]) * W_a a  ->  row norms,
//   alpha = min(beta |h| / (|h_m| + 1e-6), 1)  ->  LayerNorm(1e-5)(alpha h_m + h)  ->  dropout
//
// One persistent cooperative launch, one CTA per SM, three phases separated by grid barriers:
//   1. rows:  every warp of the grid gathers + normalises token rows (the code of embed_ln_fwd_kernel), writes the
//             fp32 embedding (residual of the shift, saved for backward) and the packed fp16 GEMM operand
//             A[t] = [h | visual | acoustic | 0] (L2-resident: 4.3 MB)
//   2. GEMM:  CTA (m, n) owns 128 tokens x 128 hidden columns of ALL FOUR gates: 4 x 128 = 512 TMEM columns.
//             warp 0 = TMA producer, warp 1 = MMA issuer (tcgen05.mma kind::f16, M 128 x N 256: gates hv|ha over all
//             14 K blocks, gates v|a only over the 2 modality K blocks -- their text weights are structurally zero),
//             warps 4-7 = epilogue: the four accumulators of a column meet in registers, bias + ReLU + the two
//             products give h_m directly; the gate pre-activations go to memory only in training (backward reads them)
//   3. rows:  |h|, |h_m|, alpha, shift, LayerNorm, dropout per token row (the code of mag_tail_fwd_kernel, reading the
//             3 KB h_m row instead of the 12 KB gate row)
// Against the three launches it replaces (embed_ln_fwd -> gate GEMM -> mag_tail_fwd) the 29.5 MB fp32 gate tensor is
// no longer read back in forward (and not written at all in eval): ~40 KB/token moved -> ~17 KB (train) / ~14 KB (eval).
#include <stdlib.h>
#include "common.cuh"
#include "gemm_sm100.cuh"
#include "host.h"
#include "rowdev.cuh"

namespace magb200 {

namespace {

constexpr int CW = 128;                 // hidden columns per CTA tile (x 4 gates = 512 TMEM columns)
constexpr int BM = 128, BK = 64;
constexpr int A_BYTES = BM * 128;       // 16 KB
constexpr int BOX_BYTES = CW * 128;     // one gate's B box: 16 KB
constexpr int STAGE_BYTES = A_BYTES + 4 * BOX_BYTES;   // 80 KB
constexpr int STAGES = 2;
constexpr int STG_PITCH = 32 * 4 + 16;  // staging row: 32 fp32 + bank skew
constexpr int STG_WARP = 32 * STG_PITCH;
constexpr int SMEM_STG = STAGES * STAGE_BYTES;
constexpr int SMEM_BIAS = SMEM_STG + 4 * STG_WARP;            // bias of the tile's 4 x 128 gate columns (per epilogue warp)
constexpr int SMEM_BAR = SMEM_BIAS + 4 * 4 * CW * 4;
constexpr int SMEM_TOTAL = SMEM_BAR + 128 + 1024;

struct MagFusedArgs {
  // phase 1 (BertEmbeddings)
  const int64_t* ids; const int64_t* tt; const float* visual; const float* acoustic;
  const float* word; const float* pos; const float* type; const float* egamma; const float* ebeta;
  float e_eps; uint32_t e_thresh; float e_scale; uint32_t e_site;
  int V, A, KA, T, S, vocab;
  float* emb32; __half* amag;
  // phase 2 (gate GEMM)
  const float* bias4; float* G32; float* hm32;
  int mtiles, ntiles;
  // phase 3 (MAG tail)
  const float* mgamma; const float* mbeta; float beta_shift, m_eps; uint32_t m_thresh; float m_scale; uint32_t m_site;
  float* out32; __half* out16;
  Seed seed;
  unsigned int* barrier;   // two zeroed words
};

// all CTAs of the (cooperative, co-resident) grid meet; writes before the barrier are visible to every thread after it
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int spins = 0;
    while (*reinterpret_cast<volatile unsigned int*>(counter) < nblocks) {
      __nanosleep(64);
      if (++spins > (1u << 26)) { printf("magb200: grid barrier timed out (block %d)\n", blockIdx.x); __trap(); }
    }
    __threadfence();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256, 1)
mag_fused_fwd_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const MagFusedArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SMEM_BAR);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 1);
  const int warp = warp_idx_uniform();
  const int lane = threadIdx.x & 31;
  const uint64_t seed = a.seed.ptr ? *a.seed.ptr : a.seed.value;

  pdl_trigger();   // the QKV GEMM that follows may run its private prologue
  // ---------------------------------------------------------------- private prologue of the GEMM phase
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      mbar_init(tmem_full_bar, 1);
      mbar_init(tmem_empty_bar, 4);          // the four epilogue warps
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }

  // ---------------------------------------------------------------- phase 1: embeddings (one warp per token row)
  const int gwarp = blockIdx.x * 8 + warp, nwarps = gridDim.x * 8;
  for (int t = gwarp; t < a.T; t += nwarps) {
    int64_t id = a.ids[t];
    id = id < 0 ? 0 : (id >= a.vocab ? a.vocab - 1 : id);
    const int64_t ty = a.tt ? a.tt[t] : 0;
    const int s = t % a.S;
    Row r, x;
    load_row(r, a.word + id * H, lane);
    load_row(x, a.pos + (int64_t)s * H, lane);
#pragma unroll
    for (int i = 0; i < 4 * VEC; ++i) r.v[i] += x.v[i];
    load_row(x, a.type + ty * H, lane);
#pragma unroll
    for (int i = 0; i < 4 * VEC; ++i) r.v[i] += x.v[i];
    float mean, rstd;
    row_stats(r, a.e_eps, mean, rstd);
    Row g, b;
    load_row(g, a.egamma, lane);
    load_row(b, a.ebeta, lane);
#pragma unroll
    for (int i = 0; i < 4 * VEC; ++i) r.v[i] = (r.v[i] - mean) * rstd * g.v[i] + b.v[i];
    row_dropout(r, seed, a.e_site, t, lane, a.e_thresh, a.e_scale);
    store_row(r, a.emb32 + (int64_t)t * H, lane);
    __half* arow = a.amag + (int64_t)t * a.KA;
    store_row_f16(r, arow, lane);
    for (int c = lane; c < a.KA - H; c += 32) {
      float v = 0.f;
      if (c < a.V) v = a.visual[(int64_t)t * a.V + c];
      else if (c < a.V + a.A) v = a.acoustic[(int64_t)t * a.A + (c - a.V)];
      arow[H + c] = __float2half_rn(v);
    }
  }
  // the operand rows were written through the generic proxy and are read by TMA (async proxy) in phase 2
  asm volatile("fence.proxy.async.global;" ::: "memory");
  tc_fence_before();
  grid_barrier(a.barrier, gridDim.x);
  tc_fence_after();
  asm volatile("fence.proxy.async.global;" ::: "memory");
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  // ---------------------------------------------------------------- phase 2: gate GEMM, tiles blockIdx.x, + gridDim.x, ...
  // (one tile per CTA at the model's sizes: 19 x 6 = 114 tiles at T = 2400; the loop serves long sequences)
  {
    const int ntot = a.mtiles * a.ntiles;
    const int KB = a.KA / BK;            // 14
    const int KB_TEXT = H / BK;          // 12: K blocks in which only the hv / ha gates have weights
    if (warp == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < ntot; tile += gridDim.x) {
        const int m0 = (tile / a.ntiles) * BM, c0 = (tile % a.ntiles) * CW;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (elect_one_sync()) {
            const bool all4 = kb >= KB_TEXT;
            mbar_arrive_expect_tx(&full_bar[s], A_BYTES + (all4 ? 4 : 2) * BOX_BYTES);
            uint8_t* dst = smem + s * STAGE_BYTES;
            tma_load_2d(dst, &tmA, &full_bar[s], kb * BK, m0);
            tma_load_2d(dst + A_BYTES, &tmB, &full_bar[s], kb * BK, c0);                       // W_hv rows
            tma_load_2d(dst + A_BYTES + BOX_BYTES, &tmB, &full_bar[s], kb * BK, H + c0);        // W_ha rows
            if (all4) {
              tma_load_2d(dst + A_BYTES + 2 * BOX_BYTES, &tmB, &full_bar[s], kb * BK, 2 * H + c0);   // W_v rows
              tma_load_2d(dst + A_BYTES + 3 * BOX_BYTES, &tmB, &full_bar[s], kb * BK, 3 * H + c0);   // W_a rows
            }
          }
          __syncwarp();
        }
      }
    } else if (warp == 1) {
      const uint32_t sb16 = smem_u32(smem) >> 4;
      const uint32_t dhi = smem_desc_hi_sw128(1024);
      constexpr uint32_t idesc = make_idesc(0u, 0u, false, false, 128, 256);
      int it = 0, n = 0;
      for (int tile = blockIdx.x; tile < ntot; tile += gridDim.x, ++n) {
        if (n > 0) {                       // the epilogue warps have drained the previous tile's accumulators
          mbar_wait(tmem_empty_bar, (uint32_t)((n - 1) & 1));
          tc_fence_after();
        }
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t soff = (uint32_t)s * STAGE_BYTES;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t adesc = smem_desc_from(dhi, sb16 + ((soff + k * 32) >> 4), 0);
              const uint64_t b01 = smem_desc_from(dhi, sb16 + ((soff + A_BYTES + k * 32) >> 4), 0);
              tc_mma_f16(tmem_base, adesc, b01, idesc, (kb | k) != 0 ? 1u : 0u);               // hv | ha: columns [0, 256)
              if (kb >= KB_TEXT) {
                const uint64_t b23 = smem_desc_from(dhi, sb16 + ((soff + A_BYTES + 2 * BOX_BYTES + k * 32) >> 4), 0);
                tc_mma_f16(tmem_base + 256, adesc, b23, idesc, (kb > KB_TEXT || k != 0) ? 1u : 0u);   // v | a: [256, 512)
              }
            }
            tc_commit(&empty_bar[s]);
          }
          __syncwarp();
        }
        if (elect_one_sync()) tc_commit(tmem_full_bar);
        __syncwarp();
      }
    } else if (warp >= 4) {
      const int q = warp & 3;
      uint8_t* stg = smem + SMEM_STG + q * STG_WARP;
      uint8_t* myrow = stg + lane * STG_PITCH;
      const uint32_t tw = tmem_base + ((uint32_t)(q * 32) << 16);
      int n = 0;
      for (int tile = blockIdx.x; tile < ntot; tile += gridDim.x, ++n) {
        const int m0 = (tile / a.ntiles) * BM, c0 = (tile % a.ntiles) * CW;
        const int m_base = m0 + q * 32;
        const int rows = min(32, a.T - m_base);
        // this warp's copy of the tile's biases: [gate][128 columns], staged before the accumulator wait
        float* sb = reinterpret_cast<float*>(smem + SMEM_BIAS) + q * 4 * CW;
        for (int i = lane * 4; i < 4 * CW; i += 128)
          *reinterpret_cast<float4*>(sb + i) = __ldg(reinterpret_cast<const float4*>(a.bias4 + (i / CW) * H + c0 + (i % CW)));
        __syncwarp();
        mbar_wait(tmem_full_bar, (uint32_t)(n & 1));
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < CW / 32; ++c) {
          const int col = c0 + c * 32;
          uint32_t r0[32], r1[32];
          float hm[32];
          // gate hv (relu) x projection v
          tmem_ld_32x32(tw + (uint32_t)(c * 32), r0);
          tmem_ld_32x32(tw + (uint32_t)(256 + c * 32), r1);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float gv = __uint_as_float(r0[j]) + sb[c * 32 + j];               // same address in every lane: broadcast
            const float pv = __uint_as_float(r1[j]) + sb[2 * CW + c * 32 + j];
            r0[j] = __float_as_uint(gv);
            r1[j] = __float_as_uint(pv);
            hm[j] = fmaxf(gv, 0.f) * pv;
          }
          if (a.G32 != nullptr && rows > 0) {          // training: the gate pre-activations are saved for backward
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r0[j]);
            row_write_f32(myrow, 0, v);
            __syncwarp();
            stage_out(stg, STG_PITCH, reinterpret_cast<uint8_t*>(a.G32 + (size_t)m_base * 4 * H + col), (size_t)4 * H * 4,
                      rows, 128, lane, false);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r1[j]);
            row_write_f32(myrow, 0, v);
            __syncwarp();
            stage_out(stg, STG_PITCH, reinterpret_cast<uint8_t*>(a.G32 + (size_t)m_base * 4 * H + 2 * H + col),
                      (size_t)4 * H * 4, rows, 128, lane, false);
            __syncwarp();
          }
          // gate ha (relu) x projection a
          tmem_ld_32x32(tw + (uint32_t)(128 + c * 32), r0);
          tmem_ld_32x32(tw + (uint32_t)(384 + c * 32), r1);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float ga = __uint_as_float(r0[j]) + sb[CW + c * 32 + j];
            const float pa = __uint_as_float(r1[j]) + sb[3 * CW + c * 32 + j];
            r0[j] = __float_as_uint(ga);
            r1[j] = __float_as_uint(pa);
            hm[j] = fmaf(fmaxf(ga, 0.f), pa, hm[j]);                                  // modeling.py:27-30
          }
          if (a.G32 != nullptr && rows > 0) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r0[j]);
            row_write_f32(myrow, 0, v);
            __syncwarp();
            stage_out(stg, STG_PITCH, reinterpret_cast<uint8_t*>(a.G32 + (size_t)m_base * 4 * H + H + col), (size_t)4 * H * 4,
                      rows, 128, lane, false);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r1[j]);
            row_write_f32(myrow, 0, v);
            __syncwarp();
            stage_out(stg, STG_PITCH, reinterpret_cast<uint8_t*>(a.G32 + (size_t)m_base * 4 * H + 3 * H + col),
                      (size_t)4 * H * 4, rows, 128, lane, false);
            __syncwarp();
          }
          if (rows > 0) {
            row_write_f32(myrow, 0, hm);
            __syncwarp();
            stage_out(stg, STG_PITCH, reinterpret_cast<uint8_t*>(a.hm32 + (size_t)m_base * H + col), (size_t)H * 4, rows, 128,
                      lane, false);
            __syncwarp();
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_empty_bar);
      }
    }
  }
  grid_barrier(a.barrier + 1, gridDim.x);

  // ---------------------------------------------------------------- phase 3: MAG tail (one warp per token row)
  {
    Row gm, bt;
    load_row(gm, a.mgamma, lane);
    load_row(bt, a.mbeta, lane);
    for (int t = gwarp; t < a.T; t += nwarps) {
      Row text, hm;
      load_row(text, a.emb32 + (int64_t)t * H, lane);
      load_row(hm, a.hm32 + (int64_t)t * H, lane);
      const float em = sqrtf(row_dot(text, text));                                    // modeling.py:32
      const float hm_raw = sqrtf(row_dot(hm, hm));                                    // :33
      const float hm_n = hm_raw == 0.f ? 1.f : hm_raw;                                // :35-36
      const float alpha = fminf(em / (hm_n + 1e-6f) * a.beta_shift, 1.f);             // :38-42
      Row s;
#pragma unroll
      for (int i = 0; i < 4 * VEC; ++i) s.v[i] = alpha * hm.v[i] + text.v[i];         // :43-48
      float mean, rstd;
      row_stats(s, a.m_eps, mean, rstd);
#pragma unroll
      for (int i = 0; i < 4 * VEC; ++i) s.v[i] = (s.v[i] - mean) * rstd * gm.v[i] + bt.v[i];
      row_dropout(s, seed, a.m_site, t, lane, a.m_thresh, a.m_scale);                 // :47-49
      store_row(s, a.out32 + (int64_t)t * H, lane);
      if (a.out16) store_row_f16(s, a.out16 + (int64_t)t * H, lane);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

inline void drop_params(float p, uint32_t& thresh, float& scale) {
  if (p <= 0.f) { thresh = 0; scale = 1.f; return; }
  double t = (double)p * 4294967296.0;
  thresh = t >= 4294967295.0 ? 4294967295u : (uint32_t)t;
  scale = 1.0f / (1.0f - p);
}

}  // namespace

size_t mag_fused_scratch_bytes(int T) { return (size_t)T * H * sizeof(float) + 256; }

// scratch: mag_fused_scratch_bytes(T) bytes (h_m rows + the two grid-barrier words); gates32 may be null (eval)
int mag_fused_fwd(const EmbedArgs& e, const __half* packed16, const float* bias4, const float* mgamma, const float* mbeta,
                  float beta_shift, float m_eps, float p_mag, uint32_t site_mag, float* gates32, void* scratch,
                  float* out32, __half* out16, cudaStream_t st) {
  MAGB_REQUIRE(e.KA == H + 128 && e.V + e.A <= 128, "mag_fused: ka must be 896 and V + A <= 128");
  MAGB_REQUIRE(e.emb32 != nullptr && e.amag != nullptr && scratch != nullptr, "mag_fused: null buffer");
  static int n_sm[16] = {0};
  int dev = 0;
  MAGB_CHECK_CUDA(cudaGetDevice(&dev));
  MAGB_REQUIRE(dev >= 0 && dev < 16, "mag_fused: device index %d out of range", dev);
  if (n_sm[dev] == 0) {
    MAGB_CHECK_CUDA(cudaFuncSetAttribute(mag_fused_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    cudaDeviceProp prop;
    MAGB_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
    int per_sm = 0;
    MAGB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mag_fused_fwd_kernel, 256, SMEM_TOTAL));
    MAGB_REQUIRE(per_sm >= 1, "mag_fused: the kernel does not fit on an SM");
    n_sm[dev] = prop.multiProcessorCount;
  }
  MagFusedArgs a;
  a.ids = e.ids; a.tt = e.tt; a.visual = e.visual; a.acoustic = e.acoustic; a.word = e.word; a.pos = e.pos; a.type = e.type;
  a.egamma = e.gamma; a.ebeta = e.beta; a.e_eps = e.eps; a.e_site = e.site;
  drop_params(e.p, a.e_thresh, a.e_scale);
  a.V = e.V; a.A = e.A; a.KA = e.KA; a.T = e.T; a.S = e.S; a.vocab = e.vocab;
  a.emb32 = e.emb32; a.amag = e.amag;
  a.bias4 = bias4; a.G32 = gates32;
  a.hm32 = reinterpret_cast<float*>(scratch);
  a.barrier = reinterpret_cast<unsigned int*>(reinterpret_cast<uint8_t*>(scratch) + (size_t)e.T * H * sizeof(float));
  a.mtiles = (e.T + BM - 1) / BM; a.ntiles = H / CW;
  a.mgamma = mgamma; a.mbeta = mbeta; a.beta_shift = beta_shift; a.m_eps = m_eps; a.m_site = site_mag;
  drop_params(p_mag, a.m_thresh, a.m_scale);
  a.out32 = out32; a.out16 = out16; a.seed = e.seed;
  const int grid = n_sm[dev];
  CUtensorMap ta, tb;
  int rc = tensor_map_2d(&ta, e.amag, DT_F16, (uint64_t)e.KA, (uint64_t)e.T, (uint64_t)e.KA, BK, BM);
  if (rc) return rc;
  rc = tensor_map_2d(&tb, packed16, DT_F16, (uint64_t)e.KA, (uint64_t)4 * H, (uint64_t)e.KA, BK, CW);
  if (rc) return rc;
  MAGB_CHECK_CUDA(cudaMemsetAsync(a.barrier, 0, 2 * sizeof(unsigned int), st));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = SMEM_TOTAL; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  MAGB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, mag_fused_fwd_kernel, ta, tb, a));
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace magb200

// tcgen05 GEMM for sm_100a:  D[M,N] = A[M,K] * B[N,K]^T  (+ epilogue)
//
//   * operands staged in shared memory by TMA (cp.async.bulk.tensor, SWIZZLE_128B),
//   * tcgen05.mma issued by one elected thread, fp32 accumulators in TMEM,
//   * tcgen05.ld epilogue by 4 warps (one TMEM lane quarter each), fused elementwise epilogues.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue.  One 128 x BN output tile per CTA; 2 CTAs co-reside per SM (3-stage
// ring each) so one CTA's epilogue overlaps the other's main loop.
//
// Operand layouts ("major" = the contiguous dimension in global memory):
//   A K-major : A stored [M][K]  (nn.Linear activations)       A MN-major: stored [K][M]
//   B K-major : B stored [N][K]  (nn.Linear weight [out,in])   B MN-major: stored [K][N]
// MN-major operands let the weight-gradient GEMM  dW[N,K] = dY[T,N]^T * X[T,K]  read dY and X in
// place (reduction over tokens, both operands token-major) with no transposed copies.
#pragma once

#include "common.cuh"

namespace magb200 {

enum GemmEpilogue : int {
  EPI_STORE_F32 = 0,    // out0(f32) = v                 [+ aux(f32) if aux != null]
  EPI_STORE_F16 = 1,    // out0(f16) = v
  EPI_STORE_BF16 = 2,   // out0(bf16) = v
  EPI_GELU_F16 = 3,     // out0(f16) = gelu(v); out1(f16) = v if out1 != null
  EPI_DGELU_F16 = 4,    // out0(f16) = v * gelu'(aux(f16))
  EPI_ATOMIC_F32 = 5,   // out0(f32) += v   (red.global.add; split-K / gradient accumulation)
  EPI_STORE_F32_F16 = 6,  // out0(f32) = v; out1(f16) = v
  EPI_RELU_F32 = 7,     // out0(f32) = max(v, 0)
};
// v = alpha * acc + bias[n]

struct GemmArgs {
  int M, N, K;        // logical problem
  int kb_total;       // number of K blocks
  int kb_per_split;   // K blocks per blockIdx.z
  int n_split;        // tiles with n0 >= n_split start at K block kb_start_hi (block-sparse B, MAG gate GEMM)
  int kb_start_hi;
  int epi;
  const float* bias;
  void* out0;
  void* out1;
  const void* aux;
  int ldo;            // leading dimension (elements) of out0/out1/aux
  float alpha;
  uint32_t idesc;
  unsigned long long* trace;   // debug: per-CTA clock64 stamps (64 slots per CTA) or null
  // optional dropout fused into the GELU / GELU' epilogues (XLNetFeedForward drops the activation):
  // keep-mask = philox(seed, site, (m * N + n) / 4), scale 1/(1-p); thresh == 0 disables it
  uint32_t drop_thresh;
  float drop_scale;
  uint32_t drop_site;
  unsigned long long drop_seed;
  const unsigned long long* drop_seed_ptr;
  int tiles_m, tiles_n, splits;   // persistent kernel: pair-tiles along M, tiles along N, K splits
  int tma_reduce;                 // atomic epilogue: 1 = reduce through the TMA engine (third tensor map = out0), 0 = red.global
  int stream_k;                   // persistent kernel, atomic epilogue only: 1 = the (tile, K block) work units are cut
                                  // into one contiguous, equally long range per cluster (stream-K) instead of whole tiles
  int dbg;   // measurement modes (results are garbage): 1 = MMA issue without waiting for operands,
             // 2 = operand delivery only (no MMA; the consumer releases each stage as soon as it lands)
};

// cached cuTensorMapEncodeTiled wrapper (gemm.cu): 2-D row-major tensor [d1 rows][d0 contiguous], row pitch ld
// elements, box [b1][b0] with b0 * element size = 128 B, SWIZZLE_128B
int tensor_map_2d(CUtensorMap* out, const void* ptr, int dtype, uint64_t d0, uint64_t d1, uint64_t ld, uint32_t b0,
                  uint32_t b1);

template <int EB>
struct GemmTile {
  static constexpr int BM = 128;
  static constexpr int BK = 128 / EB;    // elements per K block (one 128-byte swizzle row)
  static constexpr int UK = 32 / EB;     // K per tcgen05.mma
  static constexpr int ATOM = 128 / EB;  // MN elements per 128-byte row (MN-major)
};

#ifdef __CUDACC__

__device__ __forceinline__ void st_global_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// fp32 reduction of a staged [32 rows][32 columns] box into global memory by the TMA engine
// (cp.reduce.async.bulk.tensor ... .add): the sum happens at L2 and the SM issues ONE instruction per box.  The
// per-lane red.global.add.v4.f32 it replaces is paced by the LSU at about one LANE per clock whatever the vector
// width (measured: 9.6 k cycles for a CTA's 128 x 256 fp32 tile = 13.6 B/clk, against ~35 B/clk for plain stores),
// which made the reduction epilogue -- not the main loop -- the longest phase of the weight-gradient and stream-K
// GEMMs.  Rows / columns outside the tensor are clipped by the TMA unit, so ragged M needs no predication.
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, int c0, int c1, const void* ssrc) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(ssrc))
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the bulk operations of this thread have finished READING shared memory (the staging boxes may be rewritten)
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... have completed (global memory updated): before the CTA exits
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_f16(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ------------------------------------------------------------------------------------------------
// Epilogue.  tcgen05.ld (32x32b) hands thread i of a warp row (lane base + i) of the accumulator, 32
// consecutive columns at a time.  Writing that straight to global memory makes every store
// instruction touch 32 different rows (measured: 5.5k cycles for a 128x128 fp32 tile, as long as the
// whole K=768 main loop).  Instead each epilogue warp owns a [32 rows][BN cols] staging tile in the
// (now idle) pipeline shared memory: threads write their row into it, then the warp copies whole rows
// out with fully coalesced 16-byte accesses (512 B contiguous per instruction).  Auxiliary inputs
// (residual to add, pre-GELU activations) take the same path in reverse.
template <int BN>
struct EpiStage {
  static constexpr int PITCH = BN * 4 + 16;        // bytes per staged row (fp32 worst case + bank skew)
  // one warp's staging region; a multiple of 1024 B so that it can also hold SWIZZLE_128B boxes for the TMA
  // reduction epilogue (BN / 32 boxes of 32 rows x 128 B)
  static constexpr int WARP_BYTES = (32 * PITCH + 1023) / 1024 * 1024;
  static_assert((BN / 32) * 4096 <= WARP_BYTES, "reduction boxes do not fit in the staging region");
};

// global [rows][rowbytes] (pitch gpitch bytes) -> staging, coalesced.  Eight row-slices are requested before the first
// is stored: the one-load-one-store loop this replaces exposed a full global-load latency per slice (32 dependent
// round trips for an fp32 residual tile -- the dgrad GEMMs that accumulate into the residual gradient took 36 us in
// the model against 17 us without the auxiliary input).
__device__ __forceinline__ void stage_in(uint8_t* stg, int pitch, const uint8_t* __restrict__ gsrc, size_t gpitch,
                                         int rows, int rowbytes, int lane, bool serial = false) {
  const int lpr = rowbytes >> 4;                   // lanes per row (16 B each): 8, 16, 32 or 64
  const int rpi = lpr <= 32 ? 32 / lpr : 1;        // rows per instruction
  const int lr = lpr <= 32 ? lane / lpr : 0;
  const int lc0 = lpr <= 32 ? (lane % lpr) << 4 : lane << 4;
  if (serial) {                                    // measurement mode (dbg 3): the original one-slice-at-a-time loop
    for (int c = lc0; c < rowbytes; c += 512)
#pragma unroll 1
      for (int r = lr; r < rows; r += rpi) {
        const uint4 v = *reinterpret_cast<const uint4*>(gsrc + (size_t)r * gpitch + c);
        *reinterpret_cast<uint4*>(stg + r * pitch + c) = v;
      }
    return;
  }
  for (int c = lc0; c < rowbytes; c += 512) {
    for (int r = lr; r < rows; r += 8 * rpi) {
      uint4 v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (r + q * rpi < rows) v[q] = *reinterpret_cast<const uint4*>(gsrc + (size_t)(r + q * rpi) * gpitch + c);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (r + q * rpi < rows) *reinterpret_cast<uint4*>(stg + (r + q * rpi) * pitch + c) = v[q];
    }
  }
}
// staging -> global, coalesced store (or fp32 reduction when `atomic`).  Four row-slices are read from shared
// memory before their stores are issued: a store-after-load chain per slice (what a naive loop compiles to with
// volatile asm stores) reached 16 B/clk/SM, the batched form reaches the ~35 B/clk/SM the store path sustains
// (tools/micro/store_bw.cu).
__device__ __forceinline__ void stage_out(const uint8_t* stg, int pitch, uint8_t* __restrict__ gdst, size_t gpitch,
                                          int rows, int rowbytes, int lane, bool atomic) {
  const int lpr = rowbytes >> 4;
  const int rpi = lpr <= 32 ? 32 / lpr : 1;
  const int lr = lpr <= 32 ? lane / lpr : 0;
  const int lc0 = lpr <= 32 ? (lane % lpr) << 4 : lane << 4;
  for (int c = lc0; c < rowbytes; c += 512) {
    int r = lr;
    for (; r + 3 * rpi < rows; r += 4 * rpi) {
      uint4 v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) v[q] = *reinterpret_cast<const uint4*>(stg + (r + q * rpi) * pitch + c);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint8_t* d = gdst + (size_t)(r + q * rpi) * gpitch + c;
        if (!atomic) st_global_v4(d, v[q].x, v[q].y, v[q].z, v[q].w);
        else red_add_v4(reinterpret_cast<float*>(d), __uint_as_float(v[q].x), __uint_as_float(v[q].y),
                        __uint_as_float(v[q].z), __uint_as_float(v[q].w));
      }
    }
    for (; r < rows; r += rpi) {
      const uint4 v = *reinterpret_cast<const uint4*>(stg + r * pitch + c);
      uint8_t* d = gdst + (size_t)r * gpitch + c;
      if (!atomic) st_global_v4(d, v.x, v.y, v.z, v.w);
      else red_add_v4(reinterpret_cast<float*>(d), __uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z),
                      __uint_as_float(v.w));
    }
  }
}

// acc[32] (this thread's row, 32 consecutive columns) -> v = alpha * acc + bias.  `sb` points at the bias of the
// chunk's first column in SHARED memory (staged by the warp before it waits for the accumulator; a global load here
// sits in the dependent chain of every chunk: ~400 cycles each), or is null.
__device__ __forceinline__ void epi_affine(const GemmArgs& g, const float* sb, const uint32_t (&r)[32], float (&v)[32]) {
  if (sb != nullptr) {
    const float4* b4 = reinterpret_cast<const float4*>(sb);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = b4[j];                     // same address in every lane: one broadcast wavefront
      v[4 * j + 0] = fmaf(__uint_as_float(r[4 * j + 0]), g.alpha, b.x);
      v[4 * j + 1] = fmaf(__uint_as_float(r[4 * j + 1]), g.alpha, b.y);
      v[4 * j + 2] = fmaf(__uint_as_float(r[4 * j + 2]), g.alpha, b.z);
      v[4 * j + 3] = fmaf(__uint_as_float(r[4 * j + 3]), g.alpha, b.w);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * g.alpha;
  }
}
// stage the bias of columns [n0, n0 + ncols) (ncols <= 256, multiple of 32) into this warp's shared-memory slot
__device__ __forceinline__ const float* epi_stage_bias(const GemmArgs& g, float* sbias, int n0, int ncols, int lane) {
  if (g.bias == nullptr) return nullptr;
  for (int i = lane * 4; i < ncols; i += 128) {
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n0 + i < g.N) b = __ldg(reinterpret_cast<const float4*>(g.bias + n0 + i));     // N is a multiple of 32
    *reinterpret_cast<float4*>(sbias + i) = b;
  }
  __syncwarp();
  return sbias;
}
__device__ __forceinline__ void row_write_f32(uint8_t* row, int c, const float (&v)[32]) {
  float4* d = reinterpret_cast<float4*>(row + c * 128);
#pragma unroll
  for (int j = 0; j < 8; ++j) d[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
// the same 32 columns into box c of a SWIZZLE_128B staging area ([box][32 rows][128 B], 16-byte chunk j of row r lives at
// chunk j ^ (r & 7)): the layout the TMA reduction reads, and conflict-free for the row-per-lane writes
__device__ __forceinline__ void row_write_f32_sw(uint8_t* boxes, int lane, int c, const float (&v)[32]) {
  uint8_t* row = boxes + c * 4096 + lane * 128;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(row + ((j ^ (lane & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
template <bool BF16>
__device__ __forceinline__ void row_write_16(uint8_t* row, int c, const float (&v)[32]) {
  uint4* d = reinterpret_cast<uint4*>(row + c * 64);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (BF16) d[j] = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                                pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
    else d[j] = make_uint4(pack_f16(v[8 * j], v[8 * j + 1]), pack_f16(v[8 * j + 2], v[8 * j + 3]),
                           pack_f16(v[8 * j + 4], v[8 * j + 5]), pack_f16(v[8 * j + 6], v[8 * j + 7]));
  }
}

// v[32] = columns [n, n+32) of row m: zero the dropped elements, scale the kept ones
__device__ __forceinline__ void epi_dropout(const GemmArgs& g, int m, int n, float (&v)[32]) {
  const unsigned long long seed = g.drop_seed_ptr ? *g.drop_seed_ptr : g.drop_seed;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint4 u = philox4x32(seed, g.drop_site, ((unsigned long long)m * g.N + n + 4 * j) >> 2);
    v[4 * j] = u.x >= g.drop_thresh ? v[4 * j] * g.drop_scale : 0.f;
    v[4 * j + 1] = u.y >= g.drop_thresh ? v[4 * j + 1] * g.drop_scale : 0.f;
    v[4 * j + 2] = u.z >= g.drop_thresh ? v[4 * j + 2] * g.drop_scale : 0.f;
    v[4 * j + 3] = u.w >= g.drop_thresh ? v[4 * j + 3] * g.drop_scale : 0.f;
  }
}

// One epilogue warp: rows [m_base, m_base+32) x columns [n0, n0+BN) of the output.  `sbias` = epi_stage_bias() result
// for these columns.  TMEM is drained two 32-column chunks per wait (both tcgen05.ld in flight together).
struct EpiCtx {
  int epi; bool aux32, dual16, live, tma_red; int ncols, m_row, n0, pass, lane;
  uint8_t* myrow; uint8_t* stg; const float* sbias;
};
template <int BN>
__device__ __forceinline__ void epi_chunk(const GemmArgs& g, const EpiCtx& x, int c, const uint32_t (&r)[32]) {
  if (!x.live || c * 32 >= x.ncols) return;
  float v[32];
  epi_affine(g, x.sbias ? x.sbias + c * 32 : nullptr, r, v);
  uint8_t* myrow = x.myrow;
  switch (x.epi) {
    case EPI_STORE_F32:
      if (x.aux32) {
        const float4* a4 = reinterpret_cast<const float4*>(myrow + c * 128);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 a = a4[j];
          v[4 * j] += a.x; v[4 * j + 1] += a.y; v[4 * j + 2] += a.z; v[4 * j + 3] += a.w;
        }
      }
      row_write_f32(myrow, c, v);
      break;
    case EPI_ATOMIC_F32:
      if (x.tma_red) row_write_f32_sw(x.stg, x.lane, c, v);
      else row_write_f32(myrow, c, v);
      break;
    case EPI_RELU_F32:
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      row_write_f32(myrow, c, v);
      break;
    case EPI_STORE_F32_F16:
      if (x.pass == 0) row_write_f32(myrow, c, v);
      else row_write_16<false>(myrow, c, v);
      break;
    case EPI_STORE_F16:
      row_write_16<false>(myrow, c, v);
      break;
    case EPI_STORE_BF16:
      row_write_16<true>(myrow, c, v);
      break;
    case EPI_GELU_F16:
      if (x.dual16) {
        // second output: gelu'(v), saved for the backward pass (EPI_DGELU_F16 multiplies by it) -- it shares the
        // tail q and the exponential with gelu(v), so it costs ~5 extra instructions here and saves the whole
        // erf/exp evaluation (2 MUFU + ~20 FMA-pipe instructions per element) in the dgrad epilogue
        uint4* d1 = reinterpret_cast<uint4*>(myrow + BN * 2 + c * 64);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float gp[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float xv = v[8 * j + k];
            float e;
            const float q = gelu_tail(xv, e);
            gp[k] = fmaf(xv * 0.39894228040143268f, e, xv >= 0.f ? 1.0f - q : q);
            v[8 * j + k] = fmaf(-fabsf(xv), q, fmaxf(xv, 0.f));
          }
          d1[j] = make_uint4(pack_f16(gp[0], gp[1]), pack_f16(gp[2], gp[3]), pack_f16(gp[4], gp[5]), pack_f16(gp[6], gp[7]));
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
      }
      if (g.drop_thresh != 0) epi_dropout(g, x.m_row, x.n0 + c * 32, v);
      row_write_16<false>(myrow, c, v);
      break;
    case EPI_DGELU_F16: {
      // aux = gelu'(u) saved by the forward GELU epilogue
      const uint4* u4 = reinterpret_cast<const uint4*>(myrow + c * 64);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 u = u4[j];
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[t]));
          v[8 * j + 2 * t] *= f.x;
          v[8 * j + 2 * t + 1] *= f.y;
        }
      }
      if (g.drop_thresh != 0) epi_dropout(g, x.m_row, x.n0 + c * 32, v);
      row_write_16<false>(myrow, c, v);
    } break;
    default: break;
  }
}

template <int BN>
__device__ __forceinline__ void gemm_epilogue_warp(const GemmArgs& g, const CUtensorMap* tmO, uint32_t tmem_warp, int m_base,
                                                   int n0, uint8_t* stg, int lane, unsigned long long* tr,
                                                   const float* sbias = nullptr) {
  static_assert(BN % 64 == 0, "chunks are drained in pairs");
  constexpr int PITCH = EpiStage<BN>::PITCH;
  const int rows = min(32, g.M - m_base);                   // may be <= 0 for a tail tile
  const int ncols = min(BN, g.N - n0);                      // multiple of 32
  EpiCtx x;
  x.epi = g.epi;
  x.live = rows > 0 && ncols > 0;
  x.ncols = ncols; x.m_row = m_base + lane; x.n0 = n0;
  x.myrow = stg + lane * PITCH;
  x.stg = stg; x.lane = lane;
  x.tma_red = x.epi == EPI_ATOMIC_F32 && g.tma_reduce != 0;    // reduction through the TMA engine (tmO describes out0)
  x.sbias = sbias;
  if (tr) tr[0] = clock64();      // debug stamps: [0] entry, [1] first accumulator chunks in registers, [2] staged, [3] stored
  bulk_wait_read();               // an earlier reduction epilogue of this warp may still be reading the staging tile
  // -------- auxiliary input through the staging tile (coalesced)
  x.aux32 = (x.epi == EPI_STORE_F32) && g.aux != nullptr;
  const bool aux16 = (x.epi == EPI_DGELU_F16);
  if (x.live && (x.aux32 || aux16)) {
    const int es = x.aux32 ? 4 : 2;
    stage_in(stg, PITCH, reinterpret_cast<const uint8_t*>(g.aux) + ((size_t)m_base * g.ldo + n0) * es,
             (size_t)g.ldo * es, rows, ncols * es, lane, g.dbg == 3);
  }
  __syncwarp();
  // GELU with both outputs (g16 and the pre-activation u16 the backward needs): both are fp16, so the two staged
  // tiles fit side by side in the warp's region (2 x BN*2 B <= BN*4 B per row) and one TMEM pass produces both.
  x.dual16 = (x.epi == EPI_GELU_F16 && g.out1 != nullptr);
  const int npass = x.epi == EPI_STORE_F32_F16 ? 2 : 1;
  for (int pass = 0; pass < npass; ++pass) {
    x.pass = pass;
    // which output this pass produces, and its element size
    int es = 4;
    void* out = g.out0;
    if (x.epi == EPI_STORE_F16 || x.epi == EPI_STORE_BF16 || x.epi == EPI_DGELU_F16 || x.epi == EPI_GELU_F16) es = 2;
    if (x.epi == EPI_STORE_F32_F16) { es = pass == 0 ? 4 : 2; out = pass == 0 ? g.out0 : g.out1; }
    const bool interleave = !x.tma_red && g.dbg != 4;      // MAGB200_GEMM_CFG=400: store after the whole tile is staged (A/B)
#pragma unroll 1
    for (int c = 0; c < BN / 32; c += 2) {
      uint32_t r0[32], r1[32];
      tmem_ld_32x32(tmem_warp + (uint32_t)(c * 32), r0);      // warp-collective: executed by every lane
      tmem_ld_32x32(tmem_warp + (uint32_t)(c * 32 + 32), r1);
      tmem_ld_wait();
      if (tr && c == 0) tr[1] = clock64();
      epi_chunk<BN>(g, x, c, r0);
      epi_chunk<BN>(g, x, c + 1, r1);
      if (interleave && x.live && c * 32 < ncols) {
        // hand this 64-column slice to the store path now: its st.global drain while the next slice is read from TMEM and
        // converted (profiles/r02_gemm_persistent_trace.txt: staging 1.1 k + stores 1.65 k cycles per slice, back to back,
        // and the store phase of all 8 warps coincided -- 20 B/clk/SM against the 35 B/clk the store path sustains)
        __syncwarp();
        const int w = min(64, ncols - c * 32);
        stage_out(stg + c * 32 * es, PITCH, reinterpret_cast<uint8_t*>(out) + ((size_t)m_base * g.ldo + n0 + c * 32) * es,
                  (size_t)g.ldo * es, rows, w * es, lane, x.epi == EPI_ATOMIC_F32);
        if (x.dual16)
          stage_out(stg + BN * 2 + c * 64, PITCH, reinterpret_cast<uint8_t*>(g.out1) + ((size_t)m_base * g.ldo + n0 + c * 32) * 2,
                    (size_t)g.ldo * 2, rows, w * 2, lane, false);
      }
    }
    __syncwarp();
    if (tr) tr[2] = clock64();
    if (interleave) {
      // stored slice by slice above
    } else if (x.live && x.tma_red) {
      // every lane staged its own row of each box: publish to the async proxy, then one lane hands the boxes over
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        for (int c = 0; c * 32 < ncols; ++c) tma_reduce_add_2d(tmO, n0 + c * 32, m_base, stg + c * 4096);
        bulk_commit();
      }
    } else if (x.live) {
      stage_out(stg, PITCH, reinterpret_cast<uint8_t*>(out) + ((size_t)m_base * g.ldo + n0) * es, (size_t)g.ldo * es,
                rows, ncols * es, lane, x.epi == EPI_ATOMIC_F32);
      if (x.dual16)
        stage_out(stg + BN * 2, PITCH, reinterpret_cast<uint8_t*>(g.out1) + ((size_t)m_base * g.ldo + n0) * 2,
                  (size_t)g.ldo * 2, rows, ncols * 2, lane, false);
    }
    __syncwarp();
    if (tr) tr[3] = clock64();
  }
}

template <int EB, bool A_MN, bool B_MN, int BN, int STAGES>
struct GemmSmem {
  using T = GemmTile<EB>;
  static constexpr int A_BYTES = T::BM * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int BIAS_OFFSET = BAR_OFFSET + 128;                          // barriers + TMEM slot fit in 128 B
  static constexpr int BIAS_BYTES = 4 * BN * 4;                                 // one bias copy per epilogue warp
  static constexpr int TOTAL = BIAS_OFFSET + BIAS_BYTES + 1024;                 // + alignment slack
  static_assert((2 * STAGES + 1) * 8 + 4 <= 128, "barrier area");
};

// CL = cluster size along N (1 or 2).  With CL = 2 the two CTAs of a cluster compute adjacent N tiles of the same
// M tile and SHARE the A operand: each loads half of the A tile and multicasts it to both (per-SM TMA delivery
// is the main-loop limiter -- profiles/r01_gemm_pipeline_trace.txt -- so A costs each SM 8 KB instead of 16 KB
// per K block).  A stage is then released only when BOTH CTAs' MMAs have read it (empty barriers count CL
// arrivals, tcgen05.commit multicast).
template <int EB, bool A_MN, bool B_MN, int BN, int STAGES, int CL>
__global__ void __launch_bounds__(192, 2)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmO, const GemmArgs g) {
  using T = GemmTile<EB>;
  using SM = GemmSmem<EB, A_MN, B_MN, BN, STAGES>;
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN");
  static_assert((BN & (BN - 1)) == 0, "TMEM allocation must be a power of two");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = warp_idx_uniform();
  const int lane = threadIdx.x & 31;
  unsigned long long* tr = g.trace ? g.trace + 64ull * (blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) : nullptr;
  if (tr && threadIdx.x == 0) tr[0] = clock64();
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * T::BM;

  int kb_begin = blockIdx.z * g.kb_per_split;
  int kb_end = min(kb_begin + g.kb_per_split, g.kb_total);
  if (n0 >= g.n_split) kb_begin = max(kb_begin, g.kb_start_hi);
  const int num_kb = kb_end - kb_begin;   // host guarantees >= 1

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (g.tma_reduce) tma_prefetch_desc(&tmO);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], CL);
      }
      mbar_init(tmem_full_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();   // peer barriers are initialised before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_wait();        // everything above is private to the CTA; operands / bias / aux may come from the previous kernel
  pdl_trigger();
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0;
  constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1);
  if (tr && threadIdx.x == 0) tr[1] = clock64();

  // Both single-thread loops walk the stage ring with the stage index as a compile-time constant (outer loop over
  // ring revolutions, inner loop unrolled over STAGES): shared-memory addresses, barrier addresses and the UMMA
  // descriptors are then "base + constant", one uniform add each, instead of a dependent chain per instruction.
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (whole warp, one lane issues)
    {
      for (int base = 0; base < num_kb; base += STAGES) {
        const uint32_t ph = (base / STAGES) & 1;
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
          if (base + s < num_kb) {
            mbar_wait(&empty_bar[s], ph ^ 1);
            if (elect_one_sync()) {
            mbar_arrive_expect_tx(&full_bar[s], SM::STAGE_BYTES);
            uint8_t* a_dst = smem + s * SM::STAGE_BYTES;
            uint8_t* b_dst = a_dst + SM::A_BYTES;
            const int k0 = (kb_begin + base + s) * T::BK;
            if constexpr (CL == 1) {
              if constexpr (!A_MN) {
                tma_load_2d(a_dst, &tmA, &full_bar[s], k0, m0);
              } else {
#pragma unroll
                for (int i = 0; i < T::BM / T::ATOM; ++i)
                  tma_load_2d(a_dst + i * (T::BK * 128), &tmA, &full_bar[s], m0 + i * T::ATOM, k0);
              }
            } else {
              // this CTA's share of the A tile (rows / MN atoms crank*BM/CL ...), delivered to every CTA of the cluster
              constexpr int SHARE = T::BM / CL;   // A rows per CTA
              if constexpr (!A_MN) {
                tma_load_2d_mc(a_dst + crank * (SHARE * 128), &tmA, &full_bar[s], k0, m0 + crank * SHARE, kMask);
              } else {
#pragma unroll
                for (int i = 0; i < SHARE / T::ATOM; ++i) {
                  const int atom = crank * (SHARE / T::ATOM) + i;
                  tma_load_2d_mc(a_dst + atom * (T::BK * 128), &tmA, &full_bar[s], m0 + atom * T::ATOM, k0, kMask);
                }
              }
            }
            if constexpr (!B_MN) {
              tma_load_2d(b_dst, &tmB, &full_bar[s], k0, n0);
            } else {
#pragma unroll
              for (int i = 0; i < BN / T::ATOM; ++i)
                tma_load_2d(b_dst + i * (T::BK * 128), &tmB, &full_bar[s], n0 + i * T::ATOM, k0);
            }
            if (tr && base + s < 16) tr[2 + base + s] = clock64();
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp, one lane issues)
    {
      const uint32_t sb16 = smem_u32(smem) >> 4;
      const uint32_t dhi = smem_desc_hi_sw128(1024);
      // K-major : 8-row groups 1024 B apart (SBO); K advances 32 B inside the 128-B swizzle row.
      // MN-major: 64-element (128 B) MN atoms BK*128 B apart (LBO), 8-K-row groups 1024 B apart (SBO);
      //           K advances UK rows = UK*128 B.
      constexpr uint32_t A_LBO = A_MN ? T::BK * 128 : 0, B_LBO = B_MN ? T::BK * 128 : 0;
      constexpr uint32_t A_KSTEP = A_MN ? T::UK * 128 : 32, B_KSTEP = B_MN ? T::UK * 128 : 32;
      for (int base = 0; base < num_kb; base += STAGES) {
        const uint32_t ph = (base / STAGES) & 1;
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
          if (base + s < num_kb) {
            if (g.dbg != 1) mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            if (elect_one_sync()) {
            if (tr && base + s < 16) tr[20 + base + s] = clock64();
            if (g.dbg == 2) {
              mbar_arrive(&empty_bar[s]);
              if (tr && base + s < 16) tr[40 + base + s] = clock64();
            } else {
#pragma unroll
            for (int k = 0; k < T::BK / T::UK; ++k) {
              const uint64_t adesc = smem_desc_from(dhi, sb16 + ((s * SM::STAGE_BYTES + k * A_KSTEP) >> 4), A_LBO);
              const uint64_t bdesc =
                  smem_desc_from(dhi, sb16 + ((s * SM::STAGE_BYTES + SM::A_BYTES + k * B_KSTEP) >> 4), B_LBO);
              const uint32_t acc = (s | k) != 0 ? 1u : (uint32_t)(base != 0);
              if constexpr (EB == 2) tc_mma_f16(tmem_base, adesc, bdesc, g.idesc, acc);
              else tc_mma_tf32(tmem_base, adesc, bdesc, g.idesc, acc);
            }
            if constexpr (CL == 1) tc_commit(&empty_bar[s]);  // smem slot reusable once these MMAs have read it
            else tc_commit_mc(&empty_bar[s], kMask);          // ... in every CTA that multicasts into it
            if (tr && base + s < 16) tr[40 + base + s] = clock64();
            }
            }
            __syncwarp();
          }
        }
      }
      if (elect_one_sync()) {
        if (g.dbg == 2) mbar_arrive(tmem_full_bar);
        else tc_commit(tmem_full_bar);    // accumulator complete
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    // split-K: the bias rides on the first split only
    const float* sb = blockIdx.z != 0 ? nullptr
                                      : epi_stage_bias(g, reinterpret_cast<float*>(smem + SM::BIAS_OFFSET) + q * BN, n0, BN, lane);
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    if (tr && threadIdx.x == 64) tr[60] = clock64();
    // the accumulator is complete => every TMA load has landed and every MMA has read its operands:
    // the pipeline stages are idle and serve as this warp's staging tile
    static_assert(4 * EpiStage<BN>::WARP_BYTES <= STAGES * SM::STAGE_BYTES, "staging does not fit in the pipeline smem");
    gemm_epilogue_warp<BN>(g, &tmO, tmem_base + ((uint32_t)(q * 32) << 16), m0 + q * 32, n0,
                           smem + q * EpiStage<BN>::WARP_BYTES, lane, (tr && threadIdx.x == 64) ? tr + 52 : nullptr, sb);
    bulk_wait_all();          // reduction epilogue: the bulk engine may still be reading this CTA's shared memory
    if (tr && threadIdx.x == 64) tr[61] = clock64();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, BN);
  if constexpr (CL > 1) cluster_sync_all();   // do not exit while the peer can still signal our barriers
  if (tr && threadIdx.x == 0) tr[62] = clock64();
}


// ================================================================================================
// 2-CTA kernel (tcgen05 cta_group::2): a cluster of two CTAs on one SM pair computes a 256 x BN tile.  Each CTA
// stages ITS 128 rows of A and HALF of the B tile (BN/2 rows); the leader CTA's single thread issues
// tcgen05.mma.cta_group::2 (M = 256), which reads both CTAs' shared memory and writes each CTA's 128 accumulator
// rows into that CTA's TMEM.  Per K block an SM therefore ingests 16 KB (A) + BN/2*128 B (B) for 128 x BN of MMA
// work -- half the B traffic of the 1-CTA kernel, which is what bounds it (profiles/r01_gemm_pipeline_trace.txt).
//   full_bar  (leader's only): 1 arrival (leader's arrive.expect_tx of both CTAs' bytes) + complete_tx of all loads;
//             the peer's TMA loads signal the leader's barrier (cp.async.bulk.tensor.cta_group::2 + mapa address)
//   empty_bar (one per CTA)  : the leader's tcgen05.commit.cta_group::2 multicast frees the slot in both CTAs
//   tmem_full (one per CTA)  : same multicast commit after the last K block
template <int BN, int STAGES>
struct Gemm2Smem {
  static constexpr int A_BYTES = 128 * 128;
  static constexpr int B_BYTES = (BN / 2) * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int BIAS_OFFSET = BAR_OFFSET + 128;
  static constexpr int BIAS_BYTES = 4 * BN * 4;
  static constexpr int TOTAL = BIAS_OFFSET + BIAS_BYTES + 1024;
  static_assert((2 * STAGES + 1) * 8 + 4 <= 128, "barrier area");
};

template <bool A_MN, bool B_MN, int BN, int STAGES>
__global__ void __launch_bounds__(192, 1)
gemm2_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmO, const GemmArgs g) {
  using T = GemmTile<2>;
  using SM = Gemm2Smem<BN, STAGES>;
  static_assert(BN == 128 || BN == 256, "BN");
  static_assert(4 * EpiStage<BN>::WARP_BYTES <= STAGES * SM::STAGE_BYTES, "staging does not fit in the pipeline smem");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = warp_idx_uniform();
  const int lane = threadIdx.x & 31;
  const uint32_t crank = __shfl_sync(0xffffffffu, cluster_ctarank(), 0);   // 0 = leader; the pair is adjacent along grid x (cluster 2x1x1):
  const bool leader = crank == 0;                    // cta_group::2 pairs CTAs of consecutive cluster rank
  const int m0 = blockIdx.x * T::BM;                 // this CTA's 128 rows (pair: blockIdx.x = 2p, 2p+1)
  const int n0 = blockIdx.y * BN;

  int kb_begin = blockIdx.z * g.kb_per_split;
  int kb_end = min(kb_begin + g.kb_per_split, g.kb_total);
  const int num_kb = kb_end - kb_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (g.tma_reduce) tma_prefetch_desc(&tmO);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      mbar_init(tmem_full_bar, 1);
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
    // ------------------------------------------------------------------ TMA producer (both CTAs; one lane issues)
    {
      const int nb0 = n0 + (int)crank * (BN / 2);      // this CTA's half of the B tile
      for (int base = 0; base < num_kb; base += STAGES) {
        const uint32_t ph = (base / STAGES) & 1;
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
          if (base + s < num_kb) {
            mbar_wait(&empty_bar[s], ph ^ 1);
            if (elect_one_sync()) {
            if (leader) mbar_arrive_expect_tx(&full_bar[s], 2 * SM::STAGE_BYTES);   // bytes of both CTAs
            const uint32_t bar = mapa_shared(&full_bar[s], 0);                       // the leader's barrier
            uint8_t* a_dst = smem + s * SM::STAGE_BYTES;
            uint8_t* b_dst = a_dst + SM::A_BYTES;
            const int k0 = (kb_begin + base + s) * T::BK;
            if constexpr (!A_MN) {
              tma_load_2d_2sm(a_dst, &tmA, bar, k0, m0);
            } else {
#pragma unroll
              for (int i = 0; i < T::BM / T::ATOM; ++i)
                tma_load_2d_2sm(a_dst + i * (T::BK * 128), &tmA, bar, m0 + i * T::ATOM, k0);
            }
            if constexpr (!B_MN) {
              tma_load_2d_2sm(b_dst, &tmB, bar, k0, nb0);
            } else {
#pragma unroll
              for (int i = 0; i < (BN / 2) / T::ATOM; ++i)
                tma_load_2d_2sm(b_dst + i * (T::BK * 128), &tmB, bar, nb0 + i * T::ATOM, k0);
            }
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only; one lane issues)
    if (leader) {
      const uint32_t sb16 = smem_u32(smem) >> 4;
      const uint32_t dhi = smem_desc_hi_sw128(1024);
      constexpr uint32_t A_LBO = A_MN ? T::BK * 128 : 0, B_LBO = B_MN ? T::BK * 128 : 0;
      constexpr uint32_t A_KSTEP = A_MN ? T::UK * 128 : 32, B_KSTEP = B_MN ? T::UK * 128 : 32;
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
              const uint64_t adesc = smem_desc_from(dhi, sb16 + ((s * SM::STAGE_BYTES + k * A_KSTEP) >> 4), A_LBO);
              const uint64_t bdesc =
                  smem_desc_from(dhi, sb16 + ((s * SM::STAGE_BYTES + SM::A_BYTES + k * B_KSTEP) >> 4), B_LBO);
              tc_mma_f16_2sm(tmem_base, adesc, bdesc, g.idesc, (s | k) != 0 ? 1u : (uint32_t)(base != 0));
            }
            tc_commit_2sm(&empty_bar[s]);
            }
            __syncwarp();
          }
        }
      }
      if (elect_one_sync()) tc_commit_2sm(tmem_full_bar);
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ epilogue (both CTAs, warps 2..5)
    const int q = warp & 3;
    const float* sb = blockIdx.z != 0 ? nullptr
                                      : epi_stage_bias(g, reinterpret_cast<float*>(smem + SM::BIAS_OFFSET) + q * BN, n0, BN, lane);
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    gemm_epilogue_warp<BN>(g, &tmO, tmem_base + ((uint32_t)(q * 32) << 16), m0 + q * 32, n0,
                           smem + q * EpiStage<BN>::WARP_BYTES, lane, nullptr, sb);
    bulk_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // both CTAs are done with TMEM and with each other's barriers
  if (warp == 1) tmem_dealloc_2sm(tmem_base, BN);
}

// ================================================================================================
// Persistent 2-CTA kernel.  One cluster (SM pair) per resident slot walks the tile list
//   tile t -> (m pair = t % tiles_m, n tile = (t / tiles_m) % tiles_n, K split = t / (tiles_m * tiles_n)),
// t = cluster id, cluster id + #clusters, ...   The accumulator is double-buffered in TMEM (2 x BN columns), so the
// epilogue of tile i (NEPI warps per CTA, each with a private staging tile) overlaps the main loop of tile i+1, and
// the TMA/MMA ring never drains between tiles.  Why: the main loop of this model's GEMMs is bound by L2->SM
// operand delivery (~6 KB/clk chip-wide, measured with the "delivery only" mode), a 256x256 pair tile needs half
// the operand bytes per flop of four 128x128 tiles -- but only pays off when its long epilogue is hidden.
//   full_bar[s]   (leader's)  : as in the one-shot kernel
//   empty_bar[s]  (per CTA)   : multicast tcgen05.commit
//   tmem_full[b]  (per CTA)   : multicast tcgen05.commit after the last K block of the tile using buffer b
//   tmem_empty[b] (leader's)  : 2 * NEPI arrivals (every epilogue warp of both CTAs) once buffer b has been read
template <int BN, int STAGES, int NEPI>
struct Gemm2pSmem {
  static constexpr int A_BYTES = 128 * 128;
  static constexpr int B_BYTES = (BN / 2) * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SUB = 64;                                   // columns per epilogue sub-block
  static constexpr int STG_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int STG_BYTES = NEPI * EpiStage<SUB>::WARP_BYTES;
  static constexpr int BAR_OFFSET = STG_OFFSET + STG_BYTES;
  static constexpr int BIAS_OFFSET = BAR_OFFSET + 256;
  static constexpr int BIAS_BYTES = 4 * BN * 4;                     // NEPI warps x BN / (NEPI / 4) columns
  static constexpr int TOTAL = BIAS_OFFSET + BIAS_BYTES + 1024;
  static_assert((2 * STAGES + 4) * 8 + 4 <= 256, "barrier area");
};

// Work of one cluster of the persistent kernel as a sequence of segments (pair tile mp, N tile nt, K blocks
// [kb_begin, kb_end)).  Static mode: whole tiles (or whole K splits of tiles), cluster c takes c, c + #clusters, ...
// Stream-K mode: the tiles_m * tiles_n * kb_total (tile, K block) units are numbered tile-major and cluster c takes the
// contiguous range [c * U / C, (c + 1) * U / C): every cluster runs the same number of K blocks (+-1) whatever the
// tile count, at the price of a partial-sum epilogue per range boundary -- which is free for the red.add epilogue.
// At M = 2400 the model's GEMMs have 27 .. 120 pair tiles for 74 clusters, so whole-tile scheduling leaves 20-60 % of
// the machine idle in the last round; the three warp roles walk the same sequence, so no communication is needed.
template <typename F>
__device__ __forceinline__ void for_each_segment(const GemmArgs& g, int cluster_id, int num_clusters, F&& f) {
  // one loop, one call site of f (the body holds the unrolled issue loops); 32-bit arithmetic (the host checks that
  // units * clusters fits); results pass through a shuffle so the compiler keeps them in uniform registers -- the TMA
  // coordinates and loop bounds derived from them feed single-lane issue loops (see warp_idx_uniform)
  const bool sk = g.stream_k != 0;
  const uint32_t tiles_mn = (uint32_t)(g.tiles_m * g.tiles_n);
  const uint32_t units = tiles_mn * (uint32_t)g.kb_total;
  uint32_t u = sk ? units * (uint32_t)cluster_id / (uint32_t)num_clusters : (uint32_t)cluster_id;
  const uint32_t u1 = sk ? units * (uint32_t)(cluster_id + 1) / (uint32_t)num_clusters : tiles_mn * (uint32_t)g.splits;
  while (u < u1) {
    int mp, nt, kb_begin, kb_end;
    uint32_t adv;
    if (sk) {
      const uint32_t tile = u / (uint32_t)g.kb_total;
      kb_begin = (int)(u - tile * (uint32_t)g.kb_total);
      const uint32_t left = u1 - u;
      kb_end = left < (uint32_t)(g.kb_total - kb_begin) ? kb_begin + (int)left : g.kb_total;
      mp = (int)(tile % (uint32_t)g.tiles_m);
      nt = (int)(tile / (uint32_t)g.tiles_m);
      adv = (uint32_t)(kb_end - kb_begin);
    } else {
      mp = (int)(u % (uint32_t)g.tiles_m);
      const uint32_t r = u / (uint32_t)g.tiles_m;
      nt = (int)(r % (uint32_t)g.tiles_n);
      kb_begin = (int)(r / (uint32_t)g.tiles_n) * g.kb_per_split;
      kb_end = min(kb_begin + g.kb_per_split, g.kb_total);
      adv = (uint32_t)num_clusters;
    }
    mp = __shfl_sync(0xffffffffu, mp, 0);
    nt = __shfl_sync(0xffffffffu, nt, 0);
    kb_begin = __shfl_sync(0xffffffffu, kb_begin, 0);
    kb_end = __shfl_sync(0xffffffffu, kb_end, 0);
    f(mp, nt, kb_begin, kb_end);
    u += adv;
  }
}

__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

template <bool A_MN, bool B_MN, int BN, int STAGES, int NEPI>
__global__ void __launch_bounds__(64 + 32 * NEPI, 1)
gemm2p_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmO, const GemmArgs g) {
  using T = GemmTile<2>;
  using SM = Gemm2pSmem<BN, STAGES, NEPI>;
  static_assert(BN == 128 || BN == 256, "BN");
  static_assert(NEPI == 4 || NEPI == 8, "NEPI");
  constexpr int TCOLS = 2 * BN;                        // two accumulator buffers
  constexpr int WCOLS = BN / (NEPI / 4);               // columns one epilogue warp drains per tile
  static_assert(WCOLS % SM::SUB == 0, "epilogue column split");
  static_assert(SM::TOTAL <= 232448, "shared memory");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;            // [2]
  uint64_t* tmem_empty = tmem_full + 2;                // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = warp_idx_uniform();
  const int lane = threadIdx.x & 31;
  const uint32_t crank = __shfl_sync(0xffffffffu, cluster_ctarank(), 0);
  const bool leader = crank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  unsigned long long* tr = g.trace ? g.trace + 64ull * blockIdx.x : nullptr;   // debug: clock64 stamps per CTA
  if (tr && threadIdx.x == 0) tr[0] = clock64();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (g.tma_reduce) tma_prefetch_desc(&tmO);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&tmem_full[b], 1);
        mbar_init(&tmem_empty[b], 2 * NEPI);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_2sm(tmem_slot, TCOLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_wait();
  pdl_trigger();
  if (tr && threadIdx.x == 0) tr[1] = clock64();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs; one lane issues)
    {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for_each_segment(g, cluster_id, num_clusters, [&](int mp, int nt, int kb_begin, int kb_end) {
        const int m0 = (2 * mp + (int)crank) * T::BM;
        const int nb0 = nt * BN + (int)crank * (BN / 2);            // this CTA's half of the B tile
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one_sync()) {
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * SM::STAGE_BYTES);
          const uint32_t bar = mapa_shared(&full_bar[stage], 0);
          uint8_t* a_dst = smem + stage * SM::STAGE_BYTES;
          uint8_t* b_dst = a_dst + SM::A_BYTES;
          const int k0 = kb * T::BK;
          if constexpr (!A_MN) {
            tma_load_2d_2sm(a_dst, &tmA, bar, k0, m0);
          } else {
#pragma unroll
            for (int i = 0; i < T::BM / T::ATOM; ++i)
              tma_load_2d_2sm(a_dst + i * (T::BK * 128), &tmA, bar, m0 + i * T::ATOM, k0);
          }
          if constexpr (!B_MN) {
            tma_load_2d_2sm(b_dst, &tmB, bar, k0, nb0);
          } else {
#pragma unroll
            for (int i = 0; i < (BN / 2) / T::ATOM; ++i)
              tma_load_2d_2sm(b_dst + i * (T::BK * 128), &tmB, bar, nb0 + i * T::ATOM, k0);
          }
          if (tr && it < 12) tr[2 + it] = clock64();
          }
          __syncwarp();
          ++it;
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      });
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only; one lane issues)
    if (leader) {
      const uint32_t sb16 = smem_u32(smem) >> 4;
      const uint32_t dhi = smem_desc_hi_sw128(1024);
      constexpr uint32_t A_LBO = A_MN ? T::BK * 128 : 0, B_LBO = B_MN ? T::BK * 128 : 0;
      constexpr uint32_t A_KSTEP = A_MN ? T::UK * 128 : 32, B_KSTEP = B_MN ? T::UK * 128 : 32;
      int stage = 0;
      uint32_t phase = 0;
      uint32_t acc_it = 0;
      int it = 0;
      const uint32_t peer_empty = mapa_shared(&empty_bar[0], 1), peer_tfull = mapa_shared(&tmem_full[0], 1);
      for_each_segment(g, cluster_id, num_clusters, [&](int, int, int kb_begin, int kb_end) {
        const int num_kb = kb_end - kb_begin;
        const uint32_t buf = acc_it & 1, acc_phase = (acc_it >> 1) & 1;
        mbar_wait(&tmem_empty[buf], acc_phase ^ 1);      // epilogue warps of both CTAs have drained this buffer
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          if (g.dbg != 1) mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one_sync()) {
          if (tr && it < 12) tr[14 + it] = clock64();
          if (g.dbg == 2) {          // delivery only: release the stage in both CTAs without touching the tensor core
            mbar_arrive(&empty_bar[stage]);
            mbar_arrive_cluster(peer_empty + stage * 8);
          } else {
          const uint32_t soff = (uint32_t)stage * SM::STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < T::BK / T::UK; ++k) {
            const uint64_t adesc = smem_desc_from(dhi, sb16 + ((soff + k * A_KSTEP) >> 4), A_LBO);
            const uint64_t bdesc = smem_desc_from(dhi, sb16 + ((soff + SM::A_BYTES + k * B_KSTEP) >> 4), B_LBO);
            tc_mma_f16_2sm(d_tmem, adesc, bdesc, g.idesc, (kb | k) != 0 ? 1u : 0u);
          }
          tc_commit_2sm(&empty_bar[stage]);
          if (tr && it < 12) tr[26 + it] = clock64();
          }
          }
          __syncwarp();
          ++it;
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (elect_one_sync()) {
          if (g.dbg == 2) {
            mbar_arrive(&tmem_full[buf]);
            mbar_arrive_cluster(peer_tfull + buf * 8);
          } else {
            tc_commit_2sm(&tmem_full[buf]);
          }
        }
        __syncwarp();
        ++acc_it;
      });
    }
  } else {
    // ------------------------------------------------------------------ epilogue (both CTAs, warps 2 .. 2+NEPI-1)
    const int ew = warp - 2;
    const int q = warp & 3;                              // TMEM lane quarter this warp may access
    const int cg = ew >> 2;                              // column group (NEPI == 8: two warps share a lane quarter)
    uint8_t* stg = smem + SM::STG_OFFSET + ew * EpiStage<SM::SUB>::WARP_BYTES;
    const uint32_t empty_remote = mapa_shared(&tmem_empty[0], 0);
    uint32_t acc_it = 0;
    for_each_segment(g, cluster_id, num_clusters, [&](int mp, int nt, int kb_begin, int) {
      const uint32_t buf = acc_it & 1, acc_phase = (acc_it >> 1) & 1;
      // the bias belongs to the partial sum that holds the tile's first K block
      const float* sbw = kb_begin != 0 ? nullptr
                                       : epi_stage_bias(g, reinterpret_cast<float*>(smem + SM::BIAS_OFFSET) + ew * WCOLS,
                                                        nt * BN + cg * WCOLS, WCOLS, lane);
      mbar_wait(&tmem_full[buf], acc_phase);
      tc_fence_after();
      // experiment (dbg 4): shift the second column group by about half a sub-block so that its arithmetic overlaps
      // the first group's store phase (the 35 B/clk store path serialises all 8 warps when they store together)
      if (g.dbg == 4 && cg == 1 && acc_it == 0) __nanosleep(1000);
      if (tr && threadIdx.x == 64 && acc_it < 3) tr[38 + 2 * acc_it] = clock64();
      const int m_base = (2 * mp + (int)crank) * T::BM + q * 32;
#pragma unroll 1
      for (int sb = 0; sb < WCOLS / SM::SUB; ++sb) {
        const int col = cg * WCOLS + sb * SM::SUB;
        gemm_epilogue_warp<SM::SUB>(g, &tmO, tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + col, m_base, nt * BN + col,
                                    stg, lane, (tr && threadIdx.x == 64 && acc_it == 0 && sb < 2) ? tr + 44 + 4 * sb : nullptr,
                                    sbw ? sbw + sb * SM::SUB : nullptr);
      }
      tc_fence_before();
      __syncwarp();
      if (tr && threadIdx.x == 64 && acc_it < 3) tr[39 + 2 * acc_it] = clock64();
      if (lane == 0) mbar_arrive_cluster(empty_remote + buf * 8);
      ++acc_it;
    });
    bulk_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2sm(tmem_base, TCOLS);
  if (tr && threadIdx.x == 0) tr[63] = clock64();
}

#endif  // __CUDACC__

}  // namespace magb200

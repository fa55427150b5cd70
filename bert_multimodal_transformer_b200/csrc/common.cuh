// Common device helpers for libmagb200 (sm_100a only): mbarrier / TMA / tcgen05 PTX wrappers,
// warp reductions, the counter-based dropout RNG and small math helpers.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#if defined(__CUDA_ARCH__) && !defined(__CUDA_ARCH_FEAT_SM100_ALL) && !defined(__CUDA_ARCH_FEAT_SM101_ALL) && !defined(__CUDA_ARCH_FEAT_SM103_ALL)
#error "libmagb200 is written for sm_100a (tcgen05/TMEM/TMA); compile with -gencode arch=compute_100a,code=sm_100a"
#endif

namespace magb200 {

// ----------------------------------------------------------------------------------------
// error plumbing (host)
void set_last_error(const char* fmt, ...);
void count_launch();   // every kernel launch of the library bumps a process-wide counter (magb200_launch_count)
#define MAGB_CHECK_CUDA(expr)                                                                   \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      ::magb200::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return (int)_e;                                                                           \
    }                                                                                           \
  } while (0)
#define MAGB_REQUIRE(cond, ...)                                 \
  do {                                                          \
    if (!(cond)) {                                              \
      ::magb200::set_last_error(__VA_ARGS__);                   \
      return -1;                                                \
    }                                                           \
  } while (0)

// Programmatic dependent launch level (MAGB200_PDL): 0 = off, 1 = the GEMMs (default), 2 = also the LayerNorm /
// attention / column-sum kernels between them.  launch_pdl() is only for kernels that call pdl_wait() before their
// first access to global memory.
int pdl_level();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_level() >= 2 ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// cudaFuncSetAttribute is per device: "configured" flags are kept per device index (one process may drive several GPUs)
constexpr int kMaxDevices = 16;
inline int current_device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 0;
  return dev;
}

// ----------------------------------------------------------------------------------------
// element type tags shared between host API and kernels
enum DType : int { DT_F32 = 0, DT_F16 = 1, DT_BF16 = 2 };

#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must become a trap (a CUDA error the host reports), never a hang.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) {  // try_wait itself blocks for a HW-defined slice; this is seconds
      printf("magb200: mbarrier wait timed out (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z,
             threadIdx.x);
      __trap();
    }
  }
}

// ---------------------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// 2-D tiled load global -> shared, completion on an mbarrier. c0 = inner (contiguous) coordinate.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// Multicast variant: the box lands at the same shared-memory offset in every CTA of `cta_mask`, and each of
// those CTAs' mbarrier (same offset) receives the complete_tx.
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
// cta_group::2 variant (a CTA pair working on one UMMA tile): the data lands in THIS CTA's shared memory but the
// mbarrier that receives the complete_tx may live in the peer (leader) CTA.
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
// shared::cluster address of `p` (an address in this CTA's shared memory) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
// ---------------------------------------------------------------------------- warp-uniform single-lane issue
// The TMA / tcgen05 instructions take their operands from UNIFORM registers.  Issued from inside `if (lane == 0)`
// the compiler has to treat every operand as divergent and wraps each instruction in ELECT + 4 x R2UR.BROADCAST +
// a BRA.U.ANY loop (seen in SASS; measured ~200 cycles per TMA load and ~80 per MMA, which made both single-thread
// loops the bottleneck of the main loop).  Instead the whole warp walks the loop in uniform control flow -- warp index
// obtained through a shuffle so the compiler can prove it uniform -- and only the issue itself is predicated on one
// elected lane.
__device__ __forceinline__ int warp_idx_uniform() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------------------- programmatic dependent launch
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while its predecessor in the
// stream is still draining: its CTAs run their private prologue (barrier init, TMEM allocation, descriptor prefetch),
// then block in pdl_wait() until the predecessor has completed and its writes are visible.  pdl_trigger() in the
// predecessor allows that early start once all of ITS CTAs are resident.  Both are no-ops for ordinary launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------------------- clusters
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// ---- cta_group::2 (CTA pair) forms
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// issued by the leader CTA only: completion of its prior tcgen05.mma.cta_group::2 arrives on the barrier at this
// offset in both CTAs of the pair
__device__ __forceinline__ void tc_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}
// D[tmem of both CTAs] (+)= A[256 x K: 128 rows from each CTA's smem] * B[N x K: N/2 rows from each CTA's smem]
__device__ __forceinline__ void tc_mma_f16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// All prior tcgen05.mma of this thread arrive on the mbarrier when they complete
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// same, arriving on the barrier at this offset in every CTA of `cta_mask` (operand-multicast clusters: a smem
// slot is shared work, so its release must be seen by every CTA that writes into it)
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(cta_mask)
               : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; 16-bit operands (f16/bf16 selected by the instruction descriptor)
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32-bit operands read as tf32
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// TMEM -> registers: this warp's 32 lanes x 32 consecutive 32-bit columns (thread i <- lane base+i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (tcgen05), SWIZZLE_128B canonical layouts.
//   bits [0,14)  start address >> 4        bits [16,30) leading-dim byte offset >> 4
//   bits [32,46) stride byte offset >> 4   bits [46,48) descriptor version = 1 (sm_100)
//   bits [61,64) layout type: 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// The same descriptor split into its constant high word and a low word that is (smem address >> 4) plus constant
// fields, so that the issue loop needs ONE uniform add per descriptor: the single MMA-issuing thread is the
// serial resource of the kernel (measured: ~80 cycles per tcgen05.mma when the descriptor is rebuilt with a
// dependent chain of uniform-datapath instructions; profiles/r01_gemm_pipeline_trace.txt).
__device__ __forceinline__ uint32_t smem_desc_hi_sw128(uint32_t sbo_bytes) {
  return ((sbo_bytes >> 4) & 0x3FFF) | (1u << 14) | (2u << 29);     // bits [32,46) SBO, [46,48) version, [61,64) layout
}
__device__ __forceinline__ uint64_t smem_desc_from(uint32_t hi, uint32_t addr16 /* smem byte address >> 4 */,
                                                   uint32_t lbo_bytes) {
  return ((uint64_t)hi << 32) | (uint64_t)(addr16 + (((lbo_bytes >> 4) & 0x3FFF) << 16));
}

// Instruction descriptor for kind::f16 / kind::tf32, fp32 accumulate, dense, no negate.
//   [4,6) D fmt (1 = f32)  [7,10) A fmt  [10,13) B fmt  (f16 = 0, bf16 = 1, tf32 = 2)
//   bit 15 A major (1 = MN)  bit 16 B major (1 = MN)  [17,23) N >> 3  [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc(uint32_t a_fmt, uint32_t b_fmt, bool a_mn, bool b_mn, uint32_t M,
                                                  uint32_t N) {
  return (1u << 4) | (a_fmt << 7) | (b_fmt << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((N >> 3) << 17) | ((M >> 4) << 24);
}

// ---------------------------------------------------------------------------- warp helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------- math
// erf-GELU (transformers 3.0.2 ACT2FN["gelu"] = F.gelu) without erff():  gelu(x) = x Phi(x),
//   q(x) = 0.5 erfc(|x|/sqrt2) = 0.5 (a1 t + ... + a5 t^5) exp(-x^2/2),  t = 1/(1 + p |x|/sqrt2)
// (Abramowitz-Stegun 7.1.26, |erf error| <= 1.5e-7; measured |gelu error| 3.3e-7 in fp32, the same as the erff form),
//   Phi(x) = x >= 0 ? 1 - q : q,  gelu(x) = max(x, 0) - |x| q,  gelu'(x) = Phi(x) + x exp(-x^2/2) / sqrt(2 pi).
// One MUFU.RCP + one MUFU.EX2 + ~10 FMA-pipe instructions; erff() costs ~24 and the GEMM epilogues that apply it to
// every FFN activation are issue-bound (profiles/r01_gemm_persistent_trace.txt).
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// q(x) as above; e = exp(-x^2/2)
__device__ __forceinline__ float gelu_tail(float x, float& e) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = rcp_approx(fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(t, 0.5f * 1.061405429f, 0.5f * -1.453152027f);
  poly = fmaf(t, poly, 0.5f * 1.421413741f);
  poly = fmaf(t, poly, 0.5f * -0.284496736f);
  poly = fmaf(t, poly, 0.5f * 0.254829592f);
  poly *= t;
  e = ex2_approx(z * z * -1.4426950408889634f);
  return poly * e;
}
__device__ __forceinline__ float gelu_erf(float x) {
  float e;
  const float q = gelu_tail(x, e);
  return fmaf(-fabsf(x), q, fmaxf(x, 0.f));
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float e;
  const float q = gelu_tail(x, e);
  const float cdf = x >= 0.f ? 1.0f - q : q;
  return fmaf(x * 0.39894228040143268f, e, cdf);
}

// ---------------------------------------------------------------------------- dropout RNG
// Counter-based, so the backward pass regenerates the forward keep-mask from (seed, site, element index) instead of
// storing it: one call yields 4 uniform 32-bit words for the 4 consecutive elements starting at index 4*ctr.
//
// Default generator: a keyed integer hash per element -- triple32 (Wellons' hash-prospector: 3 multiplies, avalanche bias
// 0.02) of (element index ^ k0), xor k1, with (k0, k1) hashed from (seed, site).  ~12 integer ops per element, against
// ~28 for Philox-4x32-10: ncu showed the row-wise kernels issue-bound ON THE GENERATOR (dropout_residual_ln_fwd: 670
// warp instructions per row of which ~600 were Philox; ln_bwd about half; the attention kernels a third), so this is a
// kernel-time matter, not a taste one.  Dropout needs masks that are uniform, independent across elements and sites
// and reproducible -- not cryptographic quality (tests/test_gpu_ops.py checks rate, determinism and independence).
// -DMAGB200_PHILOX restores Philox-4x32-10 (the generator torch uses) at every site.
__device__ __forceinline__ uint32_t hash_triple32(uint32_t x) {
  x ^= x >> 17; x *= 0xed5ad4bbu;
  x ^= x >> 11; x *= 0xac4c1b51u;
  x ^= x >> 15; x *= 0x31848babu;
  x ^= x >> 14;
  return x;
}
__device__ __forceinline__ uint4 philox4x32_10(uint64_t seed, uint32_t site, uint64_t ctr) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = site, c3 = 0x2545F491u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ uint4 philox4x32(uint64_t seed, uint32_t site, uint64_t ctr) {
#ifdef MAGB200_PHILOX
  return philox4x32_10(seed, site, ctr);
#else
  // per-(seed, site) keys: loop-invariant at every call site, the compiler hoists them
  const uint32_t k0 = hash_triple32((uint32_t)seed ^ hash_triple32((uint32_t)(seed >> 32) ^ (site * 0x9E3779B9u + 0x7F4A7C15u)));
  const uint32_t k1 = hash_triple32(k0 ^ 0x85EBCA6Bu);
  const uint32_t hi = (uint32_t)(ctr >> 30);                       // element index bits >= 32 (tensors > 4 G elements)
  const uint32_t base = ((uint32_t)ctr << 2) ^ k0 ^ (hi ? hash_triple32(hi) : 0u);
  return make_uint4(hash_triple32(base) ^ k1, hash_triple32(base ^ 1u) ^ k1, hash_triple32(base ^ 2u) ^ k1,
                    hash_triple32(base ^ 3u) ^ k1);
#endif
}
// keep iff u32 >= p * 2^32 ; threshold precomputed on the host side of each kernel
__device__ __forceinline__ uint32_t dropout_threshold(float p) {
  return p <= 0.f ? 0u : (uint32_t)fminf(p * 4294967296.0f, 4294967295.0f);
}

#endif  // __CUDACC__

}  // namespace magb200

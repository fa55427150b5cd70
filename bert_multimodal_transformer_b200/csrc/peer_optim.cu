// Data-parallel optimizer step as ONE kernel over NVLink peer memory (no NCCL on this path):
//
//     reduce-scatter of the gradients  +  AdamW  +  all-gather of the updated parameters
//
// Every rank keeps its flat fp32 parameters and its flat fp32 gradients in symmetric memory (same allocation on every
// GPU, peer pointers exchanged once through torch's symmetric-memory rendezvous).  Rank r owns elements
// [slice_begin, slice_end) of the flat buffers (1 / world of the model, 256-aligned) and runs, for each element i:
//
//     g  = (G_0[i] + G_1[i] + ... + G_{W-1}[i]) * grad_unscale        <- P2P loads from every rank's gradient buffer
//     m, v, p: transformers-3.0.2 AdamW (head_optim.cu::adamw_kernel) <- moments exist for the owned slice only
//     P_q[i] = p for every rank q                                     <- P2P stores into every rank's parameters
//     p16[i] = fp16(p)                                                <- local operand mirror (peers recast their copy)
//
// Per rank and step the NVLink carries (W-1)/W of the model in as gradients and out as parameters -- the volume of a
// ring all-reduce -- but the HBM-bound AdamW pass shrinks by the factor W (optimizer state and update are sharded,
// ZeRO-1 style), the sum has a fixed rank order (bit-identical replicas by construction: every element is computed
// once and copied), and no collective library kernel competes with the backward GEMMs for SMs.
// Ordering between ranks is a flag barrier in symmetric memory (peer_barrier_kernel): "all gradients written" before,
// "all parameters written / all gradients read" after.
#include "common.cuh"
#include "host.h"

namespace magb200 {

constexpr int kMaxPeers = 8;
struct PeerPtrs {
  const float* grad[kMaxPeers];
  float* param[kMaxPeers];
};

__device__ __forceinline__ float4 ld_peer_f4(const float* p) {
  float4 v;      // peer memory is never cached stale here: one read per step, straight after a system-scope barrier
  asm volatile("ld.global.relaxed.sys.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_peer_f4(float* p, const float4& v) {
  asm volatile("st.global.relaxed.sys.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

struct AdamConsts {
  float lr, b1, b2, eps, wd, step_size, gscale;
};

// one float4 of the owned slice: gradients of all ranks (already loaded, fixed rank order) -> moments, parameters -> every rank
__device__ __forceinline__ void adamw_peer_element(const PeerPtrs& peers, int world, int rank, const float4* g, float* m,
                                                   float* v, __half* p16, const uint8_t* decay_chunk, int64_t i,
                                                   const AdamConsts& c) {
  float4 gg = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int q = 0; q < kMaxPeers; ++q)
    if (q < world) { gg.x += g[q].x; gg.y += g[q].y; gg.z += g[q].z; gg.w += g[q].w; }
  float4 pp = reinterpret_cast<const float4*>(peers.param[rank])[i];
  float4 mm = reinterpret_cast<float4*>(m)[i];
  float4 vv = reinterpret_cast<float4*>(v)[i];
  const float dec = decay_chunk[i >> 6] ? c.lr * c.wd : 0.f;
  float* P = &pp.x; float* G = &gg.x; float* M = &mm.x; float* V = &vv.x;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float gr = G[k] * c.gscale;
    M[k] = c.b1 * M[k] + (1.f - c.b1) * gr;
    V[k] = c.b2 * V[k] + (1.f - c.b2) * gr * gr;
    P[k] -= c.step_size * M[k] / (sqrtf(V[k]) + c.eps);
    P[k] -= dec * P[k];
  }
  reinterpret_cast<float4*>(m)[i] = mm;
  reinterpret_cast<float4*>(v)[i] = vv;
#pragma unroll
  for (int q = 0; q < kMaxPeers; ++q)
    if (q < world) st_peer_f4(peers.param[q] + 4 * i, pp);
  if (p16) {
    __half2 a = __floats2half2_rn(pp.x, pp.y), b = __floats2half2_rn(pp.z, pp.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    reinterpret_cast<uint2*>(p16)[i] = u;
  }
}

// [begin4, end4): the piece of a bucket this rank owns (reduce + AdamW + broadcast).
// [clean_begin4, clean_end4): a range every rank has finished with (a barrier lies between): the parameters the peers
// wrote into it are recast into the local fp16 operand mirror and the local gradients are cleared for the next step.
__global__ void __launch_bounds__(256) adamw_peer_kernel(const PeerPtrs peers, int world, float* __restrict__ m,
                                                         float* __restrict__ v, __half* __restrict__ p16,
                                                         const uint8_t* __restrict__ decay_chunk, int64_t begin4,
                                                         int64_t end4, int64_t clean_begin4, int64_t clean_end4,
                                                         AdamConsts c, int rank, const float* __restrict__ hyper) {
  if (hyper != nullptr) {
    c.lr = hyper[0];
    c.step_size = hyper[1];
  }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = begin4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  // two elements per trip: 2 * world remote loads in flight per thread (NVLink latency is microseconds)
  for (; i + stride < end4; i += 2 * stride) {
    float4 g0[kMaxPeers], g1[kMaxPeers];
#pragma unroll
    for (int q = 0; q < kMaxPeers; ++q)
      if (q < world) {
        g0[q] = ld_peer_f4(peers.grad[q] + 4 * i);
        g1[q] = ld_peer_f4(peers.grad[q] + 4 * (i + stride));
      }
    adamw_peer_element(peers, world, rank, g0, m, v, p16, decay_chunk, i, c);
    adamw_peer_element(peers, world, rank, g1, m, v, p16, decay_chunk, i + stride, c);
  }
  if (i < end4) {
    float4 g0[kMaxPeers];
#pragma unroll
    for (int q = 0; q < kMaxPeers; ++q)
      if (q < world) g0[q] = ld_peer_f4(peers.grad[q] + 4 * i);
    adamw_peer_element(peers, world, rank, g0, m, v, p16, decay_chunk, i, c);
  }
  float* my_grad = const_cast<float*>(peers.grad[rank]);
  for (int64_t j = clean_begin4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < clean_end4; j += stride) {
    const float4 pp = ld_peer_f4(peers.param[rank] + 4 * j);       // written by its owner over NVLink: not through L1
    __half2 a = __floats2half2_rn(pp.x, pp.y), b = __floats2half2_rn(pp.z, pp.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    if (p16) reinterpret_cast<uint2*>(p16)[j] = u;
    reinterpret_cast<float4*>(my_grad)[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// Barrier between the ranks of one node through flags in symmetric memory.  flags[q] of rank r is written by rank q.
// `epoch` lives in local device memory and is advanced by the kernel itself, so a captured graph can be replayed.
struct PeerFlags {
  unsigned int* flags[kMaxPeers];
};
__global__ void peer_barrier_kernel(const PeerFlags pf, int world, int rank, unsigned int* epoch_counter) {
  __shared__ unsigned int epoch;
  if (threadIdx.x == 0) epoch = *epoch_counter + 1;
  __syncthreads();
  const int q = threadIdx.x;
  if (q < world) {
    __threadfence_system();
    asm volatile("st.global.release.sys.u32 [%0], %1;" ::"l"(pf.flags[q] + rank), "r"(epoch) : "memory");
    unsigned int seen = 0, spins = 0;
    do {
      asm volatile("ld.global.acquire.sys.u32 %0, [%1];" : "=r"(seen) : "l"(pf.flags[rank] + q) : "memory");
      if (seen >= epoch) break;
      __nanosleep(100);
      if (++spins > (1u << 27)) {
        printf("magb200: peer barrier timed out (rank %d waiting for rank %d, epoch %u, saw %u)\n", rank, q, epoch, seen);
        __trap();
      }
    } while (true);
    __threadfence_system();
  }
  __syncthreads();
  if (threadIdx.x == 0) *epoch_counter = epoch;
}

int peer_barrier(unsigned int* const* flag_ptrs, int world, int rank, unsigned int* epoch_counter, cudaStream_t st) {
  MAGB_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "peer_barrier: bad world/rank %d/%d", world, rank);
  PeerFlags pf;
  for (int q = 0; q < kMaxPeers; ++q) pf.flags[q] = q < world ? flag_ptrs[q] : nullptr;
  peer_barrier_kernel<<<1, 32, 0, st>>>(pf, world, rank, epoch_counter);
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int adamw_peer_step(const float* const* grad_ptrs, float* const* param_ptrs, int world, int rank, float* m, float* v,
                    __half* p16, const uint8_t* decay_chunk, int64_t slice_begin, int64_t slice_end, int64_t clean_begin,
                    int64_t clean_end, float lr, float b1, float b2, float eps, float wd, int step, float gscale,
                    const float* hyper_dev, int max_blocks, cudaStream_t st) {
  MAGB_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "adamw_peer: bad world/rank %d/%d", world, rank);
  MAGB_REQUIRE(slice_begin % 256 == 0 && slice_end % 256 == 0 && slice_begin <= slice_end, "adamw_peer: slice [%lld, %lld) must be 256-aligned",
               (long long)slice_begin, (long long)slice_end);
  MAGB_REQUIRE(clean_begin % 256 == 0 && clean_end % 256 == 0 && clean_begin <= clean_end, "adamw_peer: clean-up range [%lld, %lld) must be 256-aligned",
               (long long)clean_begin, (long long)clean_end);
  if (slice_end == slice_begin && clean_end == clean_begin) return 0;
  PeerPtrs pp;
  for (int q = 0; q < kMaxPeers; ++q) {
    pp.grad[q] = q < world ? grad_ptrs[q] : nullptr;
    pp.param[q] = q < world ? param_ptrs[q] : nullptr;
  }
  AdamConsts c;
  c.lr = lr; c.b1 = b1; c.b2 = b2; c.eps = eps; c.wd = wd; c.gscale = gscale;
  c.step_size = lr * sqrtf(1.f - powf(b2, (float)step)) / (1.f - powf(b1, (float)step));
  const int64_t n4 = (slice_end - slice_begin) / 4, c4 = (clean_end - clean_begin) / 4;
  int blocks = (int)(((n4 > c4 ? n4 : c4) + 255) / 256);
  const int cap = max_blocks > 0 ? max_blocks : 148 * 8;
  if (blocks > cap) blocks = cap;
  adamw_peer_kernel<<<blocks, 256, 0, st>>>(pp, world, m, v, p16, decay_chunk, slice_begin / 4, slice_end / 4, clean_begin / 4,
                                            clean_end / 4, c, rank, hyper_dev);
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace magb200

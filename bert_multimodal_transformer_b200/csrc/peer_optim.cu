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
// The trainer issues it bucket by bucket behind the backward pass (every rank owns 1 / world of every bucket), and splits
// the word-embedding matrix by the rows the step touched (RowTable below), so the work left when backward ends is the
// bottom layer group and a few thousand embedding rows.
#include "common.cuh"
#include "host.h"

namespace magb200 {

constexpr int kMaxPeers = 8;
struct PeerPtrs {
  const float* grad[kMaxPeers];
  float* param[kMaxPeers];
};

// Peer loads / stores.  Ordering between ranks comes from the barrier KERNELS around this one (system-scope fences,
// release / acquire flags, and the kernel boundary itself), so the data accesses can be weak; the loads bypass L1
// allocation so that a line of peer memory can never be served stale from an earlier launch.  MAGB200_PEER_STRONG=1
// selects relaxed.sys accesses instead (SASS: LDG / STG.E.128.STRONG.SYS) -- the first version, 320
// GB/s per direction at 2 GPUs.
__device__ __forceinline__ float4 ld_peer_f4(const float* p, bool strong = false) {
  float4 v;
  if (strong)
    asm volatile("ld.global.relaxed.sys.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p)
                 : "memory");
  else
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_peer_f4(float* p, const float4& v, bool strong = false) {
  if (strong)
    asm volatile("st.global.relaxed.sys.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  else
    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// NVSwitch multicast (NVLS): one address that stands for the same offset in every rank's buffer.  A reducing load returns
// the SUM over all ranks, added inside the switch (each GPU's NVLink carries the element once instead of world - 1 times);
// a store is replicated by the switch to every rank, this one included.
__device__ __forceinline__ float4 multimem_ld_sum_f4(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_f4(float* mc, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

struct AdamConsts {
  float lr, b1, b2, eps, wd, step_size, gscale;
  int strong;          // peer accesses as relaxed.sys instead of weak
};

// Row-sparse table inside a bucket (the word-embedding matrix: a step touches at most batch * seq of its 30522 rows).
// any[row] != 0: some rank has a gradient for the row in this step (peer_combine_marks_kernel).  A row nobody touched has
// gradient exactly zero, so its update (moment decay + weight decay) needs no gradient exchange and does not depend on
// the backward pass: mode 1 ("early", issued before backward) updates and broadcasts exactly those rows; mode 2 ("late",
// after backward) handles the touched rows and everything outside the table.  Mode 0: dense, no table.
struct RowTable {
  const uint8_t* any;
  int64_t begin4, n4;      // the table in float4 units of the flat buffer
  int row_len4;
  int mode;
  int clean_what;          // clean-up range: bit 0 = recast the fp16 mirror, bit 1 = clear the gradients
};
enum { ROW_SKIP = 0, ROW_ZERO_GRAD = 1, ROW_FULL = 2 };
__device__ __forceinline__ int classify(const RowTable& t, int64_t i) {
  if (t.mode == 0) return ROW_FULL;
  const int64_t rel = i - t.begin4;
  if (rel < 0 || rel >= t.n4) return t.mode == 1 ? ROW_SKIP : ROW_FULL;
  const bool touched = t.any[rel / t.row_len4] != 0;
  if (t.mode == 1) return touched ? ROW_SKIP : ROW_ZERO_GRAD;
  return touched ? ROW_FULL : ROW_SKIP;
}

// one float4 of the owned slice: summed gradient -> moments, parameters -> every rank
template <int W>
__device__ __forceinline__ void adamw_peer_element(const PeerPtrs& peers, float4 gg, float4 pp, float4 mm, float4 vv,
                                                   float* m, float* v, __half* p16, float dec, int64_t i, const AdamConsts& c) {
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
  if constexpr (W == 0) {
    multimem_st_f4(peers.param[0] + 4 * i, pp);          // param[0] = the multicast address
  } else {
#pragma unroll
    for (int q = 0; q < W; ++q) st_peer_f4(peers.param[q] + 4 * i, pp, c.strong != 0);
  }
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
// wrote into it are recast into the local fp16 operand mirror and the local gradients are cleared for the next step;
// [keep_begin4, keep_end4) inside it is this rank's own piece, whose mirror the update already wrote.
// W = number of ranks (static: the per-rank pointers and the loaded gradients stay in registers), U = float4 elements per
// thread and trip.  All U * W remote loads and the 3 * U local ones of a trip are issued before the first is used: NVLink
// latency is microseconds, and the first version (2 elements, loads of p / m / v only after the gradients had arrived)
// reached 320 GB/s per direction at 2 GPUs.
template <int W, int U>
__global__ void __launch_bounds__(256) adamw_peer_kernel(const PeerPtrs peers, float* __restrict__ m, float* __restrict__ v,
                                                         __half* __restrict__ p16, const uint8_t* __restrict__ decay_chunk,
                                                         int64_t begin4, int64_t end4, int64_t clean_begin4,
                                                         int64_t clean_end4, int64_t keep_begin4, int64_t keep_end4,
                                                         AdamConsts c, const float* __restrict__ my_param,
                                                         float* __restrict__ my_grad, const float* __restrict__ hyper,
                                                         const RowTable table) {
  // my_param / my_grad = this rank's entries of `peers` (separate arguments: indexing the by-value struct with a run-time
  // rank would copy it to local memory)
  if (hyper != nullptr) {
    c.lr = hyper[0];
    c.step_size = hyper[1];
  }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = begin4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < end4; i0 += U * stride) {
    int kind[U];
    constexpr int NG = W == 0 ? 1 : W;        // W == 0: multicast form, grad[0] / param[0] are the multicast addresses
    float4 pp[U], mm[U], vv[U], g[U][NG];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      kind[u] = i < end4 ? classify(table, i) : ROW_SKIP;
      if (kind[u] != ROW_SKIP) {
        pp[u] = reinterpret_cast<const float4*>(my_param)[i];
        mm[u] = reinterpret_cast<const float4*>(m)[i];
        vv[u] = reinterpret_cast<const float4*>(v)[i];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if constexpr (W == 0) {
        g[u][0] = kind[u] == ROW_FULL ? multimem_ld_sum_f4(peers.grad[0] + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
#pragma unroll
        for (int q = 0; q < W; ++q)
          g[u][q] = kind[u] == ROW_FULL ? ld_peer_f4(peers.grad[q] + 4 * i, c.strong != 0) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (kind[u] == ROW_SKIP) continue;
      const int64_t i = i0 + u * stride;
      float4 gg = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int q = 0; q < NG; ++q) { gg.x += g[u][q].x; gg.y += g[u][q].y; gg.z += g[u][q].z; gg.w += g[u][q].w; }  // fixed rank order
      const float dec = decay_chunk[i >> 6] ? c.lr * c.wd : 0.f;
      adamw_peer_element<W>(peers, gg, pp[u], mm[u], vv[u], m, v, p16, dec, i, c);
    }
  }
  for (int64_t j = clean_begin4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < clean_end4; j += stride) {
    if (table.mode != 0) {
      // inside the table only touched rows hold gradients, and nothing reads the table's fp16 mirror
      const int64_t rel = j - table.begin4;
      if (rel >= 0 && rel < table.n4) {
        if ((table.clean_what & 2) && table.any[rel / table.row_len4] != 0)
          reinterpret_cast<float4*>(my_grad)[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        continue;
      }
    }
    if ((table.clean_what & 1) && p16 && !(j >= keep_begin4 && j < keep_end4)) {
      const float4 pp = ld_peer_f4(my_param + 4 * j, c.strong != 0);   // written by its owner over NVLink: not through L1
      __half2 a = __floats2half2_rn(pp.x, pp.y), b = __floats2half2_rn(pp.z, pp.w);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&a);
      u.y = *reinterpret_cast<uint32_t*>(&b);
      reinterpret_cast<uint2*>(p16)[j] = u;
    }
    if (table.clean_what & 2) reinterpret_cast<float4*>(my_grad)[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// marks[id] = (*counter + 1) for every token id of the step's batch; the counter is advanced by the kernel itself (one
// block, so a captured graph can be replayed).  marks live in symmetric memory: the peers read them.
__global__ void __launch_bounds__(1024) peer_mark_rows_kernel(const int64_t* __restrict__ ids, int n, int vocab,
                                                              unsigned int* __restrict__ marks, unsigned int* counter) {
  __shared__ unsigned int value;
  if (threadIdx.x == 0) value = *counter + 1;
  __syncthreads();
  const unsigned int val = value;
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    const int64_t id = ids[t];
    if (id >= 0 && id < vocab) marks[id] = val;
  }
  __syncthreads();
  if (threadIdx.x == 0) *counter = val;
}
struct PeerMarks {
  const unsigned int* mark[kMaxPeers];
};
// any[row] = 1 if some rank marked the row in this step (call after a barrier that follows every rank's marking)
__global__ void peer_combine_marks_kernel(const PeerMarks pm, int world, const unsigned int* __restrict__ counter, int vocab,
                                          uint8_t* __restrict__ any) {
  const unsigned int val = *counter;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < vocab; r += gridDim.x * blockDim.x) {
    bool hit = false;
#pragma unroll
    for (int q = 0; q < kMaxPeers; ++q)
      if (q < world) {
        unsigned int x;
        asm volatile("ld.global.relaxed.sys.u32 %0, [%1];" : "=r"(x) : "l"(pm.mark[q] + r) : "memory");
        hit |= x == val;
      }
    any[r] = hit ? 1 : 0;
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

int peer_mark_rows(const int64_t* ids, int n, int vocab, unsigned int* marks, unsigned int* counter, cudaStream_t st) {
  MAGB_REQUIRE(n >= 0 && vocab > 0, "peer_mark_rows: bad sizes %d / %d", n, vocab);
  peer_mark_rows_kernel<<<1, 1024, 0, st>>>(ids, n, vocab, marks, counter);
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int peer_combine_marks(const unsigned int* const* mark_ptrs, int world, const unsigned int* counter, int vocab, uint8_t* any,
                       cudaStream_t st) {
  MAGB_REQUIRE(world >= 1 && world <= kMaxPeers, "peer_combine_marks: bad world %d", world);
  PeerMarks pm;
  for (int q = 0; q < kMaxPeers; ++q) pm.mark[q] = q < world ? mark_ptrs[q] : nullptr;
  peer_combine_marks_kernel<<<(vocab + 255) / 256, 256, 0, st>>>(pm, world, counter, vocab, any);
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int adamw_peer_step(const float* const* grad_ptrs, float* const* param_ptrs, int world, int rank, float* m, float* v,
                    __half* p16, const uint8_t* decay_chunk, int64_t slice_begin, int64_t slice_end, int64_t clean_begin,
                    int64_t clean_end, float lr, float b1, float b2, float eps, float wd, int step, float gscale,
                    const float* hyper_dev, int max_blocks, const uint8_t* row_any, int row_mode, int64_t table_begin,
                    int table_rows, int row_len, int clean_what, int64_t keep_begin, int64_t keep_end, const float* mc_grad,
                    float* mc_param, cudaStream_t st) {
  MAGB_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "adamw_peer: bad world/rank %d/%d", world, rank);
  MAGB_REQUIRE(slice_begin % 256 == 0 && slice_end % 256 == 0 && slice_begin <= slice_end, "adamw_peer: slice [%lld, %lld) must be 256-aligned",
               (long long)slice_begin, (long long)slice_end);
  MAGB_REQUIRE(clean_begin % 256 == 0 && clean_end % 256 == 0 && clean_begin <= clean_end, "adamw_peer: clean-up range [%lld, %lld) must be 256-aligned",
               (long long)clean_begin, (long long)clean_end);
  MAGB_REQUIRE(keep_begin % 4 == 0 && keep_end % 4 == 0 && keep_begin <= keep_end, "adamw_peer: bad keep range [%lld, %lld)",
               (long long)keep_begin, (long long)keep_end);
  MAGB_REQUIRE(clean_what >= 0 && clean_what <= 3, "adamw_peer: clean_what %d (bit 0 recast, bit 1 clear)", clean_what);
  MAGB_REQUIRE(row_mode == 0 || (row_any != nullptr && row_mode >= 1 && row_mode <= 2 && table_begin % 4 == 0 && row_len % 4 == 0 &&
                                 row_len > 0 && table_rows > 0),
               "adamw_peer: bad row table (mode %d, begin %lld, rows %d, row length %d)", row_mode, (long long)table_begin,
               table_rows, row_len);
  if (slice_end == slice_begin && clean_end == clean_begin) return 0;
  PeerPtrs pp;
  for (int q = 0; q < kMaxPeers; ++q) {
    pp.grad[q] = q < world ? grad_ptrs[q] : nullptr;
    pp.param[q] = q < world ? param_ptrs[q] : nullptr;
  }
  AdamConsts c;
  c.lr = lr; c.b1 = b1; c.b2 = b2; c.eps = eps; c.wd = wd; c.gscale = gscale;
  static int strong = -1;
  if (strong < 0) {
    const char* e = getenv("MAGB200_PEER_STRONG");
    strong = (e && *e && atoi(e) != 0) ? 1 : 0;
  }
  c.strong = strong;
  c.step_size = lr * sqrtf(1.f - powf(b2, (float)step)) / (1.f - powf(b1, (float)step));
  RowTable t;
  t.any = row_any; t.mode = row_mode; t.begin4 = table_begin / 4; t.row_len4 = row_mode ? row_len / 4 : 1;
  t.n4 = row_mode ? (int64_t)table_rows * (row_len / 4) : 0;
  t.clean_what = clean_what;
  const int64_t n4 = (slice_end - slice_begin) / 4, c4 = (clean_end - clean_begin) / 4;
  int blocks = (int)(((n4 > c4 ? n4 : c4) + 255) / 256);
  const int cap = max_blocks > 0 ? max_blocks : 148 * 8;
  if (blocks > cap) blocks = cap;
  // the owner's own piece inside the clean-up range keeps the fp16 mirror its update wrote
  const int64_t keep_b = (clean_end > clean_begin) ? keep_begin / 4 : 0, keep_e = (clean_end > clean_begin) ? keep_end / 4 : 0;
#define MAGB_PEER_LAUNCH(W_, U_)                                                                                              \
  adamw_peer_kernel<W_, U_><<<blocks, 256, 0, st>>>(pp, m, v, p16, decay_chunk, slice_begin / 4, slice_end / 4, clean_begin / 4, \
                                                    clean_end / 4, keep_b, keep_e, c, pp.param[rank], const_cast<float*>(pp.grad[rank]),  \
                                                    hyper_dev, t)
  if (mc_grad != nullptr && mc_param != nullptr && world > 1) {
    const float* my_p = pp.param[rank];
    const float* my_g = pp.grad[rank];
    pp.grad[0] = mc_grad;                       // multicast form: slot 0 carries the multicast addresses,
    pp.param[0] = mc_param;
    adamw_peer_kernel<0, 4><<<blocks, 256, 0, st>>>(pp, m, v, p16, decay_chunk, slice_begin / 4, slice_end / 4, clean_begin / 4,
                                                    clean_end / 4, keep_b, keep_e, c, my_p, const_cast<float*>(my_g), hyper_dev, t);
    MAGB_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return 0;
  }
  switch (world) {
    case 1: MAGB_PEER_LAUNCH(1, 4); break;
    case 2: MAGB_PEER_LAUNCH(2, 4); break;
    case 3: MAGB_PEER_LAUNCH(3, 4); break;
    case 4: MAGB_PEER_LAUNCH(4, 4); break;
    case 5: MAGB_PEER_LAUNCH(5, 2); break;
    case 6: MAGB_PEER_LAUNCH(6, 2); break;
    case 7: MAGB_PEER_LAUNCH(7, 2); break;
    default: MAGB_PEER_LAUNCH(8, 2); break;
  }
#undef MAGB_PEER_LAUNCH
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace magb200

// Host side of the tcgen05 GEMM: TMA descriptor construction (cached), kernel selection, launch.
#include "gemm_sm100.cuh"
#include "gemm_ln_sm100.cuh"
#include "host.h"

#include <mutex>
#include <unordered_map>
#include <cstring>

namespace magb200 {

// ---------------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled is a driver entry point; fetch it through the runtime so the library
// has no link-time dependency on libcuda.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  uint64_t d0, d1, ld;
  uint32_t b0, b1, dt;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && ld == o.ld && b0 == o.b0 && b1 == o.b1 && dt == o.dt;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    auto mix = [&h](uint64_t v) { h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); };
    mix(k.d0); mix(k.d1); mix(k.ld); mix(k.b0); mix(k.b1); mix(k.dt);
    return h;
  }
};

// 2-D row-major tensor [d1 rows][d0 contiguous], row pitch ld elements, box [b1][b0], SWIZZLE_128B.
static int get_tensor_map(CUtensorMap* out, const void* ptr, int dtype, uint64_t d0, uint64_t d1, uint64_t ld,
                          uint32_t b0, uint32_t b1) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key{ptr, d0, d1, ld, b0, b1, (uint32_t)dtype};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeTiledFn enc = get_encode_fn();
  MAGB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  const uint32_t eb = dtype == DT_F32 ? 4 : 2;
  MAGB_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA operand %p is not 16-byte aligned", ptr);
  MAGB_REQUIRE((ld * eb) % 16 == 0, "TMA operand row pitch %llu B is not a multiple of 16", (unsigned long long)(ld * eb));
  MAGB_REQUIRE(b0 * eb == 128, "TMA box inner extent must be 128 bytes");
  CUtensorMapDataType cdt = dtype == DT_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                          : dtype == DT_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                            : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  cuuint64_t gdim[2] = {d0, d1};
  cuuint64_t gstr[1] = {ld * eb};
  cuuint32_t box[2] = {b0, b1};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, cdt, 2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MAGB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) ptr=%p dims=%llux%llu ld=%llu box=%ux%u", (int)r,
               ptr, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)ld, b0, b1);
  {
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() > 4096) cache.clear();
    cache.emplace(key, m);
  }
  *out = m;
  return 0;
}

int tensor_map_2d(CUtensorMap* out, const void* ptr, int dtype, uint64_t d0, uint64_t d1, uint64_t ld, uint32_t b0,
                  uint32_t b1) {
  return get_tensor_map(out, ptr, dtype, d0, d1, ld, b0, b1);
}

// Programmatic dependent launch for every GEMM (the kernels call griddepcontrol.wait after their private prologue);
// MAGB200_PDL=0 turns the attribute off.
static int pdl_attr(cudaLaunchAttribute* a) {
  if (pdl_level() < 1) return 0;
  a->id = cudaLaunchAttributeProgrammaticStreamSerialization;
  a->val.programmaticStreamSerializationAllowed = 1;
  return 1;
}

template <int EB, bool A_MN, bool B_MN, int BN, int STAGES, int CL>
static int launch_variant(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const GemmArgs& g, dim3 grid,
                          cudaStream_t st) {
  using SM = GemmSmem<EB, A_MN, B_MN, BN, STAGES>;
  auto kern = gemm_tcgen05_kernel<EB, A_MN, B_MN, BN, STAGES, CL>;
  static bool configured[kMaxDevices] = {false};
  const int dev = current_device_slot();
  if (!configured[dev]) {
    MAGB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
    configured[dev] = true;
  }
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = SM::TOTAL; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (CL > 1) {
      attr[na].id = cudaLaunchAttributeClusterDimension;
      attr[na].val.clusterDim.x = CL; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
      ++na;
    }
    na += pdl_attr(attr + na);
    cfg.attrs = attr; cfg.numAttrs = na;
    MAGB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, to, g));
  }
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

template <int EB, int BN, int STAGES, int CL>
static int launch_majors(bool a_mn, bool b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to,
                         const GemmArgs& g, dim3 grid, cudaStream_t st) {
  if (!a_mn && !b_mn) return launch_variant<EB, false, false, BN, STAGES, CL>(ta, tb, to, g, grid, st);
  if (!a_mn && b_mn) return launch_variant<EB, false, true, BN, STAGES, CL>(ta, tb, to, g, grid, st);
  if (a_mn && !b_mn) return launch_variant<EB, true, false, BN, STAGES, CL>(ta, tb, to, g, grid, st);
  return launch_variant<EB, true, true, BN, STAGES, CL>(ta, tb, to, g, grid, st);
}

// debug/tuning hooks (tools/gemm_tune.py): trace buffer and tile-configuration override
static unsigned long long* g_trace = nullptr;
static int g_cfg = 0;
static int g_dbg = 0;
static int g_policy = 1;   // roles routed to the persistent pair kernel by default: forward Linears (profiles/r01_gemm_policy.txt)
static int g_streamk_grid = 1;
static int g_streamk_min = 0;
static int g_streamk = 1;  // atomic-epilogue problems (weight gradients, accumulate-into-residual dgrads, split-K forward
                           // Linears) go to the persistent pair kernel with stream-K work ranges; MAGB200_STREAMK=0 restores
                           // the round-1 policy (one-shot pair tiles / whole-tile persistent scheduling)
void gemm_set_debug(unsigned long long* trace, int cfg) { g_trace = trace; g_cfg = cfg % 100; g_dbg = cfg / 100; }

// ---------------------------------------------------------------------------------------------
// 2-CTA (cta_group::2) path
template <bool A_MN, bool B_MN, int BN, int STAGES>
static int launch2_variant(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const GemmArgs& g, dim3 grid,
                           cudaStream_t st) {
  using SM = Gemm2Smem<BN, STAGES>;
  auto kern = gemm2_tcgen05_kernel<A_MN, B_MN, BN, STAGES>;
  static bool configured[kMaxDevices] = {false};
  const int dev = current_device_slot();
  if (!configured[dev]) {
    MAGB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
    configured[dev] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = SM::TOTAL; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1 + pdl_attr(attr + 1);
  MAGB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, to, g));
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
template <int BN, int STAGES>
static int launch2_majors(bool a_mn, bool b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to,
                          const GemmArgs& g, dim3 grid, cudaStream_t st) {
  if (!a_mn && !b_mn) return launch2_variant<false, false, BN, STAGES>(ta, tb, to, g, grid, st);
  if (!a_mn && b_mn) return launch2_variant<false, true, BN, STAGES>(ta, tb, to, g, grid, st);
  if (a_mn && !b_mn) return launch2_variant<true, false, BN, STAGES>(ta, tb, to, g, grid, st);
  return launch2_variant<true, true, BN, STAGES>(ta, tb, to, g, grid, st);
}

// Persistent 2-CTA kernel: as many clusters as fit on the device at once (queried per variant), never more than tiles.
template <bool A_MN, bool B_MN, int BN, int STAGES, int NEPI>
static int launch2p_variant(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const GemmArgs& g,
                            cudaStream_t st) {
  using SM = Gemm2pSmem<BN, STAGES, NEPI>;
  auto kern = gemm2p_tcgen05_kernel<A_MN, B_MN, BN, STAGES, NEPI>;
  static int max_clusters[16] = {0};
  int dev = 0;
  MAGB_CHECK_CUDA(cudaGetDevice(&dev));
  MAGB_REQUIRE(dev >= 0 && dev < 16, "gemm: device index %d out of range", dev);
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(64 + 32 * NEPI); cfg.dynamicSmemBytes = SM::TOTAL; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1 + pdl_attr(attr + 1);
  if (max_clusters[dev] == 0) {
    MAGB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
    cudaDeviceProp prop;
    MAGB_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
    cfg.gridDim = dim3(2 * (prop.multiProcessorCount / 2), 1, 1);
    int n = 0;
    MAGB_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
    MAGB_REQUIRE(n >= 1, "gemm: the persistent 2-CTA kernel does not fit on this device");
    max_clusters[dev] = n < prop.multiProcessorCount / 2 ? n : prop.multiProcessorCount / 2;
  }
  // whole-tile scheduling: never more clusters than tiles.  Stream-K: the (tile, K block) units are cut evenly over the
  // clusters of the launch (a range boundary costs a partial-sum epilogue, so >= 4 K blocks per cluster).
  // Stream-K grid.  Two stream-K GEMMs run side by side in backward (dgrad on the main stream, wgrad on the side stream)
  // and a persistent cluster owns its SM pair, so "every cluster of the device" for each of them serialises the two
  // (measured: 3.70 -> 4.08 ms / step).  Default (MAGB200_STREAMK_GRID=1): as many clusters as output tiles, but at least
  // MAGB200_STREAMK_MIN (the 9-tile weight gradient of the attention output Linear would otherwise crawl on 9 clusters);
  // 0 = every cluster, n > 1 = at most n clusters.
  long long total = (long long)g.tiles_m * g.tiles_n * g.splits;
  int cap = max_clusters[dev];
  if (g.stream_k) {
    long long units = (long long)g.tiles_m * g.tiles_n * g.kb_total / 4;       // >= 4 K blocks per cluster
    if (units < 1) units = 1;
    if (g_streamk_grid == 1) {
      if (total < g_streamk_min) total = g_streamk_min;
      if (total > units) total = units;
    } else {
      total = units;
      if (g_streamk_grid > 1 && g_streamk_grid < cap) cap = g_streamk_grid;
    }
  }
  const int ncl = total < cap ? (int)total : cap;
  cfg.gridDim = dim3(2 * ncl, 1, 1);
  MAGB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, to, g));
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
template <int BN, int STAGES, int NEPI>
static int launch2p_majors(bool a_mn, bool b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to,
                           const GemmArgs& g, cudaStream_t st) {
  if (!a_mn && !b_mn) return launch2p_variant<false, false, BN, STAGES, NEPI>(ta, tb, to, g, st);
  if (!a_mn && b_mn) return launch2p_variant<false, true, BN, STAGES, NEPI>(ta, tb, to, g, st);
  if (a_mn && !b_mn) return launch2p_variant<true, false, BN, STAGES, NEPI>(ta, tb, to, g, st);
  return launch2p_variant<true, true, BN, STAGES, NEPI>(ta, tb, to, g, st);
}

// Tensor map of the fp32 output for the TMA reduction epilogue (boxes of 32 rows x 32 columns, SWIZZLE_128B); for the
// other epilogues the kernels never touch the third descriptor and a copy of A's is passed.
static int g_tma_reduce = 1;   // MAGB200_TMA_REDUCE=0: per-lane red.global.add.v4.f32 (the round-1 epilogue)
static int output_map(const GemmProblem& p, const CUtensorMap& fallback, CUtensorMap* to, GemmArgs* g) {
  g->tma_reduce = 0;
  *to = fallback;
  if (p.epi != EPI_ATOMIC_F32 || !g_tma_reduce) return 0;
  if ((reinterpret_cast<uintptr_t>(p.out0) & 15) != 0 || (p.ldo * 4) % 16 != 0) return 0;
  int rc = get_tensor_map(to, p.out0, DT_F32, (uint64_t)p.N, (uint64_t)p.M, (uint64_t)p.ldo, 32, 32);
  if (rc) return rc;
  g->tma_reduce = 1;
  return 0;
}

static int gemm_launch_2cta(const GemmProblem& p, int BN, cudaStream_t st, bool persistent = false, bool stream_k = false) {
  const int BK = 64, ATOM = 64, BM = 128;
  CUtensorMap ta, tb;
  int rc;
  if (!p.a_mn) rc = get_tensor_map(&ta, p.A, p.a_dtype, p.K, p.M, p.lda, BK, BM);
  else rc = get_tensor_map(&ta, p.A, p.a_dtype, p.M, p.K, p.lda, ATOM, BK);
  if (rc) return rc;
  if (!p.b_mn) rc = get_tensor_map(&tb, p.B, p.b_dtype, p.K, p.N, p.ldb, BK, BN / 2);   // each CTA loads half the N tile
  else rc = get_tensor_map(&tb, p.B, p.b_dtype, p.N, p.K, p.ldb, ATOM, BK);
  if (rc) return rc;
  GemmArgs g;
  g.M = p.M; g.N = p.N; g.K = p.K;
  g.kb_total = (p.K + BK - 1) / BK;
  int splits = p.splits < 1 ? 1 : p.splits;
  if (splits > g.kb_total) splits = g.kb_total;
  g.kb_per_split = (g.kb_total + splits - 1) / splits;
  splits = (g.kb_total + g.kb_per_split - 1) / g.kb_per_split;
  MAGB_REQUIRE(splits == 1 || p.epi == EPI_ATOMIC_F32, "gemm: split-K requires the atomic epilogue");
  g.n_split = 0x7fffffff; g.kb_start_hi = 0;
  g.epi = p.epi; g.bias = p.bias; g.out0 = p.out0; g.out1 = p.out1; g.aux = p.aux; g.ldo = p.ldo; g.alpha = p.alpha;
  const uint32_t fmt = p.a_dtype == DT_BF16 ? 1u : 0u;
  g.idesc = make_idesc(fmt, fmt, p.a_mn, p.b_mn, 256, BN);     // the pair's UMMA: M = 256
  g.trace = persistent ? g_trace : nullptr;
  g.dbg = (persistent || g_dbg == 3) ? g_dbg : 0;   // 4: staggered epilogue column groups (persistent kernel)
  g.drop_thresh = 0; g.drop_scale = 1.f; g.drop_site = p.drop_site; g.drop_seed = p.drop_seed.value;
  g.drop_seed_ptr = reinterpret_cast<const unsigned long long*>(p.drop_seed.ptr);
  if (p.drop_p > 0.f) {
    const double t = (double)p.drop_p * 4294967296.0;
    g.drop_thresh = t >= 4294967295.0 ? 4294967295u : (uint32_t)t;
    g.drop_scale = 1.0f / (1.0f - p.drop_p);
  }
  const int m_pairs = (p.M + 2 * BM - 1) / (2 * BM);
  g.tiles_m = m_pairs; g.tiles_n = (p.N + BN - 1) / BN; g.splits = splits;
  g.stream_k = 0;
  if (stream_k) {
    MAGB_REQUIRE(persistent && p.epi == EPI_ATOMIC_F32, "gemm: stream-K needs the persistent kernel and the atomic epilogue");
    g.stream_k = 1; g.splits = 1; g.kb_per_split = g.kb_total;
    MAGB_REQUIRE((long long)g.tiles_m * g.tiles_n * g.kb_total * 128 < (1ll << 31), "gemm: stream-K unit count overflows");
  }
  CUtensorMap to;
  rc = output_map(p, ta, &to, &g);
  if (rc) return rc;
  if (persistent) {
    if (BN == 256) return launch2p_majors<256, 4, 8>(p.a_mn, p.b_mn, ta, tb, to, g, st);
    return launch2p_majors<128, 6, 8>(p.a_mn, p.b_mn, ta, tb, to, g, st);
  }
  dim3 grid(2 * m_pairs, (p.N + BN - 1) / BN, splits);      // x = M tiles (pair-adjacent), y = N tiles
  if (BN == 256) return launch2_majors<256, 5>(p.a_mn, p.b_mn, ta, tb, to, g, grid, st);
  return launch2_majors<128, 4>(p.a_mn, p.b_mn, ta, tb, to, g, grid, st);
}

// Linear + dropout + residual + LayerNorm (N = 768) on clusters of 2 x 3 CTAs (gemm_ln_sm100.cuh)
int gemm_ln_launch(const GemmLnProblem& p, cudaStream_t st) {
  constexpr int STAGES = 4;
  using SM = GemmLnSmem<STAGES>;
  MAGB_REQUIRE(p.M > 0 && p.K > 0 && p.K % 8 == 0 && p.lda % 8 == 0, "gemm_ln: bad shape M=%d K=%d lda=%d", p.M, p.K, p.lda);
  MAGB_REQUIRE(p.A && p.W && p.bias && p.res && p.gamma && p.beta && p.y32, "gemm_ln: null operand");
  CUtensorMap ta, tb;
  int rc = get_tensor_map(&ta, p.A, DT_F16, (uint64_t)p.K, (uint64_t)p.M, (uint64_t)p.lda, 64, 128);
  if (rc) return rc;
  rc = get_tensor_map(&tb, p.W, DT_F16, (uint64_t)p.K, (uint64_t)LN_N, (uint64_t)p.K, 64, LN_BN / 2);
  if (rc) return rc;
  CUtensorMap tr;   // residual [M][768] fp32, boxes of 32 rows x 32 columns
  rc = get_tensor_map(&tr, p.res, DT_F32, (uint64_t)LN_N, (uint64_t)p.M, (uint64_t)LN_N, 32, 32);
  if (rc) return rc;
  GemmLnArgs g;
  g.M = p.M; g.K = p.K; g.kb_total = (p.K + 63) / 64;
  g.bias = p.bias; g.res = p.res; g.gamma = p.gamma; g.beta = p.beta; g.eps = p.eps;
  g.y32 = p.y32; g.y16 = p.y16; g.rstd = p.rstd;
  g.idesc = make_idesc(0u, 0u, false, false, 256, LN_BN);
  g.drop_thresh = 0; g.drop_scale = 1.f; g.drop_site = p.drop_site; g.drop_seed = p.drop_seed.value;
  g.drop_seed_ptr = reinterpret_cast<const unsigned long long*>(p.drop_seed.ptr);
  if (p.drop_p > 0.f) {
    const double t = (double)p.drop_p * 4294967296.0;
    g.drop_thresh = t >= 4294967295.0 ? 4294967295u : (uint32_t)t;
    g.drop_scale = 1.0f / (1.0f - p.drop_p);
  }
  auto kern = gemm2_ln_tcgen05_kernel<STAGES>;
  static bool configured[16] = {false};
  int dev = 0;
  MAGB_CHECK_CUDA(cudaGetDevice(&dev));
  MAGB_REQUIRE(dev >= 0 && dev < 16, "gemm_ln: device index %d out of range", dev);
  if (!configured[dev]) {
    MAGB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
    configured[dev] = true;
  }
  const int m_pairs = (p.M + 255) / 256;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * m_pairs, LN_NT, 1); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = SM::TOTAL; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = LN_NT; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1 + pdl_attr(attr + 1);
  MAGB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, tr, g));
  MAGB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int gemm_launch(const GemmProblem& p, cudaStream_t st) {
  static bool env_read = false;
  if (!env_read) {            // tuning override for whole-model measurements: MAGB200_GEMM_CFG=<cfg>
    env_read = true;
    const char* e = getenv("MAGB200_GEMM_CFG");
    if (e && *e && g_cfg == 0 && g_dbg == 0) { const int v = atoi(e); g_cfg = v % 100; g_dbg = v / 100; }
    const char* pol = getenv("MAGB200_GEMM_POLICY");
    if (pol && *pol) g_policy = atoi(pol);
    const char* sk = getenv("MAGB200_STREAMK");
    if (sk && *sk) g_streamk = atoi(sk);
    const char* skg = getenv("MAGB200_STREAMK_GRID");
    if (skg && *skg) g_streamk_grid = atoi(skg);
    const char* skm = getenv("MAGB200_STREAMK_MIN");
    if (skm && *skm) g_streamk_min = atoi(skm);
    const char* tr = getenv("MAGB200_TMA_REDUCE");
    if (tr && *tr) g_tma_reduce = atoi(tr);
  }
  MAGB_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0, "gemm: empty problem %dx%dx%d", p.M, p.N, p.K);
  const bool a32 = p.a_dtype == DT_F32, b32 = p.b_dtype == DT_F32;
  // measured on B200: tcgen05.mma kind::f16 traps (illegal instruction) when A and B formats differ,
  // and kind::tf32 with MN-major operands returns wrong results with this descriptor scheme.
  MAGB_REQUIRE(p.a_dtype == p.b_dtype, "gemm: A and B must have the same element type (f16, bf16 or f32/tf32)");
  MAGB_REQUIRE(!(a32 && (p.a_mn || p.b_mn)), "gemm: tf32 operands must be K-major");
  const int EB = a32 ? 4 : 2;
  const int BK = 128 / EB, ATOM = 128 / EB, BM = 128;
  const int BN = (g_cfg == 3 || g_cfg == 4) ? 256 : 128;
  MAGB_REQUIRE(p.N % 32 == 0, "gemm: N=%d must be a multiple of 32", p.N);
  MAGB_REQUIRE(p.ldo % 8 == 0, "gemm: ldo=%d must be a multiple of 8", p.ldo);
  // the staged red.add epilogue hands whole 16-byte row pieces to lanes: a tile's column count must be 32, 64, 128 or 256
  // (with other widths lanes would repeat pieces, harmless for stores but a double count for reductions)
  MAGB_REQUIRE(p.epi != EPI_ATOMIC_F32 || p.N % 128 == 0 || p.N == 32 || p.N == 64,
               "gemm: the atomic epilogue needs N = 32, 64 or a multiple of 128 (got %d)", p.N);
  // 2-CTA (cta_group::2) pair tiles.  Measured on B200 (profiles/r01_gemm_modes.txt, r01_gemm_tile_sweep.txt):
  // the 1-CTA 128x128 SS-mode main loop is shared-memory-bandwidth bound (UMMA operand reads 128 B/clk + TMA fill
  // 128 B/clk against 128 B/clk of SMEM), the pair tile halves the per-SM B traffic.  At these problem sizes
  // (M = 2400) 256x128 pair tiles win 7-9 % on the long-K GEMMs (K >= 2048: FFN-out, its dgrad, the FFN/QKV
  // wgrads); for K = 768 the epilogue and tile quantisation dominate and the 1-CTA kernel with two co-resident
  // CTAs per SM is as fast.  cfg 5 forces 1-CTA, cfg 7 forces 256x128 pairs, cfg 8 256x256 pairs where N allows.
  const bool pair_ok = EB == 2 && p.n_split <= 0 && p.N % 128 == 0;
  // cfg 9: persistent pair kernel everywhere (256x256 tiles where N is a multiple of 256 and >= 1024, else 256x128);
  // cfg 10: persistent with 256x128 tiles only.
  // default policy (cfg 0): persistent pair kernel for the roles in g_policy (bit 0 forward, bit 1 dgrad, bit 2 wgrad;
  // MAGB200_GEMM_POLICY overrides) -- measured per role in profiles/r01_gemm_policy.txt
  // stream-K (cfg 0, or forced with cfg 11): every cluster gets the same number of K blocks; 256-column tiles whenever N
  // allows (fewer operand bytes per flop; tile quantisation no longer matters)
  if (pair_ok && p.epi == EPI_ATOMIC_F32 && ((g_cfg == 0 && g_streamk) || g_cfg == 11)) {
    const int bn2 = (p.N % 256 == 0) ? 256 : 128;
    return gemm_launch_2cta(p, bn2, st, true, true);
  }
  if (pair_ok && g_cfg == 0 && ((g_policy >> p.role) & 1)) {
    const int bn2 = (p.N % 256 == 0 && p.N >= 1024) ? 256 : 128;
    return gemm_launch_2cta(p, bn2, st, true);
  }
  if (pair_ok && (g_cfg == 9 || g_cfg == 10)) {
    const int bn2 = (g_cfg == 9 && p.N % 256 == 0 && p.N >= 1024) ? 256 : 128;
    return gemm_launch_2cta(p, bn2, st, true);
  }
  if (pair_ok && (g_cfg == 7 || g_cfg == 8 || (g_cfg == 0 && p.K >= 2048))) {
    const int bn2 = (g_cfg == 8 && p.N % 256 == 0 && p.N >= 1024) ? 256 : 128;
    return gemm_launch_2cta(p, bn2, st);
  }

  // cluster of 2 CTAs along N sharing (multicasting) the A tile: needs an even number of N tiles and identical K
  // ranges in both CTAs (no n_split).  Measured on B200 (profiles/r01_gemm_multicast.txt): correct, but no faster --
  // the main loop is bound by bytes ARRIVING per SM (~56-84 B/clk), which multicast does not reduce; it stays
  // opt-in (cfg 6) until the 2-CTA (cta_group::2) tile halves the per-SM B traffic.
  const int n_tiles = (p.N + BN - 1) / BN;
  const bool cluster2 = (EB == 2) && (BN == 128) && (n_tiles % 2 == 0) && (p.n_split <= 0) && (g_cfg == 6);
  CUtensorMap ta, tb;
  int rc;
  if (!p.a_mn) rc = get_tensor_map(&ta, p.A, p.a_dtype, p.K, p.M, p.lda, BK, cluster2 ? BM / 2 : BM);
  else rc = get_tensor_map(&ta, p.A, p.a_dtype, p.M, p.K, p.lda, ATOM, BK);
  if (rc) return rc;
  if (!p.b_mn) rc = get_tensor_map(&tb, p.B, p.b_dtype, p.K, p.N, p.ldb, BK, BN);
  else rc = get_tensor_map(&tb, p.B, p.b_dtype, p.N, p.K, p.ldb, ATOM, BK);
  if (rc) return rc;

  GemmArgs g;
  g.M = p.M; g.N = p.N; g.K = p.K;
  g.kb_total = (p.K + BK - 1) / BK;
  int splits = p.splits < 1 ? 1 : p.splits;
  if (splits > g.kb_total) splits = g.kb_total;
  g.kb_per_split = (g.kb_total + splits - 1) / splits;
  splits = (g.kb_total + g.kb_per_split - 1) / g.kb_per_split;  // no empty split
  MAGB_REQUIRE(splits == 1 || p.epi == EPI_ATOMIC_F32, "gemm: split-K requires the atomic epilogue");
  g.n_split = p.n_split > 0 ? p.n_split : 0x7fffffff;
  MAGB_REQUIRE(p.n_split <= 0 || p.n_split % BN == 0, "gemm: n_split must be a multiple of the N tile");
  g.kb_start_hi = p.k_start_hi / BK;
  MAGB_REQUIRE(p.k_start_hi % BK == 0, "gemm: k_start_hi must be a multiple of the K block");
  MAGB_REQUIRE(p.n_split <= 0 || (splits == 1 && g.kb_start_hi < g.kb_total && p.n_split % BN == 0),
               "gemm: invalid n_split configuration");
  g.epi = p.epi; g.bias = p.bias; g.out0 = p.out0; g.out1 = p.out1; g.aux = p.aux; g.ldo = p.ldo; g.alpha = p.alpha;
  const uint32_t af = p.a_dtype == DT_F32 ? 2u : (p.a_dtype == DT_BF16 ? 1u : 0u);
  const uint32_t bf = p.b_dtype == DT_F32 ? 2u : (p.b_dtype == DT_BF16 ? 1u : 0u);
  g.idesc = make_idesc(af, bf, p.a_mn, p.b_mn, BM, BN);
  g.trace = g_trace;
  g.dbg = g_dbg;
  g.drop_thresh = 0; g.drop_scale = 1.f; g.drop_site = p.drop_site; g.drop_seed = p.drop_seed.value;
  g.drop_seed_ptr = reinterpret_cast<const unsigned long long*>(p.drop_seed.ptr);
  if (p.drop_p > 0.f) {
    MAGB_REQUIRE(p.epi == EPI_GELU_F16 || p.epi == EPI_DGELU_F16, "gemm: fused dropout is only available in the GELU epilogues");
    MAGB_REQUIRE(p.N % 4 == 0, "gemm: fused dropout needs N % 4 == 0");
    const double t = (double)p.drop_p * 4294967296.0;
    g.drop_thresh = t >= 4294967295.0 ? 4294967295u : (uint32_t)t;
    g.drop_scale = 1.0f / (1.0f - p.drop_p);
  }

  CUtensorMap to;
  rc = output_map(p, ta, &to, &g);
  if (rc) return rc;
  dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, splits);
  if (EB == 2) {
    if (cluster2) return launch_majors<2, 128, 3, 2>(p.a_mn, p.b_mn, ta, tb, to, g, grid, st);
    switch (g_cfg) {
      case 1: return launch_majors<2, 128, 4, 1>(p.a_mn, p.b_mn, ta, tb, to, g, grid, st);
      case 3: return launch_majors<2, 256, 3, 1>(p.a_mn, p.b_mn, ta, tb, to, g, grid, st);
      case 4: return launch_majors<2, 256, 4, 1>(p.a_mn, p.b_mn, ta, tb, to, g, grid, st);
      default: return launch_majors<2, 128, 3, 1>(p.a_mn, p.b_mn, ta, tb, to, g, grid, st);
    }
  }
  return launch_majors<4, 128, 3, 1>(p.a_mn, p.b_mn, ta, tb, to, g, grid, st);
}

}  // namespace magb200

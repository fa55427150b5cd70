// MAG-BERT forward / backward: the kernel schedule of
//   MAG_BertForSequenceClassification.forward (bert.py:251-324) -> MAG_BertModel.forward (bert.py:106-237)
//   -> BertEmbeddings, MAG (modeling.py:25-51), 12 x BertLayer, BertPooler, classifier
// and of the autograd pass the reference triggers at multimodal_driver.py:378.
//
// Everything is enqueued on one stream; all memory (parameters, fp16 operand copies, workspace,
// gradients) belongs to the caller.  Parameter gradients are accumulated (+=), unscaled, into a flat
// buffer that mirrors the parameter buffer.  Activation gradients inside backward carry
// dims.grad_scale (loss scaling: the backward GEMM operands are fp16); every parameter-gradient
// write multiplies by 1/grad_scale.
#include <stdlib.h>
#include "common.cuh"
#include "gemm_sm100.cuh"
#include "host.h"
#include "../../include/magb200.h"

namespace magb200 {

namespace {

struct Carve {
  uint8_t* base;
  size_t off = 0;
  explicit Carve(void* b) : base(reinterpret_cast<uint8_t*>(b)) {}
  template <typename T>
  T* take(size_t n) {
    off = (off + 1023) & ~size_t(1023);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

struct LayerWs {
  __half* qkv16; __half* ctx16; float* a32; __half* a16; __half* gp16; __half* g16; float* rstd1; float* rstd2;
  float* lse;
};

struct BertWs {
  __half* amag16; float* emb32; float* G32;
  float* x32[MAGB200_MAX_LAYERS + 1];
  __half* x16[MAGB200_MAX_LAYERS + 1];
  LayerWs layer[MAGB200_MAX_LAYERS];
  float* pre32; float* pre32b; float* pre32c; float* poolpre32; float* pooled32;
  // backward scratch
  // The fp16 gradient tensors the weight-gradient GEMMs consume exist twice (layer parity): those GEMMs run on a
  // side stream and may still be reading layer l+1's copy while the main chain produces layer l's.
  float* dy32; float* dr32; __half* dpre_f16[2]; __half* dpre_a16[2]; __half* du16[2]; __half* dctx16;
  __half* dqkv16[2]; __half* dG16;
  float* dBmag32; __half* dpool16;
  float* attD;     // [B][heads][S] rowsum(dO * O) of the layer in flight (attention backward for S > 64)
  void* magscratch;   // K1 (mag_fused.cu): h_m rows + grid-barrier words
  size_t bytes;
};

BertWs carve_ws(const magb200_bert_dims& d, void* base) {
  BertWs w;
  Carve c(base);
  const size_t T = (size_t)d.batch * d.seq, H = d.hidden, I = d.inter;
  w.amag16 = c.take<__half>(T * d.ka);
  w.emb32 = c.take<float>(T * H);
  w.G32 = c.take<float>(T * 4 * H);
  for (int l = 0; l <= d.layers; ++l) {
    w.x32[l] = c.take<float>(T * H);
    w.x16[l] = c.take<__half>(T * H);
  }
  for (int l = 0; l < d.layers; ++l) {
    LayerWs& L = w.layer[l];
    L.qkv16 = c.take<__half>(T * 3 * H);
    L.ctx16 = c.take<__half>(T * H);
    L.a32 = c.take<float>(T * H);
    L.a16 = c.take<__half>(T * H);
    L.gp16 = c.take<__half>(T * I);
    L.g16 = c.take<__half>(T * I);
    L.rstd1 = c.take<float>(T);
    L.rstd2 = c.take<float>(T);
    L.lse = c.take<float>((size_t)d.batch * d.heads * d.seq);
  }
  w.pre32 = c.take<float>(T * H);
  w.pre32b = c.take<float>(T * H);
  w.pre32c = c.take<float>(T * H);
  w.poolpre32 = c.take<float>((size_t)d.batch * H);
  w.pooled32 = c.take<float>((size_t)d.batch * H);
  w.dy32 = c.take<float>(T * H);
  w.dr32 = c.take<float>(T * H);
  for (int k = 0; k < 2; ++k) {
    w.dpre_f16[k] = c.take<__half>(T * H);
    w.dpre_a16[k] = c.take<__half>(T * H);
    w.du16[k] = c.take<__half>(T * I);
    w.dqkv16[k] = c.take<__half>(T * 3 * H);
  }
  w.dctx16 = c.take<__half>(T * H);
  w.dG16 = c.take<__half>(T * 4 * H);
  w.dBmag32 = c.take<float>((size_t)4 * H * d.ka);
  w.dpool16 = c.take<__half>((size_t)d.batch * H);
  w.attD = c.take<float>((size_t)d.batch * d.heads * d.seq);
  w.magscratch = c.take<uint8_t>(mag_fused_scratch_bytes((int)T));
  w.bytes = (c.off + 1023) & ~size_t(1023);
  return w;
}

int check_dims(const magb200_bert_dims& d) {
  MAGB_REQUIRE(d.hidden == 768 && d.heads == 12, "bert: kernels are specialised for hidden=768, heads=12 (got %d, %d)",
               d.hidden, d.heads);
  MAGB_REQUIRE(d.inter % 128 == 0 && d.inter > 0, "bert: intermediate size %d must be a multiple of 128", d.inter);
  MAGB_REQUIRE(d.layers >= 1 && d.layers <= MAGB200_MAX_LAYERS, "bert: layers=%d out of range", d.layers);
  MAGB_REQUIRE(d.batch >= 1 && d.seq >= 1 && d.seq <= 512 && d.seq <= d.max_pos, "bert: bad batch/seq %d/%d", d.batch, d.seq);
  MAGB_REQUIRE(d.ka == d.hidden + 128 && d.visual_dim + d.acoustic_dim <= 128 && d.visual_dim > 0 && d.acoustic_dim > 0,
               "bert: visual_dim + acoustic_dim = %d must be <= 128 (ka=%d)", d.visual_dim + d.acoustic_dim, d.ka);
  MAGB_REQUIRE(d.num_labels >= 1 && d.num_labels <= 32, "bert: num_labels=%d out of range", d.num_labels);
  return 0;
}

// dropout sites (counter-RNG stream ids)
enum : uint32_t { SITE_EMB = 1, SITE_MAG = 2, SITE_CLS = 3, SITE_LAYER0 = 16 };
inline uint32_t site_attn_prob(int l) { return SITE_LAYER0 + 4 * l; }
inline uint32_t site_attn_out(int l) { return SITE_LAYER0 + 4 * l + 1; }
inline uint32_t site_ffn_out(int l) { return SITE_LAYER0 + 4 * l + 2; }

// Side stream for the weight-gradient GEMMs and bias column sums of the encoder layers.  They depend only on the
// activation-gradient tensor just produced and on saved activations, nothing downstream reads their results before
// the optimizer (or the gradient all-reduce of the piece), so they fill the SMs the main dgrad chain leaves idle in
// its partial last waves.  Ordering uses events only (valid eagerly and under stream capture, where the side stream
// joins the capture through the first wait and is joined back before the piece returns).
constexpr int MAGB200_MAX_DEVICES = 16;
struct SideCtx {
  cudaStream_t stream = nullptr, opt = nullptr;
  cudaEvent_t fork = nullptr, done[2] = {nullptr, nullptr}, join = nullptr, lay = nullptr, optjoin = nullptr;
  bool enabled = true;
  bool init() {
    if (stream) return true;
    const char* e = getenv("MAGB200_SIDE_STREAM");
    if (e && e[0] == '0') { enabled = false; return true; }
    if (cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess) return false;
    if (cudaStreamCreateWithFlags(&opt, cudaStreamNonBlocking) != cudaSuccess) return false;
    cudaEvent_t* evs[6] = {&fork, &done[0], &done[1], &join, &lay, &optjoin};
    for (cudaEvent_t* ev : evs)
      if (cudaEventCreateWithFlags(ev, cudaEventDisableTiming) != cudaSuccess) return false;
    return true;
  }
};
SideCtx g_side[MAGB200_MAX_DEVICES];

#define RUN(expr)            \
  do {                       \
    int _rc = (expr);        \
    if (_rc) return _rc;     \
  } while (0)

// MAGB200_FUSED_LN=1: the attention-output and FFN-output Linears run with dropout + residual + LayerNorm in their
// epilogue (gemm_ln_sm100.cuh) instead of a GEMM followed by dropout_residual_ln_fwd
bool fused_ln() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MAGB200_FUSED_LN");
    v = (e && *e) ? atoi(e) : 0;
  }
  return v != 0;
}

int linear_ln_fwd(const __half* A, int lda, const __half* W, int M, int K, const float* bias, const float* res,
                  const float* gamma, const float* beta, float eps, float p, Seed sd, uint32_t site, float* y32, __half* y16,
                  float* rstd, cudaStream_t st) {
  GemmLnProblem q;
  q.A = A; q.W = W; q.M = M; q.K = K; q.lda = lda; q.bias = bias; q.res = res; q.gamma = gamma; q.beta = beta; q.eps = eps;
  q.y32 = y32; q.y16 = y16; q.rstd = rstd; q.drop_p = p; q.drop_seed = sd; q.drop_site = site;
  return gemm_ln_launch(q, st);
}

bool fused_mag() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MAGB200_FUSED_MAG");
    v = (e && *e) ? atoi(e) : 0;
  }
  return v != 0;
}

int streamk_forward() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MAGB200_STREAMK");
    const char* f = getenv("MAGB200_STREAMK_FWD");
    v = (e && e[0] == '0') ? 0 : (f && *f ? atoi(f) : 0);
  }
  return v;
}

bool accumulate_dgrad_atomic() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MAGB200_STREAMK");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

// C[M,N] = A[M,K] * W[N,K]^T : forward Linear (both operands K-major, fp16)
int linear_fwd(const __half* A, int lda, const __half* W, int M, int N, int K, const float* bias, int epi, void* out0,
               void* out1, int ldo, cudaStream_t st, int n_split = 0, int k_start_hi = 0) {
  GemmProblem p;
  p.A = A; p.B = W; p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldb = K;
  p.a_dtype = DT_F16; p.b_dtype = DT_F16; p.epi = epi; p.bias = bias; p.out0 = out0; p.out1 = out1; p.ldo = ldo;
  p.n_split = n_split; p.k_start_hi = k_start_hi;
  return gemm_launch(p, st);
}
// dX[M,Kin] = dY[M,Nout] * W[Nout,Kin] : W is read in place as an MN-major B operand
int linear_dgrad(const __half* dY, int ldy, const __half* W, int ldw, int M, int Kin, int Nout, int epi, void* out0,
                 const void* aux, int ldo, cudaStream_t st) {
  GemmProblem p;
  p.A = dY; p.B = W; p.M = M; p.N = Kin; p.K = Nout; p.lda = ldy; p.ldb = ldw; p.b_mn = true;
  p.a_dtype = DT_F16; p.b_dtype = DT_F16; p.epi = epi; p.out0 = out0; p.aux = aux; p.ldo = ldo;
  p.role = 1;
  // accumulate into the running residual gradient (out0 == aux): a reduction, i.e. the atomic epilogue, which lets
  // the GEMM run with stream-K work ranges (no auxiliary tile to stage either)
  if (epi == EPI_STORE_F32 && aux != nullptr && aux == out0 && accumulate_dgrad_atomic()) {
    p.epi = EPI_ATOMIC_F32;
    p.aux = nullptr;
  }
  return gemm_launch(p, st);
}
// dW[Nout,Kin] += dY[T,Nout]^T * X[T,Kin] : both operands token-major, read in place (MN-major)
int linear_wgrad(const __half* dY, int ldy, const __half* X, int ldx, int T, int Nout, int Kin, float* dW, int lddw,
                 int splits, float oscale, cudaStream_t st) {
  GemmProblem p;
  p.A = dY; p.B = X; p.M = Nout; p.N = Kin; p.K = T; p.lda = ldy; p.ldb = ldx; p.a_mn = true; p.b_mn = true;
  p.a_dtype = DT_F16; p.b_dtype = DT_F16; p.epi = EPI_ATOMIC_F32; p.out0 = dW; p.ldo = lddw; p.splits = splits; p.alpha = oscale;
  p.role = 2;
  return gemm_launch(p, st);
}

}  // namespace

size_t bert_workspace_bytes(const magb200_bert_dims& d) { return carve_ws(d, nullptr).bytes; }

// Refresh the fp16 operand copies after the fp32 master parameters changed.
int bert_refresh_operands(const magb200_bert_dims& d, const int64_t* off, const float* params, __half* params16,
                          cudaStream_t st) {
  RUN(check_dims(d));
  RUN(cast_f32_f16(params, params16, off[MAGB200_OFF_TOTAL], st));
  RUN(mag_pack_weights(params + off[MAGB200_OFF_MAG_WHV], params + off[MAGB200_OFF_MAG_WHA],
                       params + off[MAGB200_OFF_MAG_WV], params + off[MAGB200_OFF_MAG_WA], d.visual_dim,
                       d.acoustic_dim, d.ka, params16 + off[MAGB200_OFF_MAG_PACKED16], st));
  return 0;
}

int bert_forward(const magb200_bert_dims& d, const int64_t* off, const float* P, const __half* P16,
                 const int64_t* ids, const int64_t* tt, const int64_t* mask, const float* visual,
                 const float* acoustic, void* workspace, size_t ws_bytes, float* logits, cudaStream_t st) {
  RUN(check_dims(d));
  BertWs w = carve_ws(d, workspace);
  MAGB_REQUIRE(ws_bytes >= w.bytes, "bert_forward: workspace too small (%zu < %zu)", ws_bytes, w.bytes);
  const int B = d.batch, S = d.seq, T = B * S, H = d.hidden, I = d.inter;
  const Seed sd{d.seed, d.seed_dev};

  // BertEmbeddings (+ packing of the MAG GEMM operand)                         bert.py:211-216
  EmbedArgs e;
  e.ids = ids; e.tt = tt; e.visual = visual; e.acoustic = acoustic;
  e.word = P + off[MAGB200_OFF_WORD]; e.pos = P + off[MAGB200_OFF_POS]; e.type = P + off[MAGB200_OFF_TYPE];
  e.gamma = P + off[MAGB200_OFF_EMB_LN_W]; e.beta = P + off[MAGB200_OFF_EMB_LN_B];
  e.eps = d.ln_eps; e.p = d.p_hidden; e.seed = sd; e.site = SITE_EMB;
  e.V = d.visual_dim; e.A = d.acoustic_dim; e.KA = d.ka; e.T = T; e.S = S; e.vocab = d.vocab;
  e.emb32 = w.emb32; e.amag = w.amag16;
  if (fused_mag()) {
    // K1: embeddings + MAG in one warp-specialised cooperative kernel (mag_fused.cu); the gate pre-activations are
    // written for the backward pass but not read back
    RUN(mag_fused_fwd(e, P16 + off[MAGB200_OFF_MAG_PACKED16], P + off[MAGB200_OFF_MAG_BIAS4], P + off[MAGB200_OFF_MAG_LN_W],
                      P + off[MAGB200_OFF_MAG_LN_B], d.beta_shift, d.mag_ln_eps, d.p_mag, SITE_MAG, w.G32, w.magscratch,
                      w.x32[0], w.x16[0], st));
  } else {
    RUN(embed_ln_fwd(e, st));
    // MAG (modeling.py:25-51): one block GEMM for the four Linears, then the fused tail   bert.py:219
    RUN(linear_fwd(w.amag16, d.ka, P16 + off[MAGB200_OFF_MAG_PACKED16], T, 4 * H, d.ka, P + off[MAGB200_OFF_MAG_BIAS4],
                   EPI_STORE_F32, w.G32, nullptr, 4 * H, st, 2 * H, H));
    RUN(mag_tail_fwd(w.G32, w.emb32, P + off[MAGB200_OFF_MAG_LN_W], P + off[MAGB200_OFF_MAG_LN_B], d.beta_shift,
                     d.mag_ln_eps, d.p_mag, sd, SITE_MAG, T, w.x32[0], w.x16[0], st));
  }
  // BertEncoder                                                                  bert.py:221-229
  // The FFN-output Linear (N = 768, K = 3072) has 30 pair tiles of 48 K blocks for 74 clusters: it can run as a stream-K
  // reduction into a zeroed buffer (red.add epilogue, bias carried by the partial sum that holds K block 0), which the
  // LayerNorm kernel that consumes it hands back zeroed.  MAGB200_STREAMK_FWD: 0 = off, 1 = FFN-out, 2 = attention
  // output too (K = 768: 5 K blocks per cluster against a full-tile reduction epilogue -- measured slower).
  const bool fln = fused_ln();
  const int skf = streamk_forward();
  const int epi_ffn = skf >= 1 ? EPI_ATOMIC_F32 : EPI_STORE_F32, epi_ao = skf >= 2 ? EPI_ATOMIC_F32 : EPI_STORE_F32;
  float* pre_ffn = skf >= 1 ? w.pre32b : w.pre32;       // reduction targets are dedicated, always-zero buffers
  float* pre_ao = skf >= 2 ? w.pre32c : w.pre32;
  if (skf >= 1) MAGB_CHECK_CUDA(cudaMemsetAsync(w.pre32b, 0, (size_t)T * H * sizeof(float), st));
  if (skf >= 2) MAGB_CHECK_CUDA(cudaMemsetAsync(w.pre32c, 0, (size_t)T * H * sizeof(float), st));
  for (int l = 0; l < d.layers; ++l) {
    const int64_t* o = off + MAGB200_OFF_LAYER0 + (int64_t)l * MAGB200_OFF_PER_LAYER;
    LayerWs& L = w.layer[l];
    RUN(linear_fwd(w.x16[l], H, P16 + o[MAGB200_LOFF_QKV_W], T, 3 * H, H, P + o[MAGB200_LOFF_QKV_B], EPI_STORE_F16,
                   L.qkv16, nullptr, 3 * H, st));
    RUN(sdpa_fwd(L.qkv16, mask, B, S, d.heads, H / d.heads, d.p_attn, sd, site_attn_prob(l), L.ctx16, L.lse, st));
    if (fln) {
      RUN(linear_ln_fwd(L.ctx16, H, P16 + o[MAGB200_LOFF_AO_W], T, H, P + o[MAGB200_LOFF_AO_B], w.x32[l],
                        P + o[MAGB200_LOFF_LN1_W], P + o[MAGB200_LOFF_LN1_B], d.ln_eps, d.p_hidden, sd, site_attn_out(l), L.a32,
                        L.a16, L.rstd1, st));
    } else {
    RUN(linear_fwd(L.ctx16, H, P16 + o[MAGB200_LOFF_AO_W], T, H, H, P + o[MAGB200_LOFF_AO_B], epi_ao, pre_ao,
                   nullptr, H, st));
    RUN(dropout_residual_ln_fwd(pre_ao, w.x32[l], P + o[MAGB200_LOFF_LN1_W], P + o[MAGB200_LOFF_LN1_B], d.ln_eps,
                                d.p_hidden, sd, site_attn_out(l), T, L.a32, L.a16, L.rstd1, st, skf >= 2 ? pre_ao : nullptr));
    }
    RUN(linear_fwd(L.a16, H, P16 + o[MAGB200_LOFF_F1_W], T, I, H, P + o[MAGB200_LOFF_F1_B], EPI_GELU_F16, L.g16, L.gp16,
                   I, st));
    if (fln) {
      RUN(linear_ln_fwd(L.g16, I, P16 + o[MAGB200_LOFF_F2_W], T, I, P + o[MAGB200_LOFF_F2_B], L.a32, P + o[MAGB200_LOFF_LN2_W],
                        P + o[MAGB200_LOFF_LN2_B], d.ln_eps, d.p_hidden, sd, site_ffn_out(l), w.x32[l + 1], w.x16[l + 1],
                        L.rstd2, st));
    } else {
    RUN(linear_fwd(L.g16, I, P16 + o[MAGB200_LOFF_F2_W], T, H, I, P + o[MAGB200_LOFF_F2_B], epi_ffn, pre_ffn,
                   nullptr, H, st));
    RUN(dropout_residual_ln_fwd(pre_ffn, L.a32, P + o[MAGB200_LOFF_LN2_W], P + o[MAGB200_LOFF_LN2_B], d.ln_eps,
                                d.p_hidden, sd, site_ffn_out(l), T, w.x32[l + 1], w.x16[l + 1], L.rstd2, st,
                                skf >= 1 ? pre_ffn : nullptr));
    }
  }
  // BertPooler on the CLS rows (row pitch S*H) + classifier                       bert.py:231, 304-307
  RUN(linear_fwd(w.x16[d.layers], S * H, P16 + off[MAGB200_OFF_POOL_W], B, H, H, P + off[MAGB200_OFF_POOL_B],
                 EPI_STORE_F32, w.poolpre32, nullptr, H, st));
  RUN(cls_fwd(w.poolpre32, P + off[MAGB200_OFF_CLS_W], P + off[MAGB200_OFF_CLS_B], B, d.num_labels, d.p_cls, sd,
              SITE_CLS, w.pooled32, logits, st));
  return 0;
}

int bert_backward(const magb200_bert_dims& d, const int64_t* off, const float* P, const __half* P16,
                  const int64_t* ids, const int64_t* tt, const int64_t* mask, void* workspace, size_t ws_bytes,
                  const float* dlogits_scaled, float* G, cudaStream_t st) {
  return bert_backward_part(d, off, P, P16, ids, tt, mask, workspace, ws_bytes, dlogits_scaled, G, d.layers - 1, 0, 3, st,
                            nullptr);
}

// Backward in pieces so the caller can all-reduce finished gradient ranges while later pieces run:
// layers layer_hi .. layer_lo (descending); flags bit 0 = run the head (classifier + pooler) first,
// bit 1 = run the tail (MAG + embeddings) afterwards.  The running activation gradient lives in the workspace.
int bert_backward_part(const magb200_bert_dims& d, const int64_t* off, const float* P, const __half* P16,
                       const int64_t* ids, const int64_t* tt, const int64_t* mask, void* workspace, size_t ws_bytes,
                       const float* dlogits_scaled, float* G, int layer_hi, int layer_lo, int flags, cudaStream_t st,
                       const OptHook* opt) {
  RUN(check_dims(d));
  MAGB_REQUIRE(layer_hi < d.layers && layer_lo >= 0, "bert_backward_part: layer range [%d, %d] out of bounds", layer_lo, layer_hi);
  BertWs w = carve_ws(d, workspace);
  MAGB_REQUIRE(ws_bytes >= w.bytes, "bert_backward: workspace too small (%zu < %zu)", ws_bytes, w.bytes);
  const int B = d.batch, S = d.seq, T = B * S, H = d.hidden, I = d.inter;
  const float us = 1.0f / d.grad_scale;   // parameter gradients are stored unscaled
  const Seed sd{d.seed, d.seed_dev};
  int dev = 0;
  MAGB_CHECK_CUDA(cudaGetDevice(&dev));
  MAGB_REQUIRE(dev >= 0 && dev < MAGB200_MAX_DEVICES, "bert_backward: device index %d out of range", dev);
  SideCtx& sc = g_side[dev];
  MAGB_REQUIRE(sc.init(), "bert_backward: cannot create the side stream / events");
  const bool side = sc.enabled;
  cudaStream_t sst = side ? sc.stream : st;
  // main -> side: everything enqueued on `st` so far is visible to what the side stream runs next
  auto fork = [&]() -> int {
    if (!side) return 0;
    MAGB_CHECK_CUDA(cudaEventRecord(sc.fork, st));
    MAGB_CHECK_CUDA(cudaStreamWaitEvent(sst, sc.fork, 0));
    return 0;
  };
  bool side_used = false, opt_used = false;
  int layers_done = 0;
  // Optimizer hook (single-process training): AdamW over a layer's slice of the flat buffers is enqueued on a third
  // stream as soon as that layer's weight gradients (side stream) are complete and the main chain has left the
  // layer (its dgrad GEMMs read the fp16 weights AdamW rewrites).  AdamW is HBM-bound, the GEMM chain is not, so
  // the update of layers L-1 .. 1 disappears behind the backward pass of the layers below.
  cudaStream_t ost = (opt && side) ? sc.opt : st;
  auto adamw_range = [&](int64_t s0, int64_t s1, cudaStream_t stream) -> int {
    if (s1 <= s0) return 0;
    MAGB_REQUIRE(s0 % 256 == 0 && s1 % 256 == 0, "backward_adamw: slice [%lld, %lld) is not 256-aligned", (long long)s0,
                 (long long)s1);
    return adamw_step(opt->params + s0, G + s0, opt->exp_avg + s0, opt->exp_avg_sq + s0,
                      opt->params16 ? opt->params16 + s0 : nullptr, opt->decay_chunk + s0 / 256, s1 - s0, opt->lr,
                      opt->beta1, opt->beta2, opt->eps, opt->weight_decay, opt->step, opt->grad_unscale, opt->zero_grad,
                      opt->hyper_dev, stream);
  };
  auto layer_start = [&](int l) -> int64_t {
    return l < d.layers ? off[MAGB200_OFF_LAYER0 + (int64_t)l * MAGB200_OFF_PER_LAYER + MAGB200_LOFF_QKV_W]
                        : off[MAGB200_OFF_POOL_W];
  };
  if (opt) {
    for (int l = 0; l < d.layers; ++l)
      MAGB_REQUIRE(layer_start(l) < layer_start(l + 1) && layer_start(0) > off[MAGB200_OFF_EMB_LN_B],
                   "backward_adamw: flat order must be [embeddings | layers ascending | pooler, MAG, classifier]");
    MAGB_REQUIRE(off[MAGB200_OFF_TOTAL] > off[MAGB200_OFF_POOL_W], "backward_adamw: bad OFF_TOTAL");
  }

  if (flags & 1) {
  // classifier + pooler                                                         bert.py:304-307, 231
  RUN(cls_bwd(dlogits_scaled, w.pooled32, P + off[MAGB200_OFF_CLS_W], B, d.num_labels, d.p_cls, sd, SITE_CLS,
              us, w.dpool16, G + off[MAGB200_OFF_CLS_W], G + off[MAGB200_OFF_CLS_B], st));
  MAGB_CHECK_CUDA(cudaMemsetAsync(w.dy32, 0, (size_t)T * H * sizeof(float), st));
  RUN(linear_dgrad(w.dpool16, H, P16 + off[MAGB200_OFF_POOL_W], H, B, H, H, EPI_STORE_F32, w.dy32, nullptr, S * H, st));
  RUN(linear_wgrad(w.dpool16, H, w.x16[d.layers], S * H, B, H, H, G + off[MAGB200_OFF_POOL_W], H, 1, us, st));
  RUN(colsum_f16(w.dpool16, B, H, us, G + off[MAGB200_OFF_POOL_B], st));
  }

  for (int l = layer_hi; l >= layer_lo; --l) {
    const int64_t* o = off + MAGB200_OFF_LAYER0 + (int64_t)l * MAGB200_OFF_PER_LAYER;
    LayerWs& L = w.layer[l];
    const int par = l & 1;
    __half* dpre_f = w.dpre_f16[par];
    __half* dpre_a = w.dpre_a16[par];
    __half* du = w.du16[par];
    __half* dqkv = w.dqkv16[par];
    // this layer overwrites the gradient copies the side stream read two layers ago
    if (side && layers_done >= 2) MAGB_CHECK_CUDA(cudaStreamWaitEvent(st, sc.done[par], 0));
    // BertOutput: LayerNorm(dropout(dense(g)) + a)
    RUN(ln_bwd(w.dy32, w.x32[l + 1], P + o[MAGB200_LOFF_LN2_W], P + o[MAGB200_LOFF_LN2_B], L.rstd2, d.p_hidden, sd,
               site_ffn_out(l), T, us, w.dr32, dpre_f, G + o[MAGB200_LOFF_LN2_W], G + o[MAGB200_LOFF_LN2_B],
               G + o[MAGB200_LOFF_F2_B], st));
    RUN(fork());
    RUN(linear_wgrad(dpre_f, H, L.g16, I, T, H, I, G + o[MAGB200_LOFF_F2_W], I, 1, us, sst));
    RUN(linear_dgrad(dpre_f, H, P16 + o[MAGB200_LOFF_F2_W], I, T, I, H, EPI_DGELU_F16, du, L.gp16, I, st));
    // BertIntermediate: gelu(dense(a))
    RUN(fork());
    RUN(colsum_f16(du, T, I, us, G + o[MAGB200_LOFF_F1_B], sst));
    RUN(linear_wgrad(du, I, L.a16, H, T, I, H, G + o[MAGB200_LOFF_F1_W], H, 1, us, sst));
    RUN(linear_dgrad(du, I, P16 + o[MAGB200_LOFF_F1_W], H, T, H, I, EPI_STORE_F32, w.dr32, w.dr32, H, st));
    // BertSelfOutput: LayerNorm(dropout(dense(ctx)) + x)
    RUN(ln_bwd(w.dr32, L.a32, P + o[MAGB200_LOFF_LN1_W], P + o[MAGB200_LOFF_LN1_B], L.rstd1, d.p_hidden, sd,
               site_attn_out(l), T, us, w.dy32, dpre_a, G + o[MAGB200_LOFF_LN1_W], G + o[MAGB200_LOFF_LN1_B],
               G + o[MAGB200_LOFF_AO_B], st));
    RUN(fork());
    RUN(linear_wgrad(dpre_a, H, L.ctx16, H, T, H, H, G + o[MAGB200_LOFF_AO_W], H, 4, us, sst));
    RUN(linear_dgrad(dpre_a, H, P16 + o[MAGB200_LOFF_AO_W], H, T, H, H, EPI_STORE_F16, w.dctx16, nullptr, H, st));
    // BertSelfAttention
    RUN(sdpa_bwd(L.qkv16, w.dctx16, L.ctx16, L.lse, mask, B, S, d.heads, H / d.heads, d.p_attn, sd, site_attn_prob(l), us,
                 dqkv, G + o[MAGB200_LOFF_QKV_B], w.attD, st));
    RUN(fork());
    RUN(linear_wgrad(dqkv, 3 * H, w.x16[l], H, T, 3 * H, H, G + o[MAGB200_LOFF_QKV_W], H, 2, us, sst));
    RUN(linear_dgrad(dqkv, 3 * H, P16 + o[MAGB200_LOFF_QKV_W], H, T, H, 3 * H, EPI_STORE_F32, w.dy32, w.dy32, H, st));
    if (side) {
      MAGB_CHECK_CUDA(cudaEventRecord(sc.done[par], sst));
      side_used = true;
    }
    ++layers_done;
    if (opt) {
      if (side) {
        MAGB_CHECK_CUDA(cudaEventRecord(sc.lay, st));
        MAGB_CHECK_CUDA(cudaStreamWaitEvent(ost, sc.lay, 0));
        MAGB_CHECK_CUDA(cudaStreamWaitEvent(ost, sc.done[par], 0));
        opt_used = true;
      }
      RUN(adamw_range(layer_start(l), layer_start(l + 1), ost));
    }
  }
  // join: the caller may all-reduce / consume this piece's gradients as soon as `st` gets here
  if (side && side_used) {
    MAGB_CHECK_CUDA(cudaEventRecord(sc.join, sst));
    MAGB_CHECK_CUDA(cudaStreamWaitEvent(st, sc.join, 0));
  }

  if (!(flags & 2)) {
    if (opt && opt_used) {
      MAGB_CHECK_CUDA(cudaEventRecord(sc.optjoin, ost));
      MAGB_CHECK_CUDA(cudaStreamWaitEvent(st, sc.optjoin, 0));
    }
    return 0;
  }
  // MAG (modeling.py:25-51)
  RUN(mag_tail_bwd(w.dy32, w.G32, w.emb32, P + off[MAGB200_OFF_MAG_LN_W], d.beta_shift, d.mag_ln_eps, d.p_mag, sd,
                   SITE_MAG, T, us, w.dG16, w.dr32, G + off[MAGB200_OFF_MAG_LN_W], G + off[MAGB200_OFF_MAG_LN_B],
                   G + off[MAGB200_OFF_MAG_BIAS4], st));
  MAGB_CHECK_CUDA(cudaMemsetAsync(w.dBmag32, 0, (size_t)4 * H * d.ka * sizeof(float), st));
  RUN(linear_wgrad(w.dG16, 4 * H, w.amag16, d.ka, T, 4 * H, d.ka, w.dBmag32, d.ka, 1, us, st));
  RUN(mag_unpack_wgrad(w.dBmag32, d.visual_dim, d.acoustic_dim, d.ka, G + off[MAGB200_OFF_MAG_WHV],
                       G + off[MAGB200_OFF_MAG_WHA], G + off[MAGB200_OFF_MAG_WV], G + off[MAGB200_OFF_MAG_WA], st));
  // text gradient through the two gate Linears: dG[:, :2H] * Bpacked[:2H, :H]
  RUN(linear_dgrad(w.dG16, 4 * H, P16 + off[MAGB200_OFF_MAG_PACKED16], d.ka, T, H, 2 * H, EPI_STORE_F32, w.dr32, w.dr32,
                   H, st));
  // BertEmbeddings
  EmbedArgs e;
  e.ids = ids; e.tt = tt;
  e.word = P + off[MAGB200_OFF_WORD]; e.pos = P + off[MAGB200_OFF_POS]; e.type = P + off[MAGB200_OFF_TYPE];
  e.gamma = P + off[MAGB200_OFF_EMB_LN_W]; e.beta = P + off[MAGB200_OFF_EMB_LN_B];
  e.eps = d.ln_eps; e.p = d.p_hidden; e.seed = sd; e.site = SITE_EMB; e.T = T; e.S = S; e.vocab = d.vocab;
  e.type_vocab = d.type_vocab;
  RUN(embed_ln_bwd(e, w.dr32, us, G + off[MAGB200_OFF_WORD], G + off[MAGB200_OFF_POS], G + off[MAGB200_OFF_TYPE],
                   G + off[MAGB200_OFF_EMB_LN_W], G + off[MAGB200_OFF_EMB_LN_B], st));
  if (opt) {
    // what no layer slice covers: the embedding tables in front, pooler / MAG / classifier behind
    RUN(adamw_range(0, layer_start(0), st));
    RUN(adamw_range(off[MAGB200_OFF_POOL_W], off[MAGB200_OFF_TOTAL], st));
    if (opt_used) {
      MAGB_CHECK_CUDA(cudaEventRecord(sc.optjoin, ost));
      MAGB_CHECK_CUDA(cudaStreamWaitEvent(st, sc.optjoin, 0));
    }
  }
  return 0;
}

int bert_attention_probs(const magb200_bert_dims& d, void* workspace, int layer, const int64_t* mask, float* out,
                         size_t out_bytes, cudaStream_t st) {
  RUN(check_dims(d));
  MAGB_REQUIRE(layer >= 0 && layer < d.layers, "attention_probs: bad layer %d", layer);
  BertWs w = carve_ws(d, workspace);
  const size_t n = (size_t)d.batch * d.heads * d.seq * d.seq * sizeof(float);
  MAGB_REQUIRE(out_bytes >= n, "attention_probs: output buffer too small (%zu < %zu)", out_bytes, n);
  return sdpa_probs(w.layer[layer].qkv16, mask, d.batch, d.seq, d.heads, d.p_attn, Seed{d.seed, d.seed_dev},
                    site_attn_prob(layer), out, st);
}

// Intermediate activations for tests (copies device->device into caller buffers; null = skip)
int bert_peek(const magb200_bert_dims& d, void* workspace, int what, int layer, void* out, size_t out_bytes,
              cudaStream_t st) {
  BertWs w = carve_ws(d, workspace);
  const size_t T = (size_t)d.batch * d.seq, H = d.hidden;
  const void* src = nullptr;
  size_t n = 0;
  switch (what) {
    case MAGB200_PEEK_EMB: src = w.emb32; n = T * H * 4; break;
    case MAGB200_PEEK_MAG: src = w.x32[0]; n = T * H * 4; break;
    case MAGB200_PEEK_LAYER_OUT:
      MAGB_REQUIRE(layer >= 0 && layer < d.layers, "peek: bad layer");
      src = w.x32[layer + 1]; n = T * H * 4; break;
    case MAGB200_PEEK_POOLED: src = w.pooled32; n = (size_t)d.batch * H * 4; break;
    case MAGB200_PEEK_GATES: src = w.G32; n = T * 4 * H * 4; break;
    default: MAGB_REQUIRE(false, "peek: unknown item %d", what);
  }
  MAGB_REQUIRE(out_bytes >= n, "peek: output buffer too small");
  MAGB_CHECK_CUDA(cudaMemcpyAsync(out, src, n, cudaMemcpyDeviceToDevice, st));
  return 0;
}

}  // namespace magb200

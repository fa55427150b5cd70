// MAG-XLNet forward / backward schedules:
//   MAG_XLNetForSequenceClassification.forward (xlnet.py:443-527) -> MAG_XLNetModel.forward (xlnet.py:148-429)
//   -> word embedding, relative positional encoding, 12 x XLNetLayer (XLNetRelativeAttention +
//      XLNetFeedForward, transformers 3.0.2) with MAG (modeling.py:25-51) applied to the content stream
//      before layer XLNET_INJECTION_INDEX (xlnet.py:371-372), SequenceSummary('last', tanh) + logits_proj.
// Content stream only (target_mapping None -> output_g None, xlnet.py:306-313).  Same conventions as
// model_bert.cu: flat parameter / gradient buffers, caller-owned workspace, loss-scaled fp16 activation
// gradients, unscaled fp32 parameter gradients.
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

struct XLayerWs {
  __half* qkv16; __half* kr16; __half* ctx16; float* a32; __half* a16; __half* gp16; __half* g16; float* rstd1; float* rstd2;
};
struct XlWs {
  __half* pos16; __half* amag16; float* G32; float* mag32; __half* mag16;
  float* y32[MAGB200_MAX_LAYERS + 1];
  __half* y16[MAGB200_MAX_LAYERS + 1];
  XLayerWs layer[MAGB200_MAX_LAYERS];
  float* pre32; __half* last16; float* sumpre32; float* pooled32;
  float* dy32; float* dr32; __half* dpre16; __half* du16; __half* dctx16; __half* dqkv16; float* dkr32; __half* dkr16;
  __half* dG16; float* dBmag32; __half* dpool16; float* dlast32;
  size_t bytes;
};

inline int pos_batch(const magb200_xlnet_dims& d) { return d.p_drop > 0.f ? d.batch : 1; }

XlWs carve_ws(const magb200_xlnet_dims& d, void* base) {
  XlWs w;
  Carve c(base);
  const size_t T = (size_t)d.batch * d.seq, H = d.hidden, I = d.inter;
  const size_t PR = (size_t)d.batch * 2 * d.seq;     // worst case rows of the position tables
  w.pos16 = c.take<__half>(PR * H);
  w.amag16 = c.take<__half>(T * d.ka);
  w.G32 = c.take<float>(T * 4 * H);
  w.mag32 = c.take<float>(T * H);
  w.mag16 = c.take<__half>(T * H);
  for (int l = 0; l <= d.layers; ++l) {
    w.y32[l] = c.take<float>(T * H);
    w.y16[l] = c.take<__half>(T * H);
  }
  for (int l = 0; l < d.layers; ++l) {
    XLayerWs& L = w.layer[l];
    L.qkv16 = c.take<__half>(T * 3 * H);
    L.kr16 = c.take<__half>(PR * H);
    L.ctx16 = c.take<__half>(T * H);
    L.a32 = c.take<float>(T * H);
    L.a16 = c.take<__half>(T * H);
    L.gp16 = c.take<__half>(T * I);
    L.g16 = c.take<__half>(T * I);
    L.rstd1 = c.take<float>(T);
    L.rstd2 = c.take<float>(T);
  }
  w.pre32 = c.take<float>(T * H);
  w.last16 = c.take<__half>((size_t)d.batch * H);
  w.sumpre32 = c.take<float>((size_t)d.batch * H);
  w.pooled32 = c.take<float>((size_t)d.batch * H);
  w.dy32 = c.take<float>(T * H);
  w.dr32 = c.take<float>(T * H);
  w.dpre16 = c.take<__half>(T * H);
  w.du16 = c.take<__half>(T * I);
  w.dctx16 = c.take<__half>(T * H);
  w.dqkv16 = c.take<__half>(T * 3 * H);
  w.dkr32 = c.take<float>(PR * H);
  w.dkr16 = c.take<__half>(PR * H);
  w.dG16 = c.take<__half>(T * 4 * H);
  w.dBmag32 = c.take<float>((size_t)4 * H * d.ka);
  w.dpool16 = c.take<__half>((size_t)d.batch * H);
  w.dlast32 = c.take<float>((size_t)d.batch * H);
  w.bytes = (c.off + 1023) & ~size_t(1023);
  return w;
}

int check_dims(const magb200_xlnet_dims& d) {
  MAGB_REQUIRE(d.hidden == 768 && d.heads == 12, "xlnet: kernels are specialised for d_model=768, 12 heads (got %d, %d)",
               d.hidden, d.heads);
  MAGB_REQUIRE(d.inter % 128 == 0 && d.inter > 0, "xlnet: d_inner %d must be a multiple of 128", d.inter);
  MAGB_REQUIRE(d.layers >= 1 && d.layers <= MAGB200_MAX_LAYERS, "xlnet: layers=%d out of range", d.layers);
  MAGB_REQUIRE(d.batch >= 1 && d.seq >= 1 && d.seq <= 64, "xlnet: seq=%d out of range [1,64] (relative attention kernel)", d.seq);
  MAGB_REQUIRE(d.ka == d.hidden + 128 && d.visual_dim + d.acoustic_dim <= 128 && d.visual_dim > 0 && d.acoustic_dim > 0,
               "xlnet: visual_dim + acoustic_dim = %d must be <= 128", d.visual_dim + d.acoustic_dim);
  MAGB_REQUIRE(d.num_labels >= 1 && d.num_labels <= 32, "xlnet: num_labels=%d out of range", d.num_labels);
  // injection_index >= layers simply never applies MAG, exactly like the loop at xlnet.py:362-372
  MAGB_REQUIRE(d.injection_index >= 0, "xlnet: injection index %d out of range", d.injection_index);
  return 0;
}

enum : uint32_t { SITE_EMB = 1, SITE_MAG = 2, SITE_CLS = 3, SITE_POS = 4, SITE_FINAL = 5, SITE_LAYER0 = 16 };
inline uint32_t site_attn_prob(int l) { return SITE_LAYER0 + 4 * l; }
inline uint32_t site_attn_out(int l) { return SITE_LAYER0 + 4 * l + 1; }
inline uint32_t site_ffn_out(int l) { return SITE_LAYER0 + 4 * l + 2; }
inline uint32_t site_ffn_mid(int l) { return SITE_LAYER0 + 4 * l + 3; }

#define RUN(expr)        \
  do {                   \
    int _rc = (expr);    \
    if (_rc) return _rc; \
  } while (0)

GemmProblem gp(const void* A, int lda, bool a_mn, const void* B, int ldb, bool b_mn, int M, int N, int K, int epi,
               const float* bias, void* out0, int ldo) {
  GemmProblem p;
  p.A = A; p.lda = lda; p.a_mn = a_mn; p.B = B; p.ldb = ldb; p.b_mn = b_mn; p.M = M; p.N = N; p.K = K;
  p.a_dtype = DT_F16; p.b_dtype = DT_F16; p.epi = epi; p.bias = bias; p.out0 = out0; p.ldo = ldo;
  return p;
}

RelAttnArgs rel_args(const magb200_xlnet_dims& d, const int64_t* off, const float* P, const XLayerWs& L, int l,
                     const int64_t* mask, const int64_t* tt, Seed sd) {
  const int64_t* o = off + MAGB200_XOFF_LAYER0 + (int64_t)l * MAGB200_XOFF_PER_LAYER;
  RelAttnArgs a;
  a.qkv = L.qkv16; a.kr16 = L.kr16; a.Bp = pos_batch(d);
  a.r_w_bias = P + o[MAGB200_XLOFF_RW_BIAS]; a.r_r_bias = P + o[MAGB200_XLOFF_RR_BIAS];
  a.r_s_bias = P + o[MAGB200_XLOFF_RS_BIAS]; a.seg_embed = P + o[MAGB200_XLOFF_SEG];
  a.mask = mask; a.tt = tt; a.B = d.batch; a.S = d.seq; a.heads = d.heads; a.head_dim = d.hidden / d.heads;
  a.p = d.p_drop; a.seed = sd; a.site = site_attn_prob(l);
  return a;
}

}  // namespace

size_t xlnet_workspace_bytes(const magb200_xlnet_dims& d) { return carve_ws(d, nullptr).bytes; }

int xlnet_refresh_operands(const magb200_xlnet_dims& d, const int64_t* off, const float* params, __half* params16,
                           cudaStream_t st) {
  RUN(check_dims(d));
  RUN(cast_f32_f16(params, params16, off[MAGB200_XOFF_TOTAL], st));
  RUN(mag_pack_weights(params + off[MAGB200_XOFF_MAG_WHV], params + off[MAGB200_XOFF_MAG_WHA],
                       params + off[MAGB200_XOFF_MAG_WV], params + off[MAGB200_XOFF_MAG_WA], d.visual_dim,
                       d.acoustic_dim, d.ka, params16 + off[MAGB200_XOFF_MAG_PACKED16], st));
  for (int l = 0; l < d.layers; ++l) {
    const int64_t* o = off + MAGB200_XOFF_LAYER0 + (int64_t)l * MAGB200_XOFF_PER_LAYER;
    RUN(xl_pack_qkv(params + o[MAGB200_XLOFF_Q], params + o[MAGB200_XLOFF_K], params + o[MAGB200_XLOFF_V],
                    params16 + o[MAGB200_XLOFF_QKV_PACKED16], st));
  }
  return 0;
}

int xlnet_forward(const magb200_xlnet_dims& d, const int64_t* off, const float* P, const __half* P16, const int64_t* ids,
                  const int64_t* tt, const int64_t* mask, const float* visual, const float* acoustic, void* workspace,
                  size_t ws_bytes, float* logits, cudaStream_t st) {
  RUN(check_dims(d));
  XlWs w = carve_ws(d, workspace);
  MAGB_REQUIRE(ws_bytes >= w.bytes, "xlnet_forward: workspace too small (%zu < %zu)", ws_bytes, w.bytes);
  const int B = d.batch, S = d.seq, T = B * S, H = d.hidden, I = d.inter;
  const Seed sd{d.seed, d.seed_dev};
  const int Bp = pos_batch(d), PR = Bp * 2 * S;

  RUN(xl_gather_fwd(ids, P + off[MAGB200_XOFF_WORD], T, d.vocab, d.p_drop, sd, SITE_EMB, w.y32[0], w.y16[0], st));   // xlnet.py:304-305
  RUN(xl_pos_emb(S, Bp, d.p_drop, sd, SITE_POS, w.pos16, st));                                                    // xlnet.py:332-333
  for (int l = 0; l < d.layers; ++l) {
    const int64_t* o = off + MAGB200_XOFF_LAYER0 + (int64_t)l * MAGB200_XOFF_PER_LAYER;
    XLayerWs& L = w.layer[l];
    const float* h32 = w.y32[l];
    const __half* h16 = w.y16[l];
    if (l == d.injection_index) {                                                                                   // xlnet.py:371-372
      RUN(mag_pack_inputs(w.y32[l], visual, acoustic, d.visual_dim, d.acoustic_dim, d.ka, T, w.amag16, st));
      GemmProblem g = gp(w.amag16, d.ka, false, P16 + off[MAGB200_XOFF_MAG_PACKED16], d.ka, false, T, 4 * H, d.ka,
                         EPI_STORE_F32, P + off[MAGB200_XOFF_MAG_BIAS4], w.G32, 4 * H);
      g.n_split = 2 * H; g.k_start_hi = H;
      RUN(gemm_launch(g, st));
      RUN(mag_tail_fwd(w.G32, w.y32[l], P + off[MAGB200_XOFF_MAG_LN_W], P + off[MAGB200_XOFF_MAG_LN_B], d.beta_shift,
                       d.mag_ln_eps, d.p_mag, sd, SITE_MAG, T, w.mag32, w.mag16, st));
      h32 = w.mag32;
      h16 = w.mag16;
    }
    // XLNetRelativeAttention: head projections (no biases), positions, core, output projection + LN
    RUN(gemm_launch(gp(h16, H, false, P16 + o[MAGB200_XLOFF_QKV_PACKED16], 3 * H, true, T, 3 * H, H, EPI_STORE_F16, nullptr,
                       L.qkv16, 3 * H), st));
    RUN(gemm_launch(gp(w.pos16, H, false, P16 + o[MAGB200_XLOFF_R], H, true, PR, H, H, EPI_STORE_F16, nullptr, L.kr16, H),
                    st));
    RUN(xl_rel_attn_fwd(rel_args(d, off, P, L, l, mask, tt, sd), L.ctx16, st));
    RUN(gemm_launch(gp(L.ctx16, H, false, P16 + o[MAGB200_XLOFF_O], H, false, T, H, H, EPI_STORE_F32, nullptr, w.pre32, H),
                    st));
    RUN(dropout_residual_ln_fwd(w.pre32, h32, P + o[MAGB200_XLOFF_LN1_W], P + o[MAGB200_XLOFF_LN1_B], d.ln_eps, d.p_drop,
                                sd, site_attn_out(l), T, L.a32, L.a16, L.rstd1, st));
    // XLNetFeedForward: layer_1 -> gelu -> dropout -> layer_2 -> dropout -> +residual -> LN
    GemmProblem f1 = gp(L.a16, H, false, P16 + o[MAGB200_XLOFF_F1_W], H, false, T, I, H, EPI_GELU_F16,
                        P + o[MAGB200_XLOFF_F1_B], L.g16, I);
    f1.out1 = L.gp16; f1.drop_p = d.p_drop; f1.drop_seed = sd; f1.drop_site = site_ffn_mid(l);
    RUN(gemm_launch(f1, st));
    RUN(gemm_launch(gp(L.g16, I, false, P16 + o[MAGB200_XLOFF_F2_W], I, false, T, H, I, EPI_STORE_F32,
                       P + o[MAGB200_XLOFF_F2_B], w.pre32, H), st));
    RUN(dropout_residual_ln_fwd(w.pre32, L.a32, P + o[MAGB200_XLOFF_LN2_W], P + o[MAGB200_XLOFF_LN2_B], d.ln_eps, d.p_drop,
                                sd, site_ffn_out(l), T, w.y32[l + 1], w.y16[l + 1], L.rstd2, st));
  }
  // output dropout (xlnet.py:396), SequenceSummary 'last' + tanh + dropout, logits_proj (xlnet.py:506-509)
  RUN(xl_last_token_fwd(w.y32[d.layers], B, S, d.p_drop, sd, SITE_FINAL, w.last16, st));
  RUN(gemm_launch(gp(w.last16, H, false, P16 + off[MAGB200_XOFF_SUM_W], H, false, B, H, H, EPI_STORE_F32,
                     P + off[MAGB200_XOFF_SUM_B], w.sumpre32, H), st));
  RUN(cls_fwd(w.sumpre32, P + off[MAGB200_XOFF_CLS_W], P + off[MAGB200_XOFF_CLS_B], B, d.num_labels, d.p_summary, sd,
              SITE_CLS, w.pooled32, logits, st));
  return 0;
}

int xlnet_backward(const magb200_xlnet_dims& d, const int64_t* off, const float* P, const __half* P16, const int64_t* ids,
                   const int64_t* tt, const int64_t* mask, void* workspace, size_t ws_bytes, const float* dlogits_scaled,
                   float* G, cudaStream_t st) {
  RUN(check_dims(d));
  XlWs w = carve_ws(d, workspace);
  MAGB_REQUIRE(ws_bytes >= w.bytes, "xlnet_backward: workspace too small (%zu < %zu)", ws_bytes, w.bytes);
  const int B = d.batch, S = d.seq, T = B * S, H = d.hidden, I = d.inter;
  const float us = 1.0f / d.grad_scale;
  const Seed sd{d.seed, d.seed_dev};
  const int Bp = pos_batch(d), PR = Bp * 2 * S;

  auto wgrad = [&](const __half* dY, int ldy, const __half* X, int ldx, int K, int Nout, int Kin, float* dW, int lddw,
                   int splits) {
    GemmProblem p = gp(dY, ldy, true, X, ldx, true, Nout, Kin, K, EPI_ATOMIC_F32, nullptr, dW, lddw);
    p.splits = splits; p.alpha = us;
    return gemm_launch(p, st);
  };

  // head
  RUN(cls_bwd(dlogits_scaled, w.pooled32, P + off[MAGB200_XOFF_CLS_W], B, d.num_labels, d.p_summary, sd, SITE_CLS, us,
              w.dpool16, G + off[MAGB200_XOFF_CLS_W], G + off[MAGB200_XOFF_CLS_B], st));
  RUN(gemm_launch(gp(w.dpool16, H, false, P16 + off[MAGB200_XOFF_SUM_W], H, true, B, H, H, EPI_STORE_F32, nullptr,
                     w.dlast32, H), st));
  RUN(wgrad(w.dpool16, H, w.last16, H, B, H, H, G + off[MAGB200_XOFF_SUM_W], H, 1));
  RUN(colsum_f16(w.dpool16, B, H, us, G + off[MAGB200_XOFF_SUM_B], st));
  RUN(xl_last_token_bwd(w.dlast32, B, S, d.p_drop, sd, SITE_FINAL, w.dy32, st));

  for (int l = d.layers - 1; l >= 0; --l) {
    const int64_t* o = off + MAGB200_XOFF_LAYER0 + (int64_t)l * MAGB200_XOFF_PER_LAYER;
    XLayerWs& L = w.layer[l];
    const bool inj = (l == d.injection_index);
    const float* h32 = inj ? w.mag32 : w.y32[l];
    const __half* h16 = inj ? w.mag16 : w.y16[l];
    (void)h32;
    // XLNetFeedForward
    RUN(ln_bwd(w.dy32, w.y32[l + 1], P + o[MAGB200_XLOFF_LN2_W], P + o[MAGB200_XLOFF_LN2_B], L.rstd2, d.p_drop, sd,
               site_ffn_out(l), T, us, w.dr32, w.dpre16, G + o[MAGB200_XLOFF_LN2_W], G + o[MAGB200_XLOFF_LN2_B],
               G + o[MAGB200_XLOFF_F2_B], st));
    RUN(wgrad(w.dpre16, H, L.g16, I, T, H, I, G + o[MAGB200_XLOFF_F2_W], I, 1));
    GemmProblem dg = gp(w.dpre16, H, false, P16 + o[MAGB200_XLOFF_F2_W], I, true, T, I, H, EPI_DGELU_F16, nullptr, w.du16, I);
    dg.aux = L.gp16; dg.drop_p = d.p_drop; dg.drop_seed = sd; dg.drop_site = site_ffn_mid(l);
    RUN(gemm_launch(dg, st));
    RUN(colsum_f16(w.du16, T, I, us, G + o[MAGB200_XLOFF_F1_B], st));
    RUN(wgrad(w.du16, I, L.a16, H, T, I, H, G + o[MAGB200_XLOFF_F1_W], H, 1));
    GemmProblem d1 = gp(w.du16, I, false, P16 + o[MAGB200_XLOFF_F1_W], H, true, T, H, I, EPI_STORE_F32, nullptr, w.dr32, H);
    d1.aux = w.dr32;
    RUN(gemm_launch(d1, st));
    // XLNetRelativeAttention.post_attention: LN(dropout(o-projection) + h)   (no bias on the projection)
    RUN(ln_bwd(w.dr32, L.a32, P + o[MAGB200_XLOFF_LN1_W], P + o[MAGB200_XLOFF_LN1_B], L.rstd1, d.p_drop, sd,
               site_attn_out(l), T, us, w.dy32, w.dpre16, G + o[MAGB200_XLOFF_LN1_W], G + o[MAGB200_XLOFF_LN1_B], nullptr,
               st));
    RUN(wgrad(w.dpre16, H, L.ctx16, H, T, H, H, G + o[MAGB200_XLOFF_O], H, 4));
    RUN(gemm_launch(gp(w.dpre16, H, false, P16 + o[MAGB200_XLOFF_O], H, true, T, H, H, EPI_STORE_F16, nullptr, w.dctx16, H),
                    st));
    // relative attention core
    MAGB_CHECK_CUDA(cudaMemsetAsync(w.dkr32, 0, (size_t)PR * H * sizeof(float), st));
    RUN(xl_rel_attn_bwd(rel_args(d, off, P, L, l, mask, tt, sd), w.dctx16, us, w.dqkv16, w.dkr32,
                        G + o[MAGB200_XLOFF_RW_BIAS], G + o[MAGB200_XLOFF_RR_BIAS], G + o[MAGB200_XLOFF_RS_BIAS],
                        G + o[MAGB200_XLOFF_SEG], st));
    RUN(cast_f32_f16(w.dkr32, w.dkr16, (int64_t)PR * H, st));
    // projections: parameters are [in][out]  =>  dW[in][out] += X^T dY
    RUN(wgrad(w.pos16, H, w.dkr16, H, PR, H, H, G + o[MAGB200_XLOFF_R], H, 1));
    RUN(wgrad(h16, H, w.dqkv16, 3 * H, T, H, H, G + o[MAGB200_XLOFF_Q], H, 4));
    RUN(wgrad(h16, H, w.dqkv16 + H, 3 * H, T, H, H, G + o[MAGB200_XLOFF_K], H, 4));
    RUN(wgrad(h16, H, w.dqkv16 + 2 * H, 3 * H, T, H, H, G + o[MAGB200_XLOFF_V], H, 4));
    // dh += dqkv [T,2304] . Wqkv^T : the packed [in][q|k|v] operand read K-major (K = 2304)
    GemmProblem dh = gp(w.dqkv16, 3 * H, false, P16 + o[MAGB200_XLOFF_QKV_PACKED16], 3 * H, false, T, H, 3 * H, EPI_STORE_F32,
                        nullptr, w.dy32, H);
    dh.aux = w.dy32;
    RUN(gemm_launch(dh, st));
    if (inj) {
      // MAG backward: dy32 = gradient wrt the MAG output; produce the gradient wrt y32[l] (the text input)
      RUN(mag_tail_bwd(w.dy32, w.G32, w.y32[l], P + off[MAGB200_XOFF_MAG_LN_W], d.beta_shift, d.mag_ln_eps, d.p_mag, sd,
                       SITE_MAG, T, us, w.dG16, w.dr32, G + off[MAGB200_XOFF_MAG_LN_W], G + off[MAGB200_XOFF_MAG_LN_B],
                       G + off[MAGB200_XOFF_MAG_BIAS4], st));
      MAGB_CHECK_CUDA(cudaMemsetAsync(w.dBmag32, 0, (size_t)4 * H * d.ka * sizeof(float), st));
      RUN(wgrad(w.dG16, 4 * H, w.amag16, d.ka, T, 4 * H, d.ka, w.dBmag32, d.ka, 1));
      RUN(mag_unpack_wgrad(w.dBmag32, d.visual_dim, d.acoustic_dim, d.ka, G + off[MAGB200_XOFF_MAG_WHV],
                           G + off[MAGB200_XOFF_MAG_WHA], G + off[MAGB200_XOFF_MAG_WV], G + off[MAGB200_XOFF_MAG_WA], st));
      GemmProblem dt = gp(w.dG16, 4 * H, false, P16 + off[MAGB200_XOFF_MAG_PACKED16], d.ka, true, T, H, 2 * H, EPI_STORE_F32,
                          nullptr, w.dy32, H);
      dt.aux = w.dr32;
      RUN(gemm_launch(dt, st));
    }
  }
  RUN(xl_gather_bwd(w.dy32, ids, T, d.vocab, d.p_drop, sd, SITE_EMB, us, G + off[MAGB200_XOFF_WORD], st));
  return 0;
}

int xlnet_peek(const magb200_xlnet_dims& d, void* workspace, int what, int layer, void* out, size_t out_bytes,
               cudaStream_t st) {
  XlWs w = carve_ws(d, workspace);
  const size_t T = (size_t)d.batch * d.seq, H = d.hidden;
  const void* src = nullptr;
  size_t n = 0;
  switch (what) {
    case MAGB200_PEEK_MAG: src = w.mag32; n = T * H * 4; break;
    case MAGB200_PEEK_LAYER_OUT:
      MAGB_REQUIRE(layer >= 0 && layer < d.layers, "peek: bad layer");
      src = w.y32[layer + 1]; n = T * H * 4; break;
    case MAGB200_PEEK_POOLED: src = w.pooled32; n = (size_t)d.batch * H * 4; break;
    default: MAGB_REQUIRE(false, "peek: unknown item %d", what);
  }
  MAGB_REQUIRE(out_bytes >= n, "peek: output buffer too small");
  MAGB_CHECK_CUDA(cudaMemcpyAsync(out, src, n, cudaMemcpyDeviceToDevice, st));
  return 0;
}

}  // namespace magb200

// extern "C" boundary for the row-wise operators, attention, head, optimizer and the whole-model schedules.
#include "common.cuh"
#include "gemm_sm100.cuh"
#include "host.h"
#include "../../include/magb200.h"

using namespace magb200;

#define ST reinterpret_cast<cudaStream_t>(stream)
#define RUN(expr)        \
  do {                   \
    int _rc = (expr);    \
    if (_rc) return _rc; \
  } while (0)

extern "C" {

int magb200_dropout_mask(uint8_t* out, int64_t n, float p, uint64_t seed, uint32_t site, void* stream) {
  return dropout_mask(out, n, p, Seed{seed, nullptr}, site, ST);
}

int magb200_cast_f32_f16(const float* in, void* out_f16, int64_t n, void* stream) {
  return cast_f32_f16(in, reinterpret_cast<__half*>(out_f16), n, ST);
}

int magb200_embed_ln_fwd(const int64_t* ids, const int64_t* token_type_ids, const float* visual, const float* acoustic,
                         const float* word, const float* pos, const float* type, const float* gamma, const float* beta,
                         float eps, float p, uint64_t seed, uint32_t site, int V, int A, int ka, int T, int S, int vocab,
                         float* emb32, void* amag16, void* stream) {
  EmbedArgs e;
  e.ids = ids; e.tt = token_type_ids; e.visual = visual; e.acoustic = acoustic; e.word = word; e.pos = pos;
  e.type = type; e.gamma = gamma; e.beta = beta; e.eps = eps; e.p = p; e.seed = Seed{seed, nullptr}; e.site = site; e.V = V; e.A = A;
  e.KA = ka; e.T = T; e.S = S; e.vocab = vocab; e.emb32 = emb32; e.amag = reinterpret_cast<__half*>(amag16);
  return embed_ln_fwd(e, ST);
}

int magb200_embed_ln_bwd(const float* dout, const int64_t* ids, const int64_t* token_type_ids, const float* word,
                         const float* pos, const float* type, const float* gamma, float eps, float p, uint64_t seed,
                         uint32_t site, int T, int S, int vocab, float out_scale, float* dword, float* dpos, float* dtype,
                         float* dgamma, float* dbeta, void* stream) {
  EmbedArgs e;
  e.ids = ids; e.tt = token_type_ids; e.word = word; e.pos = pos; e.type = type; e.gamma = gamma; e.eps = eps; e.p = p;
  e.seed = Seed{seed, nullptr}; e.site = site; e.T = T; e.S = S; e.vocab = vocab;
  return embed_ln_bwd(e, dout, out_scale, dword, dpos, dtype, dgamma, dbeta, ST);
}

int magb200_mag_pack_weights(const float* W_hv, const float* W_ha, const float* W_v, const float* W_a, int V, int A,
                             int ka, void* packed16, void* stream) {
  return mag_pack_weights(W_hv, W_ha, W_v, W_a, V, A, ka, reinterpret_cast<__half*>(packed16), ST);
}

int magb200_mag_fwd(const float* text32, const float* visual, const float* acoustic, const void* packed16,
                    const float* bias4, const float* ln_gamma, const float* ln_beta, float beta_shift, float ln_eps,
                    float p, uint64_t seed, uint32_t site, int V, int A, int ka, int T, void* amag16, float* gates32,
                    float* out32, void* out16, void* stream) {
  const int H = 768;
  MAGB_REQUIRE(ka == H + 128, "mag: ka must be 896");
  RUN(mag_pack_inputs(text32, visual, acoustic, V, A, ka, T, reinterpret_cast<__half*>(amag16), ST));
  GemmProblem g;
  g.A = amag16; g.B = packed16; g.M = T; g.N = 4 * H; g.K = ka; g.lda = ka; g.ldb = ka;
  g.a_dtype = DT_F16; g.b_dtype = DT_F16; g.epi = EPI_STORE_F32; g.bias = bias4; g.out0 = gates32; g.ldo = 4 * H;
  g.n_split = 2 * H; g.k_start_hi = H;
  RUN(gemm_launch(g, ST));
  return mag_tail_fwd(gates32, text32, ln_gamma, ln_beta, beta_shift, ln_eps, p, Seed{seed, nullptr}, site, T, out32,
                      reinterpret_cast<__half*>(out16), ST);
}

size_t magb200_embed_mag_scratch_bytes(int T) { return mag_fused_scratch_bytes(T); }

int magb200_embed_mag_fwd(const int64_t* ids, const int64_t* token_type_ids, const float* visual, const float* acoustic,
                          const float* word, const float* pos, const float* type, const float* emb_gamma,
                          const float* emb_beta, float emb_eps, float p_emb, uint64_t seed, uint32_t site_emb, int V, int A,
                          int ka, int T, int S, int vocab, const void* packed16, const float* bias4, const float* mag_gamma,
                          const float* mag_beta, float beta_shift, float mag_eps, float p_mag, uint32_t site_mag,
                          float* emb32, void* amag16, float* gates32, void* scratch, size_t scratch_bytes, float* out32,
                          void* out16, void* stream) {
  MAGB_REQUIRE(scratch_bytes >= mag_fused_scratch_bytes(T), "embed_mag_fwd: scratch too small (%zu < %zu)", scratch_bytes,
               mag_fused_scratch_bytes(T));
  EmbedArgs e;
  e.ids = ids; e.tt = token_type_ids; e.visual = visual; e.acoustic = acoustic; e.word = word; e.pos = pos; e.type = type;
  e.gamma = emb_gamma; e.beta = emb_beta; e.eps = emb_eps; e.p = p_emb; e.seed = Seed{seed, nullptr}; e.site = site_emb;
  e.V = V; e.A = A; e.KA = ka; e.T = T; e.S = S; e.vocab = vocab; e.emb32 = emb32; e.amag = reinterpret_cast<__half*>(amag16);
  return mag_fused_fwd(e, reinterpret_cast<const __half*>(packed16), bias4, mag_gamma, mag_beta, beta_shift, mag_eps, p_mag,
                       site_mag, gates32, scratch, out32, reinterpret_cast<__half*>(out16), ST);
}

int magb200_mag_fwd_exact(const float* text32, const float* visual, const float* acoustic, const float* W_hv,
                          const float* W_ha, const float* W_v, const float* W_a, const float* bias4,
                          const float* ln_gamma, const float* ln_beta, float beta_shift, float ln_eps, float p,
                          uint64_t seed, uint32_t site, int V, int A, int ka, int T, float* split_scratch,
                          float* gates32, float* out32, void* out16, void* stream) {
  const int H = 768;
  MAGB_REQUIRE(ka == H + 128, "mag: ka must be 896");
  MAGB_REQUIRE(split_scratch != nullptr, "mag_fwd_exact: scratch of 2*(T + 4*768)*ka floats required");
  float* a_hi = split_scratch;
  float* a_lo = a_hi + (size_t)T * ka;
  float* b_hi = a_lo + (size_t)T * ka;
  float* b_lo = b_hi + (size_t)4 * H * ka;
  RUN(mag_pack_split(text32, visual, acoustic, W_hv, W_ha, W_v, W_a, V, A, ka, T, a_hi, a_lo, b_hi, b_lo, ST));
  // gates = A_hi B_hi^T + bias, += A_hi B_lo^T, += A_lo B_hi^T   (three TF32 GEMMs, fp32 accumulation)
  const float* As[3] = {a_hi, a_hi, a_lo};
  const float* Bs[3] = {b_hi, b_lo, b_hi};
  for (int i = 0; i < 3; ++i) {
    GemmProblem g;
    g.A = As[i]; g.B = Bs[i]; g.M = T; g.N = 4 * H; g.K = ka; g.lda = ka; g.ldb = ka;
    g.a_dtype = DT_F32; g.b_dtype = DT_F32; g.epi = EPI_STORE_F32; g.out0 = gates32; g.ldo = 4 * H;
    g.bias = i == 0 ? bias4 : nullptr;
    g.aux = i == 0 ? nullptr : gates32;
    g.n_split = 2 * H; g.k_start_hi = H;
    RUN(gemm_launch(g, ST));
  }
  return mag_tail_fwd(gates32, text32, ln_gamma, ln_beta, beta_shift, ln_eps, p, Seed{seed, nullptr}, site, T, out32,
                      reinterpret_cast<__half*>(out16), ST);
}

int magb200_mag_bwd(const float* dout, const float* text32, const void* amag16, const float* gates32,
                    const void* packed16, const float* ln_gamma, float beta_shift, float ln_eps, float p, uint64_t seed,
                    uint32_t site, int V, int A, int ka, int T, float out_scale, void* dgates16, float* dpacked32,
                    float* dtext32, float* dW_hv, float* dW_ha, float* dW_v, float* dW_a, float* dbias4, float* dgamma, float* dbeta,
                    void* stream) {
  const int H = 768;
  MAGB_REQUIRE(ka == H + 128, "mag: ka must be 896");
  RUN(mag_tail_bwd(dout, gates32, text32, ln_gamma, beta_shift, ln_eps, p, Seed{seed, nullptr}, site, T, out_scale,
                   reinterpret_cast<__half*>(dgates16), dtext32, dgamma, dbeta, dbias4, ST));
  MAGB_CHECK_CUDA(cudaMemsetAsync(dpacked32, 0, (size_t)4 * H * ka * sizeof(float), ST));
  GemmProblem w;  // dB[4H][ka] += dG^T * Amag
  w.A = dgates16; w.B = amag16; w.M = 4 * H; w.N = ka; w.K = T; w.lda = 4 * H; w.ldb = ka; w.a_mn = true; w.b_mn = true;
  w.a_dtype = DT_F16; w.b_dtype = DT_F16; w.epi = EPI_ATOMIC_F32; w.out0 = dpacked32; w.ldo = ka; w.alpha = out_scale;
  RUN(gemm_launch(w, ST));
  RUN(mag_unpack_wgrad(dpacked32, V, A, ka, dW_hv, dW_ha, dW_v, dW_a, ST));
  GemmProblem d;  // dtext += dG[:, :2H] * B[:2H, :H]
  d.A = dgates16; d.B = packed16; d.M = T; d.N = H; d.K = 2 * H; d.lda = 4 * H; d.ldb = ka; d.b_mn = true;
  d.a_dtype = DT_F16; d.b_dtype = DT_F16; d.epi = EPI_STORE_F32; d.out0 = dtext32; d.aux = dtext32; d.ldo = H;
  return gemm_launch(d, ST);
}

int magb200_dropout_residual_ln_fwd(const float* pre, const float* res, const float* gamma, const float* beta,
                                    float eps, float p, uint64_t seed, uint32_t site, int T, float* y32, void* y16,
                                    float* rstd, void* stream) {
  return dropout_residual_ln_fwd(pre, res, gamma, beta, eps, p, Seed{seed, nullptr}, site, T, y32, reinterpret_cast<__half*>(y16), rstd,
                                 ST);
}

int magb200_linear_ln_fwd(const void* A16, int lda, const void* W16, int M, int K, const float* bias, const float* res,
                          const float* gamma, const float* beta, float eps, float p, uint64_t seed, uint32_t site, float* y32,
                          void* y16, float* rstd, void* stream) {
  GemmLnProblem q;
  q.A = A16; q.W = W16; q.M = M; q.K = K; q.lda = lda; q.bias = bias; q.res = res; q.gamma = gamma; q.beta = beta; q.eps = eps;
  q.y32 = y32; q.y16 = reinterpret_cast<__half*>(y16); q.rstd = rstd; q.drop_p = p; q.drop_seed = Seed{seed, nullptr};
  q.drop_site = site;
  return gemm_ln_launch(q, ST);
}

int magb200_ln_bwd(const float* dy32, const float* y32, const float* gamma, const float* beta, const float* rstd,
                   float p, uint64_t seed, uint32_t site, int T, float out_scale, float* dz32, void* dpre16,
                   float* dgamma, float* dbeta, float* dbias, void* stream) {
  return ln_bwd(dy32, y32, gamma, beta, rstd, p, Seed{seed, nullptr}, site, T, out_scale, dz32, reinterpret_cast<__half*>(dpre16), dgamma, dbeta,
                dbias, ST);
}

int magb200_colsum_f16(const void* in_f16, int T, int N, float out_scale, float* out, void* stream) {
  return colsum_f16(reinterpret_cast<const __half*>(in_f16), T, N, out_scale, out, ST);
}

int magb200_sdpa_fwd(const void* qkv16, const int64_t* mask, int B, int S, int heads, int head_dim, float p,
                     uint64_t seed, uint32_t site, void* ctx16, float* lse, void* stream) {
  return sdpa_fwd(reinterpret_cast<const __half*>(qkv16), mask, B, S, heads, head_dim, p, Seed{seed, nullptr}, site,
                  reinterpret_cast<__half*>(ctx16), lse, ST);
}

int magb200_sdpa_bwd(const void* qkv16, const void* dctx16, const void* ctx16, const float* lse, const int64_t* mask,
                     int B, int S, int heads, int head_dim, float p, uint64_t seed, uint32_t site, float out_scale,
                     void* dqkv16, float* dbias, void* workspace, size_t workspace_bytes, void* stream) {
  MAGB_REQUIRE(workspace == nullptr || workspace_bytes >= sdpa_flash_scratch_bytes(B, S, heads),
               "sdpa_bwd: workspace too small (%zu < %zu)", workspace_bytes, sdpa_flash_scratch_bytes(B, S, heads));
  return sdpa_bwd(reinterpret_cast<const __half*>(qkv16), reinterpret_cast<const __half*>(dctx16),
                  reinterpret_cast<const __half*>(ctx16), lse, mask, B, S, heads, head_dim, p, Seed{seed, nullptr}, site,
                  out_scale, reinterpret_cast<__half*>(dqkv16), dbias, reinterpret_cast<float*>(workspace), ST);
}

size_t magb200_sdpa_workspace_bytes(int B, int S, int heads) { return sdpa_flash_scratch_bytes(B, S, heads); }

int magb200_cls_fwd(const float* pre, const float* Wc, const float* bc, int B, int num_labels, float p, uint64_t seed,
                    uint32_t site, float* pooled, float* logits, void* stream) {
  return cls_fwd(pre, Wc, bc, B, num_labels, p, Seed{seed, nullptr}, site, pooled, logits, ST);
}

int magb200_cls_bwd(const float* dlogits, const float* pooled, const float* Wc, int B, int num_labels, float p,
                    uint64_t seed, uint32_t site, float out_scale, void* dpre16, float* dWc, float* dbc, void* stream) {
  return cls_bwd(dlogits, pooled, Wc, B, num_labels, p, Seed{seed, nullptr}, site, out_scale, reinterpret_cast<__half*>(dpre16), dWc, dbc, ST);
}

int magb200_mse_loss(const float* logits, const float* labels, int n, float grad_scale, float* loss, float* dlogits,
                     void* stream) {
  return mse_loss(logits, labels, n, grad_scale, loss, dlogits, ST);
}

int magb200_adamw_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, void* p16,
                       const uint8_t* decay_chunk, int64_t n, float lr, float beta1, float beta2, float eps,
                       float weight_decay, int step, float grad_unscale, int zero_grad, const float* hyper_dev,
                       void* stream) {
  return adamw_step(params, grads, exp_avg, exp_avg_sq, reinterpret_cast<__half*>(p16), decay_chunk, n, lr, beta1, beta2,
                    eps, weight_decay, step, grad_unscale, zero_grad, hyper_dev, ST);
}

int magb200_peer_barrier(void* const* flag_ptrs, int world, int rank, void* epoch_counter, void* stream) {
  return peer_barrier(reinterpret_cast<unsigned int* const*>(flag_ptrs), world, rank, reinterpret_cast<unsigned int*>(epoch_counter), ST);
}

int magb200_adamw_peer_step(const void* const* grad_ptrs, void* const* param_ptrs, int world, int rank, float* exp_avg,
                            float* exp_avg_sq, void* p16, const uint8_t* decay_chunk, int64_t slice_begin, int64_t slice_end,
                            int64_t clean_begin, int64_t clean_end, float lr, float beta1, float beta2, float eps,
                            float weight_decay, int step, float grad_unscale, const float* hyper_dev, int max_blocks,
                            const uint8_t* row_any, int row_mode, int64_t table_begin, int table_rows, int row_len,
                            int clean_what, int64_t keep_begin, int64_t keep_end, const void* multicast_grad,
                            void* multicast_param, void* stream) {
  return adamw_peer_step(reinterpret_cast<const float* const*>(grad_ptrs), reinterpret_cast<float* const*>(param_ptrs), world,
                         rank, exp_avg, exp_avg_sq, reinterpret_cast<__half*>(p16), decay_chunk, slice_begin, slice_end, clean_begin,
                         clean_end, lr, beta1, beta2, eps, weight_decay, step, grad_unscale, hyper_dev, max_blocks, row_any, row_mode,
                         table_begin, table_rows, row_len, clean_what, keep_begin, keep_end,
                         reinterpret_cast<const float*>(multicast_grad), reinterpret_cast<float*>(multicast_param), ST);
}

int magb200_peer_mark_rows(const int64_t* ids, int n, int vocab, void* marks, void* counter, void* stream) {
  return peer_mark_rows(ids, n, vocab, reinterpret_cast<unsigned int*>(marks), reinterpret_cast<unsigned int*>(counter), ST);
}

int magb200_peer_combine_marks(const void* const* mark_ptrs, int world, const void* counter, int vocab, uint8_t* row_any,
                               void* stream) {
  return peer_combine_marks(reinterpret_cast<const unsigned int* const*>(mark_ptrs), world,
                            reinterpret_cast<const unsigned int*>(counter), vocab, row_any, ST);
}

size_t magb200_bert_workspace_bytes(const magb200_bert_dims* dims) { return bert_workspace_bytes(*dims); }

int magb200_bert_refresh_operands(const magb200_bert_dims* dims, const int64_t* off, const float* params,
                                  void* params16, void* stream) {
  return bert_refresh_operands(*dims, off, params, reinterpret_cast<__half*>(params16), ST);
}

int magb200_bert_forward(const magb200_bert_dims* dims, const int64_t* off, const float* params, const void* params16,
                         const int64_t* input_ids, const int64_t* token_type_ids, const int64_t* attention_mask,
                         const float* visual, const float* acoustic, void* workspace, size_t workspace_bytes,
                         float* logits, void* stream) {
  return bert_forward(*dims, off, params, reinterpret_cast<const __half*>(params16), input_ids, token_type_ids,
                      attention_mask, visual, acoustic, workspace, workspace_bytes, logits, ST);
}

int magb200_bert_backward(const magb200_bert_dims* dims, const int64_t* off, const float* params, const void* params16,
                          const int64_t* input_ids, const int64_t* token_type_ids, const int64_t* attention_mask,
                          void* workspace, size_t workspace_bytes, const float* dlogits_scaled, float* grads,
                          void* stream) {
  return bert_backward(*dims, off, params, reinterpret_cast<const __half*>(params16), input_ids, token_type_ids,
                       attention_mask, workspace, workspace_bytes, dlogits_scaled, grads, ST);
}

int magb200_bert_backward_part(const magb200_bert_dims* dims, const int64_t* off, const float* params,
                               const void* params16, const int64_t* input_ids, const int64_t* token_type_ids,
                               const int64_t* attention_mask, void* workspace, size_t workspace_bytes,
                               const float* dlogits_scaled, float* grads, int layer_hi, int layer_lo, int flags,
                               void* stream) {
  return bert_backward_part(*dims, off, params, reinterpret_cast<const __half*>(params16), input_ids, token_type_ids,
                            attention_mask, workspace, workspace_bytes, dlogits_scaled, grads, layer_hi, layer_lo, flags, ST);
}

int magb200_bert_backward_adamw(const magb200_bert_dims* dims, const int64_t* off, float* params, void* params16,
                                const int64_t* input_ids, const int64_t* token_type_ids, const int64_t* attention_mask,
                                void* workspace, size_t workspace_bytes, const float* dlogits_scaled, float* grads,
                                float* exp_avg, float* exp_avg_sq, const uint8_t* decay_chunk, float lr, float beta1,
                                float beta2, float eps, float weight_decay, int step, float grad_unscale, int zero_grad,
                                const float* hyper_dev, void* stream) {
  OptHook h;
  h.params = params; h.params16 = reinterpret_cast<__half*>(params16); h.exp_avg = exp_avg; h.exp_avg_sq = exp_avg_sq;
  h.decay_chunk = decay_chunk; h.lr = lr; h.beta1 = beta1; h.beta2 = beta2; h.eps = eps; h.weight_decay = weight_decay;
  h.step = step; h.grad_unscale = grad_unscale; h.zero_grad = zero_grad; h.hyper_dev = hyper_dev;
  MAGB_REQUIRE(params != nullptr && grads != nullptr && exp_avg != nullptr && exp_avg_sq != nullptr &&
                   decay_chunk != nullptr, "magb200_bert_backward_adamw: null optimizer buffer");
  return bert_backward_part(*dims, off, params, reinterpret_cast<const __half*>(params16), input_ids, token_type_ids,
                            attention_mask, workspace, workspace_bytes, dlogits_scaled, grads, dims->layers - 1, 0, 3, ST,
                            &h);
}

int magb200_bert_attention_probs(const magb200_bert_dims* dims, void* workspace, int layer, const int64_t* attention_mask,
                                 float* out, size_t out_bytes, void* stream) {
  return bert_attention_probs(*dims, workspace, layer, attention_mask, out, out_bytes, ST);
}

int magb200_bert_peek(const magb200_bert_dims* dims, void* workspace, int what, int layer, void* out, size_t out_bytes,
                      void* stream) {
  return bert_peek(*dims, workspace, what, layer, out, out_bytes, ST);
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ XLNet
extern "C" {

size_t magb200_xlnet_workspace_bytes(const magb200_xlnet_dims* dims) { return xlnet_workspace_bytes(*dims); }

int magb200_xlnet_refresh_operands(const magb200_xlnet_dims* dims, const int64_t* off, const float* params,
                                   void* params16, void* stream) {
  return xlnet_refresh_operands(*dims, off, params, reinterpret_cast<__half*>(params16), ST);
}

int magb200_xlnet_forward(const magb200_xlnet_dims* dims, const int64_t* off, const float* params, const void* params16,
                          const int64_t* input_ids, const int64_t* token_type_ids, const int64_t* attention_mask,
                          const float* visual, const float* acoustic, void* workspace, size_t workspace_bytes,
                          float* logits, void* stream) {
  return xlnet_forward(*dims, off, params, reinterpret_cast<const __half*>(params16), input_ids, token_type_ids,
                       attention_mask, visual, acoustic, workspace, workspace_bytes, logits, ST);
}

int magb200_xlnet_backward(const magb200_xlnet_dims* dims, const int64_t* off, const float* params, const void* params16,
                           const int64_t* input_ids, const int64_t* token_type_ids, const int64_t* attention_mask,
                           void* workspace, size_t workspace_bytes, const float* dlogits_scaled, float* grads,
                           void* stream) {
  return xlnet_backward(*dims, off, params, reinterpret_cast<const __half*>(params16), input_ids, token_type_ids,
                        attention_mask, workspace, workspace_bytes, dlogits_scaled, grads, ST);
}

int magb200_xlnet_peek(const magb200_xlnet_dims* dims, void* workspace, int what, int layer, void* out,
                       size_t out_bytes, void* stream) {
  return xlnet_peek(*dims, workspace, what, layer, out, out_bytes, ST);
}

static RelAttnArgs make_rel(const void* qkv16, const void* kr16, int Bp, const float* r_w_bias, const float* r_r_bias,
                            const float* r_s_bias, const float* seg_embed, const int64_t* mask, const int64_t* tt, int B,
                            int S, int heads, int head_dim, float p, uint64_t seed, uint32_t site) {
  RelAttnArgs a;
  a.qkv = reinterpret_cast<const __half*>(qkv16); a.kr16 = reinterpret_cast<const __half*>(kr16); a.Bp = Bp;
  a.r_w_bias = r_w_bias; a.r_r_bias = r_r_bias; a.r_s_bias = r_s_bias; a.seg_embed = seg_embed; a.mask = mask; a.tt = tt;
  a.B = B; a.S = S; a.heads = heads; a.head_dim = head_dim; a.p = p; a.seed = Seed{seed, nullptr}; a.site = site;
  return a;
}

int magb200_xlnet_rel_attn_fwd(const void* qkv16, const void* kr16, int Bp, const float* r_w_bias, const float* r_r_bias,
                               const float* r_s_bias, const float* seg_embed, const int64_t* attention_mask,
                               const int64_t* token_type_ids, int B, int S, int heads, int head_dim, float p, uint64_t seed,
                               uint32_t site, void* ctx16, void* stream) {
  return xl_rel_attn_fwd(make_rel(qkv16, kr16, Bp, r_w_bias, r_r_bias, r_s_bias, seg_embed, attention_mask, token_type_ids,
                                  B, S, heads, head_dim, p, seed, site),
                         reinterpret_cast<__half*>(ctx16), ST);
}

int magb200_xlnet_rel_attn_bwd(const void* qkv16, const void* kr16, int Bp, const float* r_w_bias, const float* r_r_bias,
                               const float* r_s_bias, const float* seg_embed, const int64_t* attention_mask,
                               const int64_t* token_type_ids, const void* dctx16, int B, int S, int heads, int head_dim,
                               float p, uint64_t seed, uint32_t site, float out_scale, void* dqkv16, float* dkr32,
                               float* d_r_w_bias, float* d_r_r_bias, float* d_r_s_bias, float* d_seg_embed, void* stream) {
  return xl_rel_attn_bwd(make_rel(qkv16, kr16, Bp, r_w_bias, r_r_bias, r_s_bias, seg_embed, attention_mask, token_type_ids,
                                  B, S, heads, head_dim, p, seed, site),
                         reinterpret_cast<const __half*>(dctx16), out_scale, reinterpret_cast<__half*>(dqkv16), dkr32,
                         d_r_w_bias, d_r_r_bias, d_r_s_bias, d_seg_embed, ST);
}

int magb200_xlnet_pos_emb(int S, int Bp, float p, uint64_t seed, uint32_t site, void* pos16, void* stream) {
  return xl_pos_emb(S, Bp, p, Seed{seed, nullptr}, site, reinterpret_cast<__half*>(pos16), ST);
}

}  // extern "C"

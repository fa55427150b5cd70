// Internal host-side declarations shared by the translation units of libmagb200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

namespace magb200 {

// Dropout seed of a step: by value, or read from device memory at kernel run time (ptr != null) so a
// captured CUDA graph can be replayed with a fresh seed per step.
struct Seed {
  uint64_t value = 0;
  const uint64_t* ptr = nullptr;
};

struct GemmProblem {
  const void* A = nullptr;   // K-major: [M][K] pitch lda ; MN-major: [K][M] pitch lda
  const void* B = nullptr;   // K-major: [N][K] pitch ldb ; MN-major: [K][N] pitch ldb
  int M = 0, N = 0, K = 0;
  int lda = 0, ldb = 0;      // in elements
  int a_dtype = 1, b_dtype = 1;  // DType
  bool a_mn = false, b_mn = false;
  int splits = 1;            // split-K (atomic epilogue only)
  int n_split = 0;           // >0: output tiles with n0 >= n_split skip K < k_start_hi
  int k_start_hi = 0;
  int epi = 0;               // GemmEpilogue
  const float* bias = nullptr;
  void* out0 = nullptr;
  void* out1 = nullptr;
  const void* aux = nullptr;
  int ldo = 0;
  float alpha = 1.0f;
  // dropout fused into the GELU / GELU' epilogues (p = 0: off); element index = m * N + n
  float drop_p = 0.f;
  Seed drop_seed;
  uint32_t drop_site = 0;
  int role = 0;              // scheduling hint: 0 = forward Linear, 1 = dgrad, 2 = wgrad (kernel selection policy)
};

int gemm_launch(const GemmProblem& p, cudaStream_t stream);

// y = LayerNorm(dropout(A W^T + bias) + res), hidden size 768, as one kernel (gemm_ln_sm100.cuh): A [M][K] fp16, W [768][K]
// fp16 (both K-major), res / y32 [M][768] fp32, y16 (nullable) fp16 copy, rstd (nullable) [M] saved for the backward pass
struct GemmLnProblem {
  const void* A = nullptr; const void* W = nullptr;
  int M = 0, K = 0, lda = 0;
  const float* bias = nullptr; const float* res = nullptr; const float* gamma = nullptr; const float* beta = nullptr;
  float eps = 1e-12f;
  float* y32 = nullptr; __half* y16 = nullptr; float* rstd = nullptr;
  float drop_p = 0.f; Seed drop_seed; uint32_t drop_site = 0;
};
int gemm_ln_launch(const GemmLnProblem& p, cudaStream_t stream);
void gemm_set_debug(unsigned long long* trace, int cfg);

// ---------------------------------------------------------------- rowops.cu
struct EmbedArgs {
  const int64_t* ids = nullptr;      // [T]
  const int64_t* tt = nullptr;       // [T] or null (all zero)
  const float* visual = nullptr;     // [T][V]
  const float* acoustic = nullptr;   // [T][A]
  const float* word = nullptr; const float* pos = nullptr; const float* type = nullptr;
  const float* gamma = nullptr; const float* beta = nullptr;
  float eps = 1e-12f;
  float p = 0.f; Seed seed; uint32_t site = 0;
  int V = 0, A = 0, KA = 896, T = 0, S = 0, vocab = 0, hidden = 768;
  int type_vocab = 2;                // rows of the token-type table (backward: type >= type_vocab is clamped)
  float* emb32 = nullptr;            // [T][768]
  __half* amag = nullptr;            // [T][KA] or null
};
int embed_ln_fwd(const EmbedArgs& a, cudaStream_t st);
int embed_ln_bwd(const EmbedArgs& a, const float* dout, float oscale, float* dword, float* dpos, float* dtype,
                 float* dgamma, float* dbeta, cudaStream_t st);
// mag_fused.cu (K1): embeddings + MAG as one warp-specialised cooperative kernel.  e.emb32 / e.amag are outputs (saved for
// backward); gates32 [T][3072] may be null (eval); scratch: mag_fused_scratch_bytes(T) bytes
size_t mag_fused_scratch_bytes(int T);
int mag_fused_fwd(const EmbedArgs& e, const __half* packed16, const float* bias4, const float* mgamma, const float* mbeta,
                  float beta_shift, float m_eps, float p_mag, uint32_t site_mag, float* gates32, void* scratch,
                  float* out32, __half* out16, cudaStream_t st);
int mag_pack_inputs(const float* text, const float* visual, const float* acoustic, int V, int A, int KA, int T,
                    __half* amag, cudaStream_t st);
int mag_pack_split(const float* text, const float* visual, const float* acoustic, const float* W_hv, const float* W_ha,
                   const float* W_v, const float* W_a, int V, int A, int KA, int T, float* a_hi, float* a_lo, float* b_hi,
                   float* b_lo, cudaStream_t st);
int mag_tail_fwd(const float* G, const float* text32, const float* gamma, const float* beta, float beta_shift, float eps,
                 float p, Seed seed, uint32_t site, int T, float* out32, __half* out16, cudaStream_t st);
int mag_tail_bwd(const float* dout, const float* G, const float* text32, const float* gamma, float beta_shift, float eps,
                 float p, Seed seed, uint32_t site, int T, float oscale, __half* dG, float* dtext, float* dgamma,
                 float* dbeta, float* dbias4, cudaStream_t st);
// zero_after (nullable, normally == pre): rows are rewritten with zeros once read -- `pre` is then a reduction target
// that the next stream-K GEMM can red.add into without a memset
int dropout_residual_ln_fwd(const float* pre, const float* res, const float* gamma, const float* beta, float eps, float p,
                            Seed seed, uint32_t site, int T, float* y32, __half* y16, float* rstd, cudaStream_t st,
                            float* zero_after = nullptr);
int ln_bwd(const float* dy32, const float* y32, const float* gamma, const float* beta, const float* rstd, float p,
           Seed seed, uint32_t site, int T, float oscale, float* dz32, __half* dpre, float* dgamma, float* dbeta,
           float* dbias, cudaStream_t st);
int colsum_f16(const __half* in, int T, int N, float oscale, float* out, cudaStream_t st);
int cast_f32_f16(const float* in, __half* out, int64_t n, cudaStream_t st);
int mag_pack_weights(const float* W_hv, const float* W_ha, const float* W_v, const float* W_a, int V, int A, int KA,
                     __half* Bm, cudaStream_t st);
int mag_unpack_wgrad(const float* dB, int V, int A, int KA, float* dW_hv, float* dW_ha, float* dW_v, float* dW_a,
                     cudaStream_t st);
int dropout_mask(uint8_t* out, int64_t n, float p, Seed seed, uint32_t site, cudaStream_t st);


// ---------------------------------------------------------------- attention.cu
// lse (fwd out / bwd in, [B][heads][S] fp32) and ctx select the warp-MMA kernels for S <= 64; dbias (nullable)
// receives oscale * column sums of dqkv (the fused q|k|v bias gradient)
int sdpa_fwd(const __half* qkv, const int64_t* mask, int B, int S, int heads, int head_dim, float p, Seed seed,
             uint32_t site, __half* ctx, float* lse, cudaStream_t st);
// scratch (nullable): sdpa_flash_scratch_bytes(B, S, heads) bytes, needed by the S > 64 backward (rowsum(dO * O))
int sdpa_bwd(const __half* qkv, const __half* dctx, const __half* ctx, const float* lse, const int64_t* mask, int B,
             int S, int heads, int head_dim, float p, Seed seed, uint32_t site, float oscale, __half* dqkv, float* dbias,
             float* scratch, cudaStream_t st);
// attention_flash.cu: non-materialising kernels for any S (used for S > 64)
size_t sdpa_flash_scratch_bytes(int B, int S, int heads);
int sdpa_fwd_flash(const __half* qkv, const int64_t* mask, int B, int S, int heads, float p, Seed seed, uint32_t site,
                   __half* ctx, float* lse, cudaStream_t st);
int sdpa_bwd_flash(const __half* qkv, const __half* dctx, const __half* ctx, const float* lse, const int64_t* mask, int B,
                   int S, int heads, float p, Seed seed, uint32_t site, float* Dq_scratch, __half* dqkv, cudaStream_t st);
int sdpa_probs(const __half* qkv, const int64_t* mask, int B, int S, int heads, float p, Seed seed, uint32_t site,
               float* probs, cudaStream_t st);
int sdpa_fwd_mma(const __half* qkv, const int64_t* mask, int B, int S, int heads, float p, Seed seed, uint32_t site,
                 __half* ctx, float* lse, cudaStream_t st);
int sdpa_bwd_mma(const __half* qkv, const __half* dctx, const __half* ctx, const float* lse, const int64_t* mask, int B,
                 int S, int heads, float p, Seed seed, uint32_t site, float oscale, __half* dqkv, float* dbias,
                 cudaStream_t st);

// ---------------------------------------------------------------- head_optim.cu
int cls_fwd(const float* pre, const float* Wc, const float* bc, int B, int L, float p, Seed seed, uint32_t site,
            float* pooled, float* logits, cudaStream_t st);
int cls_bwd(const float* dlogits, const float* pooled, const float* Wc, int B, int L, float p, Seed seed,
            uint32_t site, float oscale, __half* dpre, float* dWc, float* dbc, cudaStream_t st);
int mse_loss(const float* logits, const float* labels, int n, float gscale, float* loss, float* dlogits, cudaStream_t st);
int adamw_step(float* p, float* g, float* m, float* v, __half* p16, const uint8_t* decay_chunk, int64_t n, float lr,
               float b1, float b2, float eps, float wd, int step, float gscale, int zero_grad, const float* hyper_dev,
               cudaStream_t st);


// ---------------------------------------------------------------- peer_optim.cu
// data-parallel step over NVLink peer memory: flag barrier between the ranks, and reduce-scatter + AdamW + all-gather as one
// kernel (grad_ptrs / param_ptrs: the flat fp32 buffers of every rank, this rank included; m, v, p16: local, indexed like
// the flat buffer; only elements [slice_begin, slice_end) are touched)
int peer_barrier(unsigned int* const* flag_ptrs, int world, int rank, unsigned int* epoch_counter, cudaStream_t st);
int adamw_peer_step(const float* const* grad_ptrs, float* const* param_ptrs, int world, int rank, float* m, float* v,
                    __half* p16, const uint8_t* decay_chunk, int64_t slice_begin, int64_t slice_end, int64_t clean_begin,
                    int64_t clean_end, float lr, float b1, float b2, float eps, float wd, int step, float gscale,
                    const float* hyper_dev, int max_blocks, const uint8_t* row_any, int row_mode, int64_t table_begin,
                    int table_rows, int row_len, int clean_what, int64_t keep_begin, int64_t keep_end,
                    const float* mc_grad, float* mc_param, cudaStream_t st);
int peer_mark_rows(const int64_t* ids, int n, int vocab, unsigned int* marks, unsigned int* counter, cudaStream_t st);
int peer_combine_marks(const unsigned int* const* mark_ptrs, int world, const unsigned int* counter, int vocab, uint8_t* any,
                       cudaStream_t st);

// ---------------------------------------------------------------- xlnet_ops.cu
struct RelAttnArgs {
  const __half* qkv = nullptr;      // [B*S][3*heads*64]  (q | k | v head projections, no bias)
  const __half* kr16 = nullptr;     // [Bp][2S][heads*64] projected relative positions (Bp = 1 or B)
  int Bp = 1;
  const float* r_w_bias = nullptr; const float* r_r_bias = nullptr; const float* r_s_bias = nullptr;   // [heads][64]
  const float* seg_embed = nullptr; // [2][heads][64]
  const int64_t* mask = nullptr;    // [B][S] attention_mask (1 = token, 0 = pad) or null
  const int64_t* tt = nullptr;      // [B][S] segment ids or null
  int B = 0, S = 0, heads = 12, head_dim = 64;
  float p = 0.f; Seed seed; uint32_t site = 0;
};
int xl_gather_fwd(const int64_t* ids, const float* word, int T, int vocab, float p, Seed seed, uint32_t site, float* x32,
                  __half* x16, cudaStream_t st);
int xl_gather_bwd(const float* dx, const int64_t* ids, int T, int vocab, float p, Seed seed, uint32_t site, float oscale,
                  float* dword, cudaStream_t st);
int xl_pos_emb(int S, int Bp, float p, Seed seed, uint32_t site, __half* pos16, cudaStream_t st);
int xl_last_token_fwd(const float* x32, int B, int S, float p, Seed seed, uint32_t site, __half* out16, cudaStream_t st);
int xl_last_token_bwd(const float* dlast, int B, int S, float p, Seed seed, uint32_t site, float* dx32, cudaStream_t st);
int xl_pack_qkv(const float* q, const float* k, const float* v, __half* out, cudaStream_t st);
int xl_rel_attn_fwd(const RelAttnArgs& a, __half* ctx, cudaStream_t st);
int xl_rel_attn_bwd(const RelAttnArgs& a, const __half* dctx, float oscale, __half* dqkv, float* dkr32, float* d_rw,
                    float* d_rr, float* d_rs, float* d_seg, cudaStream_t st);

}  // namespace magb200

// ---------------------------------------------------------------- model_bert.cu
struct magb200_bert_dims;
namespace magb200 {
size_t bert_workspace_bytes(const magb200_bert_dims& d);
int bert_refresh_operands(const magb200_bert_dims& d, const int64_t* off, const float* params, __half* params16,
                          cudaStream_t st);
int bert_forward(const magb200_bert_dims& d, const int64_t* off, const float* P, const __half* P16, const int64_t* ids,
                 const int64_t* tt, const int64_t* mask, const float* visual, const float* acoustic, void* workspace,
                 size_t ws_bytes, float* logits, cudaStream_t st);
int bert_backward(const magb200_bert_dims& d, const int64_t* off, const float* P, const __half* P16, const int64_t* ids,
                  const int64_t* tt, const int64_t* mask, void* workspace, size_t ws_bytes, const float* dlogits_scaled,
                  float* G, cudaStream_t st);
// AdamW fused into the backward schedule (model_bert.cu): slices of the flat buffers are updated while the
// backward pass of the layers below is still running.  Field meaning as magb200_adamw_step.
struct OptHook {
  float* params = nullptr; __half* params16 = nullptr; float* exp_avg = nullptr; float* exp_avg_sq = nullptr;
  const uint8_t* decay_chunk = nullptr;
  float lr = 0.f, beta1 = 0.9f, beta2 = 0.999f, eps = 1e-6f, weight_decay = 0.f;
  int step = 1; float grad_unscale = 1.f; int zero_grad = 1; const float* hyper_dev = nullptr;
};
int bert_backward_part(const magb200_bert_dims& d, const int64_t* off, const float* P, const __half* P16,
                       const int64_t* ids, const int64_t* tt, const int64_t* mask, void* workspace, size_t ws_bytes,
                       const float* dlogits_scaled, float* G, int layer_hi, int layer_lo, int flags, cudaStream_t st,
                       const OptHook* opt = nullptr);
int bert_peek(const magb200_bert_dims& d, void* workspace, int what, int layer, void* out, size_t out_bytes,
              cudaStream_t st);
// attention probabilities of `layer` ([B][heads][S][S] fp32, after dropout) recomputed from the workspace of a forward
int bert_attention_probs(const magb200_bert_dims& d, void* workspace, int layer, const int64_t* mask, float* out,
                         size_t out_bytes, cudaStream_t st);
}  // namespace magb200

// ---------------------------------------------------------------- model_xlnet.cu
struct magb200_xlnet_dims;
namespace magb200 {
size_t xlnet_workspace_bytes(const magb200_xlnet_dims& d);
int xlnet_refresh_operands(const magb200_xlnet_dims& d, const int64_t* off, const float* params, __half* params16,
                           cudaStream_t st);
int xlnet_forward(const magb200_xlnet_dims& d, const int64_t* off, const float* P, const __half* P16, const int64_t* ids,
                  const int64_t* tt, const int64_t* mask, const float* visual, const float* acoustic, void* workspace,
                  size_t ws_bytes, float* logits, cudaStream_t st);
int xlnet_backward(const magb200_xlnet_dims& d, const int64_t* off, const float* P, const __half* P16, const int64_t* ids,
                   const int64_t* tt, const int64_t* mask, void* workspace, size_t ws_bytes, const float* dlogits_scaled,
                   float* G, cudaStream_t st);
int xlnet_peek(const magb200_xlnet_dims& d, void* workspace, int what, int layer, void* out, size_t out_bytes,
               cudaStream_t st);
}  // namespace magb200


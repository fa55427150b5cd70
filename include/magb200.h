/* libmagb200 -- C ABI of the B200-native MAG-BERT / MAG-XLNet hot path.
 *
 * The reference (WasifurRahman/BERT_multimodal_transformer) has no FFI of its own: its hot path
 * is Python calling torch ops.  This header is the boundary a maintainer would bind from the
 * reference's modules (ctypes stub in INTEGRATION.md); each entry point cites the reference
 * lines whose arithmetic it replaces.
 *
 * Conventions: every pointer is a DEVICE pointer owned by the caller (16-byte aligned, row-major,
 * contiguous unless a pitch is given); nothing is allocated or freed by the library; work is
 * enqueued on `stream` (a cudaStream_t) and the call returns without synchronising.  Return
 * value 0 = success, otherwise a cudaError_t-style code or -1 for argument errors, with the
 * message available from magb200_last_error() (thread-local).
 */
#ifndef MAGB200_H_
#define MAGB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAGB200_VERSION 100

/* element types */
#define MAGB200_F32 0  /* fp32 in memory; GEMM operands are read as tf32 */
#define MAGB200_F16 1
#define MAGB200_BF16 2

/* GEMM epilogues: v = alpha * acc + bias[n] */
#define MAGB200_EPI_STORE_F32 0      /* out0(f32) = v (+ aux(f32) when aux != NULL) */
#define MAGB200_EPI_STORE_F16 1      /* out0(f16) = v */
#define MAGB200_EPI_STORE_BF16 2     /* out0(bf16) = v */
#define MAGB200_EPI_GELU_F16 3       /* out0(f16) = gelu_erf(v); out1(f16) = gelu_erf'(v) when out1 != NULL (saved for backward) */
#define MAGB200_EPI_DGELU_F16 4      /* out0(f16) = v * aux(f16), aux = the derivative saved by MAGB200_EPI_GELU_F16 */
#define MAGB200_EPI_ATOMIC_F32 5     /* out0(f32) += v */
#define MAGB200_EPI_STORE_F32_F16 6  /* out0(f32) = v; out1(f16) = v */
#define MAGB200_EPI_RELU_F32 7       /* out0(f32) = max(v, 0) */

const char* magb200_last_error(void);
int magb200_version(void);
/* number of kernels this library has launched (or recorded into a capturing stream) in this process */
unsigned long long magb200_launch_count(void);
/* 0 iff the current device is an sm_100 part */
int magb200_device_check(void);

/* tuning/debug hook: trace_dev (device, 64 x uint64 per CTA, or NULL) receives clock64 stamps of the GEMM
 * pipeline; cfg selects the tile configuration: 0 = default (128x128, 3 stages, 2-CTA clusters multicasting the A
 * tile whenever the number of N tiles is even), 5 = the same without clusters, 1 = 128x128 4 stages,
 * 3 = 128x256x3, 4 = 128x256x4 (no clusters). */
void magb200_gemm_set_debug(void* trace_dev, int cfg);

/* D[M,N] = A[M,K] * B[N,K]^T on tcgen05 tensor cores (TMA-fed, TMEM accumulators).
 * Replaces every nn.Linear / einsum contraction on the reference path (F.linear at
 * modeling.py:27-30; BertSelfAttention/BertSelfOutput/BertIntermediate/BertOutput Linears
 * called through bert.py:221-229; einsum projections called through xlnet.py:374-385) and
 * their autograd dgrad/wgrad counterparts (multimodal_driver.py:378).
 *   a_mn_major = 0: A stored [M][K] (pitch lda);  1: A stored [K][M] (pitch lda)
 *   b_mn_major = 0: B stored [N][K] (pitch ldb);  1: B stored [K][N] (pitch ldb)
 *   splits > 1 (ATOMIC epilogue only): split-K over blockIdx.z
 *   n_split > 0: output columns >= n_split only accumulate K >= k_start_hi (block-sparse B). */
int magb200_gemm(const void* A, const void* B, int M, int N, int K, int lda, int ldb, int a_dtype, int b_dtype,
                 int a_mn_major, int b_mn_major, int epilogue, const float* bias, void* out0, void* out1,
                 const void* aux, int ldo, float alpha, int splits, int n_split, int k_start_hi, void* stream);


/* ------------------------------------------------------------------------------------------------
 * Row-wise operators (fp32 activations, hidden size 768).  Dropout uses a counter-based RNG:
 * the keep-mask of a site is a pure function of (seed, site, element index), so backward
 * regenerates it.  p = 0 disables dropout (eval mode). */

/* keep-mask bytes (1 = keep) the kernels use for (seed, site) over n elements -- test hook */
int magb200_dropout_mask(uint8_t* out, int64_t n, float p, uint64_t seed, uint32_t site, void* stream);

/* fp32 -> fp16 copy (GEMM operand refresh) */
int magb200_cast_f32_f16(const float* in, void* out_f16, int64_t n, void* stream);

/* BertEmbeddings.forward (transformers 3.0.2, called at bert.py:211-216):
 *   emb32[t] = dropout(LayerNorm(word[ids[t]] + pos[t % S] + type[tt[t]]))
 * and, when amag16 != NULL, the packed MAG GEMM operand row
 *   amag16[t] = [ fp16(emb32[t]) | fp16(visual[t]) | fp16(acoustic[t]) | 0.. ]   (pitch ka = 896) */
int magb200_embed_ln_fwd(const int64_t* ids, const int64_t* token_type_ids, const float* visual, const float* acoustic,
                         const float* word, const float* pos, const float* type, const float* gamma, const float* beta,
                         float eps, float p, uint64_t seed, uint32_t site, int V, int A, int ka, int T, int S, int vocab,
                         float* emb32, void* amag16, void* stream);
int magb200_embed_ln_bwd(const float* dout, const int64_t* ids, const int64_t* token_type_ids, const float* word,
                         const float* pos, const float* type, const float* gamma, float eps, float p, uint64_t seed,
                         uint32_t site, int T, int S, int vocab, float out_scale, float* dword, float* dpos, float* dtype,
                         float* dgamma, float* dbeta, void* stream);

/* K1 -- BertEmbeddings (bert.py:211-216) and MAG.forward (modeling.py:25-51, called at bert.py:219) as ONE
 * warp-specialised cooperative kernel: embedding gather + LayerNorm(emb_eps) + dropout(p_emb), the four gate Linears
 * on tcgen05 (operand A = [h | visual | acoustic] packed in amag16, B = packed16 from magb200_mag_pack_weights), ReLU
 * gates, h_m, alpha = min(beta_shift |h| / (|h_m| + 1e-6), 1), LayerNorm(mag_eps)(alpha h_m + h), dropout(p_mag).
 * Outputs: out32 / out16 (optional) [T][768]; emb32 [T][768] and amag16 [T][ka] (both read by the backward pass);
 * gates32 [T][3072] = the gate pre-activations for magb200_mag_bwd, or NULL to skip them (inference).
 * scratch: magb200_embed_mag_scratch_bytes(T) bytes.  Same results as magb200_embed_ln_fwd -> magb200_mag_fwd. */
size_t magb200_embed_mag_scratch_bytes(int T);
int magb200_embed_mag_fwd(const int64_t* ids, const int64_t* token_type_ids, const float* visual, const float* acoustic,
                          const float* word, const float* pos, const float* type, const float* emb_gamma,
                          const float* emb_beta, float emb_eps, float p_emb, uint64_t seed, uint32_t site_emb, int V, int A,
                          int ka, int T, int S, int vocab, const void* packed16, const float* bias4, const float* mag_gamma,
                          const float* mag_beta, float beta_shift, float mag_eps, float p_mag, uint32_t site_mag,
                          float* emb32, void* amag16, float* gates32, void* scratch, size_t scratch_bytes, float* out32,
                          void* out16, void* stream);

/* MAG.forward (modeling.py:25-51), all of it:
 *   gates GEMM on tcgen05 (the four Linears W_hv, W_ha, W_v, W_a as one block matrix) followed by the
 *   fused tail (ReLU gates, h_m, the two L2 norms, alpha = min(beta*|text|/(|h_m|+1e-6), 1) with the
 *   hm_norm == 0 guard, shift, LayerNorm(eps), dropout).
 * text32 [T][768], visual [T][V], acoustic [T][A]; packed16 = magb200_mag_pack_weights output;
 * bias4 = [b_hv | b_ha | b_v | b_a]; scratch: amag16 [T][ka] fp16, gates32 [T][3072] fp32.
 * out32 [T][768]; out16 optional fp16 copy. */
int magb200_mag_pack_weights(const float* W_hv, const float* W_ha, const float* W_v, const float* W_a, int V, int A,
                             int ka, void* packed16, void* stream);
int magb200_mag_fwd(const float* text32, const float* visual, const float* acoustic, const void* packed16,
                    const float* bias4, const float* ln_gamma, const float* ln_beta, float beta_shift, float ln_eps,
                    float p, uint64_t seed, uint32_t site, int V, int A, int ka, int T, void* amag16, float* gates32,
                    float* out32, void* out16, void* stream);
/* The same forward with the four Linears evaluated to fp32 accuracy on the tensor cores: operands split into TF32
 * hi + lo parts, three kind::tf32 GEMMs (hi*hi + hi*lo + lo*hi) -- for BASELINE.json configs[0]'s 1e-5 tolerance on
 * MAG.forward (modeling.py:25-51).  Takes the four fp32 weight matrices directly; split_scratch holds
 * 2 * (T + 4*768) * ka floats.  gates32 is compatible with magb200_mag_bwd (which stays fp16). */
int magb200_mag_fwd_exact(const float* text32, const float* visual, const float* acoustic, const float* W_hv,
                          const float* W_ha, const float* W_v, const float* W_a, const float* bias4,
                          const float* ln_gamma, const float* ln_beta, float beta_shift, float ln_eps, float p,
                          uint64_t seed, uint32_t site, int V, int A, int ka, int T, float* split_scratch,
                          float* gates32, float* out32, void* out16, void* stream);
/* Backward of magb200_mag_fwd from saved amag16 / gates32.  dout is the (loss-scaled) gradient of out32.
 * Writes dtext32 [T][768] (still scaled); accumulates (+=) out_scale * {dW_hv, dW_ha, dW_v, dW_a, dbias4, dgamma, dbeta}.
 * scratch: dgates16 [T][3072] fp16, dpacked32 [3072][ka] fp32. */
int magb200_mag_bwd(const float* dout, const float* text32, const void* amag16, const float* gates32,
                    const void* packed16, const float* ln_gamma, float beta_shift, float ln_eps, float p, uint64_t seed,
                    uint32_t site, int V, int A, int ka, int T, float out_scale, void* dgates16, float* dpacked32,
                    float* dtext32, float* dW_hv, float* dW_ha, float* dW_v, float* dW_a, float* dbias4, float* dgamma,
                    float* dbeta, void* stream);

/* BertSelfOutput / BertOutput tail (transformers 3.0.2, inside bert.py:221-229):
 *   y = LayerNorm(dropout(pre) + res);  y16 optional fp16 copy; rstd optional [T] (needed by backward) */
int magb200_dropout_residual_ln_fwd(const float* pre, const float* res, const float* gamma, const float* beta,
                                    float eps, float p, uint64_t seed, uint32_t site, int T, float* y32, void* y16,
                                    float* rstd, void* stream);
/* BertSelfOutput / BertOutput as ONE kernel: y = LayerNorm(dropout(A W^T + bias) + res) with hidden size 768 -- the Linear
 * on tcgen05 (A16 [M][K] pitch lda, W16 [768][K], fp16 K-major) and dropout + residual + LayerNorm in its epilogue
 * (clusters of 2 x 3 CTAs own 256 rows x 768 columns and exchange the row statistics through distributed shared memory).
 * Same outputs and dropout mask as magb200_gemm(EPI_STORE_F32) -> magb200_dropout_residual_ln_fwd: y32 [M][768],
 * y16 (optional) fp16 copy, rstd (optional) [M] for magb200_ln_bwd. */
int magb200_linear_ln_fwd(const void* A16, int lda, const void* W16, int M, int K, const float* bias, const float* res,
                          const float* gamma, const float* beta, float eps, float p, uint64_t seed, uint32_t site, float* y32,
                          void* y16, float* rstd, void* stream);
/* dz32 = gradient wrt (dropout(pre)+res); dpre16 = fp16 dropout-masked gradient wrt pre; dgamma/dbeta/dbias += */
int magb200_ln_bwd(const float* dy32, const float* y32, const float* gamma, const float* beta, const float* rstd,
                   float p, uint64_t seed, uint32_t site, int T, float out_scale, float* dz32, void* dpre16,
                   float* dgamma, float* dbeta, float* dbias, void* stream);
/* out[n] += sum_t in[t][n], fp16 input */
int magb200_colsum_f16(const void* in_f16, int T, int N, float out_scale, float* out, void* stream);

/* BertSelfAttention core (transformers 3.0.2, inside bert.py:221-229):
 *   ctx = dropout(softmax(q k^T / 8 + (1 - mask) * -10000)) v  per (batch, head)
 * qkv16 [B*S][3*heads*64] fp16 (q | k | v), mask [B][S] int64 (1 = attend) or NULL, ctx16 [B*S][heads*64]
 * lse (optional, [B][heads][S] fp32): forward writes max + log-sum-exp of every score row; with lse and ctx16
 * present the tensor-core kernels run and probabilities never leave registers: one tile per (batch, head) for
 * S <= 64, a flash-style loop over 64-key blocks with online softmax for longer sequences (S <= 512, the limit
 * of BERT's position table); the backward of the latter needs `workspace` (magb200_sdpa_workspace_bytes: one fp32
 * per (batch, head, query)).  Without lse the general CUDA-core kernels run (forward S <= 512, backward S <= 128).
 * dbias (optional) += out_scale * column sums of dqkv, i.e. the gradient of the fused query|key|value bias. */
int magb200_sdpa_fwd(const void* qkv16, const int64_t* mask, int B, int S, int heads, int head_dim, float p,
                     uint64_t seed, uint32_t site, void* ctx16, float* lse, void* stream);
int magb200_sdpa_bwd(const void* qkv16, const void* dctx16, const void* ctx16, const float* lse, const int64_t* mask,
                     int B, int S, int heads, int head_dim, float p, uint64_t seed, uint32_t site, float out_scale,
                     void* dqkv16, float* dbias, void* workspace, size_t workspace_bytes, void* stream);
size_t magb200_sdpa_workspace_bytes(int B, int S, int heads);

/* BertPooler tanh + dropout + classifier (bert.py:231, 304-307): pre [B][768] = pooler dense output */
int magb200_cls_fwd(const float* pre, const float* Wc, const float* bc, int B, int num_labels, float p, uint64_t seed,
                    uint32_t site, float* pooled, float* logits, void* stream);
int magb200_cls_bwd(const float* dlogits, const float* pooled, const float* Wc, int B, int num_labels, float p,
                    uint64_t seed, uint32_t site, float out_scale, void* dpre16, float* dWc, float* dbc, void* stream);
/* MSELoss()(logits.view(-1), labels.view(-1)) (multimodal_driver.py:371-373; bert.py:313-317):
 * loss[0] = mean((logits-labels)^2); dlogits (optional) = 2 (logits-labels) / n * grad_scale */
int magb200_mse_loss(const float* logits, const float* labels, int n, float grad_scale, float* loss, float* dlogits,
                     void* stream);
/* transformers 3.0.2 AdamW.step (multimodal_driver.py:345-350, 383) over a flat buffer of n (multiple of 256)
 * parameters: grads are multiplied by grad_unscale first; decay_chunk[i] != 0 applies weight decay to elements
 * [256 i, 256 i + 256).  Optionally zeroes the gradient and refreshes the fp16 copy p16 in the same pass. */
int magb200_adamw_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, void* p16,
                       const uint8_t* decay_chunk, int64_t n, float lr, float beta1, float beta2, float eps,
                       float weight_decay, int step, float grad_unscale, int zero_grad, const float* hyper_dev,
                       void* stream);
/* hyper_dev (optional, device): {lr, step_size = lr*sqrt(1-b2^t)/(1-b1^t)} read at run time instead of the
 * by-value lr/step, so a captured CUDA graph follows the LR schedule (multimodal_driver.py:345-350, 384). */

/* ------------------------------------------------------------------------------------------------
 * Whole-model schedules: MAG_BertForSequenceClassification.forward (bert.py:251-324) and the autograd
 * pass of multimodal_driver.py:378, as one enqueue each.
 *
 * Parameters live in ONE flat fp32 buffer `params` (fp16 mirror `params16` at the same element
 * offsets, gradients `grads` likewise); `off` is a host array of element offsets indexed by the
 * MAGB200_OFF_* constants, every tensor starting on a 256-element boundary.  Required adjacency:
 * query|key|value weights contiguous ([3H][H]) and their biases contiguous; b_hv|b_ha|b_v|b_a
 * contiguous.  params16 additionally holds the packed MAG gate matrix at off[MAGB200_OFF_MAG_PACKED16]. */
#define MAGB200_MAX_LAYERS 48
enum {
  MAGB200_OFF_WORD = 0, MAGB200_OFF_POS, MAGB200_OFF_TYPE, MAGB200_OFF_EMB_LN_W, MAGB200_OFF_EMB_LN_B,
  MAGB200_OFF_MAG_WHV, MAGB200_OFF_MAG_WHA, MAGB200_OFF_MAG_WV, MAGB200_OFF_MAG_WA, MAGB200_OFF_MAG_BIAS4,
  MAGB200_OFF_MAG_LN_W, MAGB200_OFF_MAG_LN_B, MAGB200_OFF_POOL_W, MAGB200_OFF_POOL_B, MAGB200_OFF_CLS_W,
  MAGB200_OFF_CLS_B, MAGB200_OFF_MAG_PACKED16, MAGB200_OFF_TOTAL, MAGB200_OFF_RESERVED0, MAGB200_OFF_RESERVED1,
  MAGB200_OFF_LAYER0 = 20
};
enum {
  MAGB200_LOFF_QKV_W = 0, MAGB200_LOFF_QKV_B, MAGB200_LOFF_AO_W, MAGB200_LOFF_AO_B, MAGB200_LOFF_LN1_W,
  MAGB200_LOFF_LN1_B, MAGB200_LOFF_F1_W, MAGB200_LOFF_F1_B, MAGB200_LOFF_F2_W, MAGB200_LOFF_F2_B,
  MAGB200_LOFF_LN2_W, MAGB200_LOFF_LN2_B, MAGB200_OFF_PER_LAYER = 12
};
enum { MAGB200_PEEK_EMB = 0, MAGB200_PEEK_MAG, MAGB200_PEEK_LAYER_OUT, MAGB200_PEEK_POOLED, MAGB200_PEEK_GATES };

typedef struct magb200_bert_dims {
  int32_t batch, seq, hidden, heads, inter, layers, vocab, max_pos, type_vocab;
  int32_t visual_dim, acoustic_dim, num_labels, ka;   /* ka = hidden + 128: padded K of the MAG gate GEMM */
  float ln_eps, mag_ln_eps, beta_shift;
  float p_hidden, p_attn, p_mag, p_cls;               /* dropout probabilities (0 = eval) */
  float grad_scale;                                    /* loss scale carried by every gradient in backward */
  uint64_t seed;                                       /* dropout seed of this step */
  const uint64_t* seed_dev;                            /* if non-NULL the kernels read the seed from this device word
                                                          at run time (lets a captured CUDA graph change seed per replay) */
} magb200_bert_dims;

/* Data-parallel optimizer step over NVLink peer memory (one node, world <= 8; multimodal_driver.py:378-386 under data
 * parallelism): every rank's flat fp32 parameters and gradients live in symmetric memory, grad_ptrs[q] / param_ptrs[q]
 * are rank q's buffers as mapped in THIS process (host arrays of device pointers, this rank included).
 * magb200_adamw_peer_step: for the 256-aligned slice [slice_begin, slice_end) this rank owns -- sums the gradients of all
 * ranks in rank order (P2P loads), applies the transformers-3.0.2 AdamW of magb200_adamw_step to its local moments, and
 * stores the updated parameters into every rank's buffer (P2P stores) plus the local fp16 mirror: reduce-scatter + AdamW +
 * all-gather in one kernel (with multicast_grad / multicast_param non-NULL -- the NVSwitch multicast addresses of the two
 * buffers -- the sum is a multimem.ld_reduce inside the switch and the broadcast one multimem.st); max_blocks > 0 bounds the grid (SMs left to concurrent GEMMs).  [clean_begin, clean_end) is a
 * range all ranks finished earlier (a barrier lies between): its local gradients are cleared and the parameters the peers
 * stored into it are recast into the local fp16 mirror (empty range: nothing; clean_what bit 0 = recast, bit 1 = clear;
 * [keep_begin, keep_end) inside it is not recast -- the caller's own piece, whose mirror its update wrote).
 * magb200_peer_barrier: flag barrier between the ranks (flag_ptrs[q]: rank q's flag array of >= world uint32 in symmetric
 * memory, zero-initialised; epoch_counter: one local device uint32, zero-initialised, advanced by the kernel itself so a
 * captured graph can be replayed). */
int magb200_peer_barrier(void* const* flag_ptrs, int world, int rank, void* epoch_counter, void* stream);
int magb200_adamw_peer_step(const void* const* grad_ptrs, void* const* param_ptrs, int world, int rank, float* exp_avg,
                            float* exp_avg_sq, void* p16, const uint8_t* decay_chunk, int64_t slice_begin, int64_t slice_end,
                            int64_t clean_begin, int64_t clean_end, float lr, float beta1, float beta2, float eps,
                            float weight_decay, int step, float grad_unscale, const float* hyper_dev, int max_blocks,
                            const uint8_t* row_any, int row_mode, int64_t table_begin, int table_rows, int row_len,
                            int clean_what, int64_t keep_begin, int64_t keep_end, const void* multicast_grad,
                            void* multicast_param, void* stream);
/* Row-sparse table inside a bucket (the word-embedding matrix, bert.py:211 / nn.Embedding backward: only the rows of the
 * step's token ids receive a gradient).  magb200_peer_mark_rows stamps marks[id] (uint32[vocab] in symmetric memory) for
 * the batch's ids with (*counter + 1) and advances the device counter; after a barrier magb200_peer_combine_marks writes
 * row_any[row] = 1 where any rank stamped the row.  magb200_adamw_peer_step then takes row_mode 1 (before backward: update
 * and broadcast the table rows NO rank touched -- their gradient is exactly zero; skip everything else) or row_mode 2
 * (after backward: touched rows and everything outside the table); row_mode 0 = dense. */
int magb200_peer_mark_rows(const int64_t* ids, int n, int vocab, void* marks, void* counter, void* stream);
int magb200_peer_combine_marks(const void* const* mark_ptrs, int world, const void* counter, int vocab, uint8_t* row_any,
                               void* stream);

size_t magb200_bert_workspace_bytes(const magb200_bert_dims* dims);
/* params -> params16 (cast) + packed MAG gate matrix; call after the master weights change */
int magb200_bert_refresh_operands(const magb200_bert_dims* dims, const int64_t* off, const float* params,
                                  void* params16, void* stream);
int magb200_bert_forward(const magb200_bert_dims* dims, const int64_t* off, const float* params, const void* params16,
                         const int64_t* input_ids, const int64_t* token_type_ids, const int64_t* attention_mask,
                         const float* visual, const float* acoustic, void* workspace, size_t workspace_bytes,
                         float* logits, void* stream);
/* dlogits_scaled = dL/dlogits * dims->grad_scale (activation gradients are loss-scaled fp16 inside);
 * grads (+=) receive dL/dparam, unscaled.  Backward kernels that write parameter gradients take
 * out_scale (= 1/grad_scale) for the same reason. */
int magb200_bert_backward(const magb200_bert_dims* dims, const int64_t* off, const float* params, const void* params16,
                          const int64_t* input_ids, const int64_t* token_type_ids, const int64_t* attention_mask,
                          void* workspace, size_t workspace_bytes, const float* dlogits_scaled, float* grads,
                          void* stream);
/* The same backward in pieces, so a data-parallel caller can all-reduce the gradient range of finished layers
 * while the next piece runs: layers layer_hi .. layer_lo (descending); flags bit 0 = classifier + pooler first,
 * bit 1 = MAG + embeddings last.  magb200_bert_backward == part(layers-1, 0, 3). */
int magb200_bert_backward_part(const magb200_bert_dims* dims, const int64_t* off, const float* params,
                               const void* params16, const int64_t* input_ids, const int64_t* token_type_ids,
                               const int64_t* attention_mask, void* workspace, size_t workspace_bytes,
                               const float* dlogits_scaled, float* grads, int layer_hi, int layer_lo, int flags,
                               void* stream);
/* Backward + optimizer of one training step as ONE schedule (multimodal_driver.py:378 + :383 for
 * gradient_accumulation_step == 1, single process): identical results to magb200_bert_backward followed by
 * magb200_adamw_step over the whole flat buffer, but the AdamW update of each encoder layer's slice is enqueued on an
 * internal stream as soon as that layer's gradients are final, overlapping the HBM-bound update with the GEMM-bound
 * backward of the layers below.  All work is ordered after what `stream` held at the call and joined back into
 * `stream` before returning (legal under stream capture).  Optimizer arguments as magb200_adamw_step. */
int magb200_bert_backward_adamw(const magb200_bert_dims* dims, const int64_t* off, float* params, void* params16,
                                const int64_t* input_ids, const int64_t* token_type_ids, const int64_t* attention_mask,
                                void* workspace, size_t workspace_bytes, const float* dlogits_scaled, float* grads,
                                float* exp_avg, float* exp_avg_sq, const uint8_t* decay_chunk, float lr, float beta1,
                                float beta2, float eps, float weight_decay, int step, float grad_unscale, int zero_grad,
                                const float* hyper_dev, void* stream);
/* output_attentions (bert.py:122-124, 233-237): the attention probabilities of one layer, [B][heads][S][S] fp32 after
 * dropout, recomputed from the q|k|v activations a forward left in `workspace` (the fused attention kernels never
 * materialise them). */
int magb200_bert_attention_probs(const magb200_bert_dims* dims, void* workspace, int layer, const int64_t* attention_mask,
                                 float* out, size_t out_bytes, void* stream);
/* copy a saved activation out of the workspace (tests): MAGB200_PEEK_* */
int magb200_bert_peek(const magb200_bert_dims* dims, void* workspace, int what, int layer, void* out, size_t out_bytes,
                      void* stream);


/* ------------------------------------------------------------------------------------------------
 * MAG-XLNet: MAG_XLNetForSequenceClassification.forward (xlnet.py:443-527) through
 * MAG_XLNetModel.forward (xlnet.py:148-429), content stream only (target_mapping / perm_mask / mems
 * None -- the path multimodal_driver.py exercises), and its backward.  Tokens are batch-major
 * (t = b*S + s); the reference's seq-first transposes (xlnet.py:206-239) are layout only.
 * Flat parameter buffer as for BERT; projection tensors q,k,v,r are [d_model][heads*64] ("in x out"),
 * o is [d_model][heads*64] read as [out][in].  params16 = fp16 mirror + packed MAG gate matrix
 * (MAGB200_XOFF_MAG_PACKED16) + per layer the packed [768][2304] q|k|v operand (MAGB200_XLOFF_QKV_PACKED16). */
enum {
  MAGB200_XOFF_WORD = 0, MAGB200_XOFF_MAG_WHV, MAGB200_XOFF_MAG_WHA, MAGB200_XOFF_MAG_WV, MAGB200_XOFF_MAG_WA,
  MAGB200_XOFF_MAG_BIAS4, MAGB200_XOFF_MAG_LN_W, MAGB200_XOFF_MAG_LN_B, MAGB200_XOFF_SUM_W, MAGB200_XOFF_SUM_B,
  MAGB200_XOFF_CLS_W, MAGB200_XOFF_CLS_B, MAGB200_XOFF_MAG_PACKED16, MAGB200_XOFF_TOTAL,
  MAGB200_XOFF_LAYER0 = 16
};
enum {
  MAGB200_XLOFF_Q = 0, MAGB200_XLOFF_K, MAGB200_XLOFF_V, MAGB200_XLOFF_O, MAGB200_XLOFF_R, MAGB200_XLOFF_RR_BIAS,
  MAGB200_XLOFF_RS_BIAS, MAGB200_XLOFF_RW_BIAS, MAGB200_XLOFF_SEG, MAGB200_XLOFF_LN1_W, MAGB200_XLOFF_LN1_B,
  MAGB200_XLOFF_F1_W, MAGB200_XLOFF_F1_B, MAGB200_XLOFF_F2_W, MAGB200_XLOFF_F2_B, MAGB200_XLOFF_LN2_W,
  MAGB200_XLOFF_LN2_B, MAGB200_XLOFF_QKV_PACKED16, MAGB200_XOFF_PER_LAYER = 18
};

typedef struct magb200_xlnet_dims {
  int32_t batch, seq, hidden, heads, inter, layers, vocab, visual_dim, acoustic_dim, num_labels, ka, injection_index;
  float ln_eps, mag_ln_eps, beta_shift;
  float p_drop, p_mag, p_summary;                      /* config.dropout, MAG dropout, summary_last_dropout (0 = eval) */
  float grad_scale;
  uint64_t seed;
  const uint64_t* seed_dev;
} magb200_xlnet_dims;

size_t magb200_xlnet_workspace_bytes(const magb200_xlnet_dims* dims);
int magb200_xlnet_refresh_operands(const magb200_xlnet_dims* dims, const int64_t* off, const float* params,
                                   void* params16, void* stream);
int magb200_xlnet_forward(const magb200_xlnet_dims* dims, const int64_t* off, const float* params, const void* params16,
                          const int64_t* input_ids, const int64_t* token_type_ids, const int64_t* attention_mask,
                          const float* visual, const float* acoustic, void* workspace, size_t workspace_bytes,
                          float* logits, void* stream);
int magb200_xlnet_backward(const magb200_xlnet_dims* dims, const int64_t* off, const float* params, const void* params16,
                           const int64_t* input_ids, const int64_t* token_type_ids, const int64_t* attention_mask,
                           void* workspace, size_t workspace_bytes, const float* dlogits_scaled, float* grads,
                           void* stream);
/* what: MAGB200_PEEK_MAG (MAG output), MAGB200_PEEK_LAYER_OUT (layer), MAGB200_PEEK_POOLED (summary) */
int magb200_xlnet_peek(const magb200_xlnet_dims* dims, void* workspace, int what, int layer, void* out,
                       size_t out_bytes, void* stream);

/* XLNetRelativeAttention core (transformers 3.0.2 rel_attn_core + rel_shift_bnij, called through xlnet.py:374-385):
 *   score[i][j] = ((q_i + r_w_bias).k_j + (q_i + r_r_bias).kr[j + S - i] + (q_i + r_s_bias).seg_embed[tt_i != tt_j]) / 8
 *                 - 1e30 * (pad_j and i != j);  ctx = dropout(softmax(score)) v          S <= 64, head_dim 64
 * kr16: projected relative positions [Bp][2S][heads*64] (Bp = 1 shared, or B when the position dropout differs per
 * batch element). */
int magb200_xlnet_rel_attn_fwd(const void* qkv16, const void* kr16, int Bp, const float* r_w_bias, const float* r_r_bias,
                               const float* r_s_bias, const float* seg_embed, const int64_t* attention_mask,
                               const int64_t* token_type_ids, int B, int S, int heads, int head_dim, float p, uint64_t seed,
                               uint32_t site, void* ctx16, void* stream);
/* dkr32 (fp32, [Bp][2S][heads*64]) and the four small parameter gradients are accumulated (+=); dqkv16 is written */
int magb200_xlnet_rel_attn_bwd(const void* qkv16, const void* kr16, int Bp, const float* r_w_bias, const float* r_r_bias,
                               const float* r_s_bias, const float* seg_embed, const int64_t* attention_mask,
                               const int64_t* token_type_ids, const void* dctx16, int B, int S, int heads, int head_dim,
                               float p, uint64_t seed, uint32_t site, float out_scale, void* dqkv16, float* dkr32,
                               float* d_r_w_bias, float* d_r_r_bias, float* d_r_s_bias, float* d_seg_embed, void* stream);
/* relative positional encoding (xlnet.py:93-146, attn_type "bi", bi_data False, clamp_len -1): pos16 [Bp][2S][768] */
int magb200_xlnet_pos_emb(int S, int Bp, float p, uint64_t seed, uint32_t site, void* pos16, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MAGB200_H_ */

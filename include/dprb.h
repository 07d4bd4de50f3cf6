/* dprb.h — C ABI of libdprb.so: the B200 (sm_100a) drop-in for the arithmetic of dpr-scale's
 * bi-encoder contrastive training step.
 *
 * The reference (facebookresearch/dpr-scale) is pure Python and has no FFI of its own; every FLOP on
 * the path is delegated to torch / HuggingFace modules.  Each entry point below therefore cites the
 * reference call site (file:line under /root/reference, or site-packages/transformers for the
 * third-party code the reference calls) whose arithmetic it replaces.  INTEGRATION.md shows the
 * ctypes binding a dpr-scale maintainer would add.
 *
 * Conventions
 *  - Every function returns 0 on success; non-zero = error, text via dprb_last_error() (thread-local).
 *  - All pointers are DEVICE pointers owned by the caller (PyTorch caching allocator); the library
 *    allocates nothing persistent and keeps no global mutable state besides cached device attributes.
 *  - `stream` is a cudaStream_t; all work is enqueued on it and no call synchronises the device.
 *  - bf16 tensors are row-major with 16-byte aligned rows; H, I multiples of 8; head_dim == 64.
 *  - One process per GPU; collectives are NOT issued here (torch.distributed/NCCL does that).
 */
#ifndef DPRB_H_
#define DPRB_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* dprb_stream_t; /* cudaStream_t */

#define DPRB_VERSION 100

int dprb_version(void);
const char* dprb_last_error(void);
/* SM count of the current device (cached); <=0 when no device. */
int dprb_num_sms(void);
/* Number of kernels this library has launched in this process since load (every <<<>>> / cudaLaunchKernelEx site
 * increments it).  bench.py reports the difference over its timed region as `gpu_launches` - a count, not an estimate.
 * No reference counterpart (the reference launches through PyTorch). */
int64_t dprb_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * GEMM (tcgen05 / TMA / TMEM):  D[M,N] = epilogue( alpha * sum_k A(m,k) * B(n,k) )
 * Replaces torch.nn.Linear forward/backward inside HF BertLayer:
 *   site-packages/transformers/models/bert/modeling_bert.py:179-181 (Q,K,V — fused here into one
 *   [3H,H] weight), :295 (attention output dense), :340 (intermediate dense), :353 (output dense).
 * colsum (optional, bf16 epilogues except BIAS_GELU): colsum[n] += sum_m D(m,n) — the bias gradient of the
 * Linear whose output gradient this GEMM produces, fused into the epilogue instead of a separate pass.
 * dropout_p / dropout_site_seed (BIAS_RESIDUAL only): D = dropout(acc + bias) + aux — HF's hidden dropout between the
 * dense layer and the residual add (modeling_bert.py:296-297, :354-355); the mask is a counter-based hash of the
 * element index (never stored; dprb_ln_bwd re-derives it, dprb_dropout_mask exports it for tests).
 * a_mn_major/b_mn_major = 0: operand stored [MN, K] (K contiguous, leading dim ld);
 *                       = 1: operand stored [K, MN] (MN contiguous, leading dim ld).
 * ------------------------------------------------------------------------------------------- */
enum {
  DPRB_EPI_BIAS = 0,           /* D(bf16) = acc + bias                      (bias may be NULL)        */
  DPRB_EPI_BIAS_GELU = 1,      /* D(bf16) = gelu(acc + bias) ; out2(bf16, optional) = gelu'(acc + bias) */
  DPRB_EPI_BIAS_RESIDUAL = 2,  /* D(bf16) = acc + bias + aux(bf16)                                     */
  DPRB_EPI_DGELU = 3,          /* D(bf16) = acc * aux(bf16), aux = the gelu' saved by BIAS_GELU          */
  DPRB_EPI_F32_ATOMIC_ADD = 4, /* D(fp32) += acc  (split-K over `splits` CTAs; 0 = choose)            */
  DPRB_EPI_F32_STORE = 5,      /* D(fp32) = acc + bias                                                 */
  DPRB_EPI_DGELU_PRE = 6,      /* D(bf16) = acc * gelu'(aux), aux(bf16) = the PRE-activation saved by BIAS_GELU|SAVE_PRE */
  DPRB_EPI_COUNT = 7
};
/* OR-ed into `epilogue`: the named 16-bit operand holds IEEE fp16 instead of bf16 (tcgen05 kind::f16 takes either
 * format per operand).  The encoder keeps its residual stream - LayerNorm inputs and outputs - in fp16 (11 significand
 * bits in the same 2 bytes): these are the tensors HF's autocast keeps in fp32 (modeling_bert.py:296-298, :354-356
 * run LayerNorm and the residual add outside the 16-bit region).  AUX / OUT apply to the BIAS and BIAS_RESIDUAL
 * epilogues. */
enum { DPRB_GEMM_A_F16 = 0x100, DPRB_GEMM_B_F16 = 0x200, DPRB_GEMM_AUX_F16 = 0x400, DPRB_GEMM_OUT_F16 = 0x800,
       /* BIAS_GELU: out2 receives the pre-activation instead of gelu'(pre) - the "lean activations" mode of the encoder,
        * which saves ONE [T, I] tensor per layer (pre) instead of two (gelu', gelu) and rebuilds both in backward. */
       DPRB_GEMM_SAVE_PRE = 0x1000 };
/* (A_F16 and B_F16 must be given together: the hardware rejects an fp16 x bf16 operand pair.) */
int dprb_gemm_bf16(const void* A, const void* B, void* D, int M, int N, int K, int64_t lda, int64_t ldb,
                   int64_t ldd, int a_mn_major, int b_mn_major, int epilogue, const float* bias,
                   const void* aux, int64_t ld_aux, void* out2, float alpha, int splits, float* colsum,
                   float dropout_p, uint64_t dropout_site_seed, dprb_stream_t stream);

/* Measurement aid (bench.py roofline leg): when enabled, every GEMM launch is bracketed by CUDA events on
 * its launch stream; dprb_gemm_profile_read sums the per-launch durations and algorithmic FLOPs (2*M*N*K). */
int dprb_gemm_profile_enable(int enable, int max_launches);
int dprb_gemm_profile_read(double* total_ms, double* total_flops, int64_t* launches);

/* ---------------------------------------------------------------------------------------------
 * Embeddings + LayerNorm.  Replaces BertEmbeddings.forward (modeling_bert.py:72-112):
 *   z = word[ids] + type[type_ids] + pos[pos_ids];  y = LN(z) (eps, gamma, beta).
 * Tables and LN parameters are the fp32 master weights.  stats[t] = (mean, rstd).
 * ------------------------------------------------------------------------------------------- */
int dprb_embed_ln_fwd(const int64_t* ids, const int64_t* type_ids, const int64_t* pos_ids, const float* word,
                      const float* pos, const float* type, const float* gamma, const float* beta, void* y_bf16,
                      float* stats, int T, int H, int vocab, int max_pos, int type_vocab, float eps,
                      float dropout_p, uint64_t dropout_seed, void* y_res_f16, dprb_stream_t stream);
/* The residual stream in fp16.  y_res_f16 (here and in dprb_ln_fwd, optional): a second copy of the LayerNorm output
 * in IEEE fp16, read by the NEXT residual add (DPRB_GEMM_AUX_F16) while the bf16 copy y feeds the next GEMM - kind::f16
 * cannot mix fp16 activations with bf16 weights, and HF's autocast keeps exactly this path (LayerNorm output ->
 * residual add -> LayerNorm input) out of 16-bit bf16.  z_f16 (dprb_ln_fwd / dprb_ln_bwd): the pre-LayerNorm sum z,
 * written by a DPRB_GEMM_OUT_F16 epilogue, holds fp16. */
/* Backward: dz = LN'(dy); dgamma/dbeta accumulated; scatter-add of dz into the three table grads. */
int dprb_embed_ln_bwd(const void* dy_bf16, const int64_t* ids, const int64_t* type_ids, const int64_t* pos_ids,
                      const float* word, const float* pos, const float* type, const float* gamma,
                      const float* stats, float* dword, float* dpos, float* dtype, float* dgamma, float* dbeta,
                      int T, int H, float dropout_p, uint64_t dropout_seed, dprb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * LayerNorm over rows of z (the residual sum is produced by the preceding GEMM epilogue).
 * Replaces BertSelfOutput / BertOutput LayerNorm (modeling_bert.py:294-298, :352-356).
 * If cls_out != NULL, rows t with t % cls_stride == 0 are also written in fp32 to
 * cls_out[t / cls_stride, :] — the CLS pooling + .clone() of hf_model.py:39-41.
 * ------------------------------------------------------------------------------------------- */
int dprb_ln_fwd(const void* z_bf16, const float* gamma, const float* beta, void* y_bf16, float* stats,
                float* cls_out, int cls_stride, int T, int H, float eps, int z_f16, void* y_res_f16,
                dprb_stream_t stream);
/* dz = LN'(dy; z, stats); dgamma += sum dy*xhat; dbeta += sum dy; if dbias != NULL: dbias += sum_t dz
 * (the bias gradient of the Linear that produced z).  If dy_cls != NULL, dy is implicit: zero
 * everywhere except rows t % cls_stride == 0 which take dy_cls[t / cls_stride, :] (fp32). */
/* With hidden dropout (dropout_p > 0): dzm_bf16 receives dz * mask/(1-p) — the gradient of the Linear output that was
 * dropped before the residual add — and dbias sums dzm instead of dz. */
int dprb_ln_bwd(const void* dy_bf16, const float* dy_cls, int cls_stride, const void* z_bf16,
                const float* stats, const float* gamma, void* dz_bf16, float* dgamma, float* dbeta,
                float* dbias, int T, int H, void* dzm_bf16, float dropout_p, uint64_t dropout_site_seed,
                int z_f16, dprb_stream_t stream);
/* Dropout sites: 0 embeddings [T,H], 1 attention probabilities [nseq*heads*S, S], 2 attention-output dense [T,H],
 * 3 FFN-output dense [T,H].  Element (r, c) of a site is kept iff its 16-bit lane of h(r, c/8, (c/2)%4, site seed)
 * is >= round(p * 65536) - one hash chain per group of 8 columns, one multiply-xorshift finaliser per column pair
 * (csrc/common.cuh: Drop); site seed = fold32(dropout_seed + (layer*8 + site + 1) * 0x9E3779B97F4A7C15).
 * dprb_ln_bwd: dbias accumulates the column sums of the Linear's own output gradient (dzm when dropout is on).
 * dprb_dropout_mask materialises keep[r * cols + c] of one site (test aid). */
uint64_t dprb_dropout_site_seed(uint64_t dropout_seed, int layer, int site);
int dprb_dropout_mask(uint8_t* keep, int64_t rows, int cols, float dropout_p, uint64_t dropout_seed, int layer,
                      int site, dprb_stream_t stream);

/* out(bf16) = gelu(pre(bf16)), elementwise over n (multiple of 8) values: rebuilds BertIntermediate's activation
 * (modeling_bert.py:339-342) in backward when the encoder ran with lean activations (save_for_backward = 2). */
int dprb_gelu_from_pre(const void* pre_bf16, void* out_bf16, int64_t n, dprb_stream_t stream);
/* Column sums: out[n] += sum_t x[t, n]  (bias gradients; x bf16 [T, N] with leading dim ld). */
int dprb_colsum_bf16(const void* x_bf16, int64_t ld, float* out, int T, int N, dprb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Self-attention core for head_dim 64, S <= 256.  Replaces BertSelfAttention.forward's
 * scaled_dot_product_attention (modeling_bert.py:168-207, integrations/sdpa_attention.py:92-101):
 *   ctx = softmax(Q K^T / 8 + key_mask) V      per (sequence, head).
 * qkv: bf16 [nseq*S, 3H] = [Q | K | V] column blocks, head h at columns h*64.
 * attn_mask: int32 [nseq, S], 1 = real token, 0 = padding (HF attention_mask); may be NULL.
 * lse: fp32 [nseq, heads, S] natural-log row log-sum-exp, written by fwd (may be NULL for
 *      forward-only use) and consumed by bwd together with the forward output ctx.
 * ------------------------------------------------------------------------------------------- */
int dprb_attn_fwd(const void* qkv_bf16, const int32_t* attn_mask, void* ctx_bf16, float* lse, int nseq, int S,
                  int heads, float dropout_p, uint64_t dropout_site_seed, dprb_stream_t stream);
/* dbias (optional, fp32 [3H]): dbias[n] += sum_t dqkv[t, n] — the bias gradient of the fused QKV projection. */
int dprb_attn_bwd(const void* qkv_bf16, const int32_t* attn_mask, const void* ctx_bf16, const float* lse,
                  const void* dctx_bf16, void* dqkv_bf16, float* dbias, int nseq, int S, int heads,
                  float dropout_p, uint64_t dropout_site_seed, dprb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Fused in-batch-negative scoring + softmax cross-entropy.
 * Replaces dpr_scale/task/dpr_task.py:98-105 (sim_score: q @ c.T, scores[mask] = -inf),
 * :197 (mask.repeat), :211 (scores /= T), :212 (nn.CrossEntropyLoss, mean over Q).
 *   q fp32 [Q,d], c fp32 [C,d], col_mask u8 [C] (1 = dummy ctx -> -inf), labels i64 [Q];
 *   pair_mask u8 [Q,C] (optional, 1 -> -inf): the per-query block mask of the non-in-batch branch (:199-207).
 * Outputs: lse[Q], loss_sum (sum over rows of lse - logit[label]; caller divides by Q),
 *          logits fp32 [Q,C] (masked columns = -inf) if non-NULL; required when backward follows.
 * Backward of mean-over-Q loss (grad_scale = upstream dL, normally 1):
 *   dq[q0:q0+nq, :]  (rows owned by this rank)  and  dc[c0:c0+nc, :] (columns owned by this rank),
 *   reproducing dpr_task.py:163-195 where remote slices are detached constants.
 * ------------------------------------------------------------------------------------------- */
int dprb_score_ce_fwd(const float* q, const float* c, const uint8_t* col_mask, const uint8_t* pair_mask,
                      const int64_t* labels, float inv_temperature, float* lse, float* loss_sum, float* logits,
                      int Q, int C, int d, dprb_stream_t stream);
int dprb_score_ce_bwd(const float* q, const float* c, const float* logits, const int64_t* labels,
                      const float* lse, float grad_scale, float inv_temperature, float* dq, float* dc, int Q,
                      int C, int d, int q0, int nq, int c0, int nc, dprb_stream_t stream);

/* The same operator on the tensor cores, in ONE pass (the form BASELINE.json's north_star names): similarity tile via
 * tcgen05.mma into TMEM, online row max / sum-exp / label pick straight from tcgen05.ld, loss accumulated by the last
 * tile of each row block; logits reach HBM only if `logits` is non-NULL.  fp32 fidelity comes from an exact 3-way bf16
 * split of q and c (six partial products per k-block, fp32 accumulate): logits agree with the fp32 product of
 * dpr_task.py:99 to ~1e-6 relative.  Backward RECOMPUTES the tiles of the rank-local row block and column block
 * (no stored logits) and runs dq = W_rows c, dc = W_cols^T q on the tcgen05 GEMM.
 *   nq / nc: the local row / column counts backward will ask for (sizes the workspace; -1 = all).
 *   workspace: caller-owned, 256-byte aligned, >= dprb_score_tc_workspace_bytes(...); it carries the operand splits
 *   from the forward call to the backward call of the same step.
 * Requires d % 8 == 0 (dprb_score_tc_supported); other shapes use dprb_score_ce_fwd/bwd above. */
int dprb_score_tc_supported(int Q, int C, int d);
int64_t dprb_score_tc_workspace_bytes(int Q, int C, int d, int nq, int nc);
int dprb_score_tc_fwd(const float* q, const float* c, const uint8_t* col_mask, const uint8_t* pair_mask,
                      const int64_t* labels, float inv_temperature, float* lse, float* loss_sum, float* logits, int Q,
                      int C, int d, int nq, int nc, void* workspace, int64_t workspace_bytes, dprb_stream_t stream);
int dprb_score_tc_bwd(const uint8_t* col_mask, const uint8_t* pair_mask, const int64_t* labels, const float* lse,
                      float grad_scale, float inv_temperature, float* dq, float* dc, int Q, int C, int d, int q0, int nq,
                      int c0, int nc, void* workspace, int64_t workspace_bytes, dprb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Optimizer step over a flat fp32 parameter arena.  Replaces torch.optim.AdamW
 * (conf/task/optim/adamw.yaml via dpr_task.py:124) + clip_grad_norm_(2.0)
 * (conf/trainer/gpu_1_host.yaml:8) + the bf16 weight shadow refresh.
 *   dprb_sumsq: out[0] += sum g^2.
 *   dprb_adamw_step: coef = min(1, max_norm / (sqrt(*sumsq) * grad_div_inv... see DESIGN.md) applied to g.
 * ------------------------------------------------------------------------------------------- */
int dprb_sumsq_f32(const float* g, int64_t n, float* out, dprb_stream_t stream);
int dprb_adamw_step(float* p, const float* g, float* m, float* v, void* shadow_bf16, int64_t n, float lr,
                    float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                    const float* sumsq, float max_norm, dprb_stream_t stream);
/* fp32 -> bf16 shadow refresh (after a state_dict load). */
int dprb_cast_f32_bf16(const float* src, void* dst_bf16, int64_t n, dprb_stream_t stream);
/* bf16 -> fp32: unpacks a bf16-compressed gradient slice after its all-reduce (the `fp16_grads` path:
 * torch's fp16_compress_hook registered at dpr_scale/task/dpr_task.py:90-92 casts, all-reduces and casts back). */
int dprb_cast_bf16_f32(const void* src_bf16, float* dst, int64_t n, dprb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Whole-encoder forward / backward: the BertModel stack of modeling_bert.py:628-691 as called from
 * dpr_scale/models/hf_model.py:36-41, minus the unused pooler.  See dprb_encoder.h-style struct below.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  /* dims */
  int32_t hidden, inter, layers, heads, vocab, max_pos, type_vocab;
  float ln_eps;
  /* flat arenas (see dprb_param_offsets): fp32 master, bf16 shadow, fp32 grads */
  const float* master;
  const void* shadow;
  float* grads;
  /* offsets (in elements) into the arenas */
  int64_t off_word, off_pos, off_type, off_emb_ln_g, off_emb_ln_b;
  int64_t off_layer0;     /* first layer block */
  int64_t layer_stride;   /* elements per layer block */
  /* per-layer relative offsets */
  int64_t rel_wqkv, rel_bqkv, rel_wo, rel_bo, rel_ln1_g, rel_ln1_b, rel_w1, rel_b1, rel_w2, rel_b2, rel_ln2_g,
      rel_ln2_b;
} dprb_encoder_weights;

typedef struct {
  int32_t nseq, S;
  const int64_t* ids;      /* [nseq*S] */
  const int64_t* type_ids; /* [nseq*S] */
  const int64_t* pos_ids;  /* [nseq*S] */
  const int32_t* attn_mask;/* [nseq*S] or NULL */
  /* activation workspace, caller-allocated: see dprb_encoder_workspace_bytes */
  void* workspace;
  int64_t workspace_bytes;
  int32_t save_for_backward; /* 0: forward-only (generate_embeddings path) reuses per-layer buffers; 1: keep every
                              * activation backward reads; 2: "lean" - per layer keep qkv, the two pre-LayerNorm sums,
                              * x1, the FFN pre-activation and the layer output, and REBUILD the attention output
                              * (one more attention forward) and gelu / gelu' (from the pre-activation) in backward:
                              * 22 KB instead of 32 KB per token and layer at RoBERTa-large, which is what lets
                              * BASELINE config 4 (278 528 tokens x 24 layers per GPU) fit 180 GB without recomputing
                              * the whole forward. */
  float dropout_p;           /* hidden + attention-probability dropout (HFEncoder's `dropout`); 0 in eval mode */
  uint64_t dropout_seed;     /* per-forward seed; backward must be given the same value */
} dprb_encoder_batch;

int64_t dprb_encoder_workspace_bytes(const dprb_encoder_weights* w, int nseq, int S, int save_for_backward);
/* pooled fp32 [nseq, hidden] = last-layer hidden state of token 0 of each sequence. */
int dprb_encoder_fwd(const dprb_encoder_weights* w, const dprb_encoder_batch* b, float* pooled,
                     dprb_stream_t stream);
/* Accumulates parameter gradients into w->grads given dpooled fp32 [nseq, hidden].
 * Layers [layer_hi-1 .. layer_lo] are processed (layer_lo == 0 also runs the embedding backward), so
 * the host can interleave gradient all-reduce buckets between calls. */
int dprb_encoder_bwd(const dprb_encoder_weights* w, const dprb_encoder_batch* b, const float* dpooled,
                     int layer_lo, int layer_hi, dprb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Brute-force retrieval: replaces search_index() of
 * dpr_scale/run_retrieval_pytorch.py:141-176 -- einsum('ik,jk->ij') in fp16 followed by torch.topk over the
 * materialised [Q, N] score matrix -- with one fused pass: the corpus is streamed once per block of 128 queries
 * through tcgen05 and a running top-k is kept per query; no score matrix is written.
 *   queries [Q, d], corpus [N, d]: row-major 16-bit (dtype 0 = fp16 as in build_index() :178-190, 1 = bf16),
 *   d % 8 == 0, 16-byte aligned, N < 2^31 - 256, 1 <= k <= min(N, 1024).
 *   out_scores [Q, k] fp32 (fp32-accumulated inner products, descending; ties towards the lower row id),
 *   out_index  [Q, k] int64 = corpus row id + index_offset (the shard offset of :225-227).
 *   workspace: >= dprb_search_workspace_bytes(Q, k) bytes of device memory, caller-owned.
 *   dtype | DPRB_SEARCH_RANK_FP16: rank by (and return) the score ROUNDED TO fp16 - what the reference's topk sees,
 *   because its einsum on fp16 tensors returns fp16 (:150-151).  Ids then equal the reference's wherever its fp16
 *   scores are distinct; the default ranks by the exact fp32-accumulated score (a finer, deterministic order).
 * dprb_topk_merge replaces the per-shard merge of :272-277 (topk over the concatenated shard results + gather):
 *   scores / index [Q, total] -> the k best per row (ties towards the earlier position), workspace
 *   >= dprb_topk_merge_workspace_bytes(Q, total).
 * ------------------------------------------------------------------------------------------- */
enum { DPRB_SEARCH_RANK_FP16 = 0x100 };
int64_t dprb_search_workspace_bytes(int64_t Q, int k);
int dprb_search_topk(const void* queries, const void* corpus, int dtype, int64_t Q, int64_t N, int d, int k,
                     int64_t index_offset, float* out_scores, int64_t* out_index, void* workspace,
                     int64_t workspace_bytes, dprb_stream_t stream);
int64_t dprb_topk_merge_workspace_bytes(int64_t Q, int total);
int dprb_topk_merge(const float* scores, const int64_t* index, int64_t Q, int total, int k, float* out_scores,
                    int64_t* out_index, void* workspace, int64_t workspace_bytes, dprb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DPRB_H_ */

// Internal C++ declarations shared by the kernel translation units of libdprb.so.
#pragma once
#include <cuda_runtime.h>
#include "../../include/dprb.h"

namespace dprb {

int gemm_bf16(const void* A, const void* B, void* D, int M, int N, int K, long long lda, long long ldb,
              long long ldd, int a_mn_major, int b_mn_major, int epilogue, const float* bias, const void* aux,
              long long ld_aux, void* out2, float alpha, int splits, float* colsum, float dropout_p,
              unsigned long long drop_site_seed, cudaStream_t stream);

int gemm_profile_enable(int enable, int max_launches);
int gemm_profile_read(double* total_ms, double* total_flops, long long* launches);

int embed_ln_fwd(const int64_t* ids, const int64_t* type_ids, const int64_t* pos_ids, const float* word,
                 const float* pos, const float* type, const float* gamma, const float* beta, void* y, float* stats,
                 int T, int H, int vocab, int max_pos, int type_vocab, float eps, float dropout_p,
                 unsigned long long seed, void* y_res, cudaStream_t stream);
int embed_ln_bwd(const void* dy, const int64_t* ids, const int64_t* type_ids, const int64_t* pos_ids,
                 const float* word, const float* pos, const float* type, const float* gamma, const float* stats,
                 float* dword, float* dpos, float* dtype, float* dgamma, float* dbeta, int T, int H,
                 float dropout_p, unsigned long long seed, cudaStream_t stream);
int ln_fwd(const void* z, const float* gamma, const float* beta, void* y, float* stats, float* cls_out,
           int cls_stride, int T, int H, float eps, int z_f16, void* y_res, cudaStream_t stream);
int ln_bwd(const void* dy, const float* dy_cls, int cls_stride, const void* z, const float* stats,
           const float* gamma, void* dz, float* dgamma, float* dbeta, float* dbias, int T, int H, void* dzm,
           float dropout_p, unsigned long long site_seed, int z_f16, cudaStream_t stream);
int dropout_mask(uint8_t* out, long long rows, int cols, float p, unsigned long long seed, int layer, int site,
                 cudaStream_t stream);
unsigned long long drop_site_seed(unsigned long long seed, int layer, int site);
int gelu_from_pre(const void* pre, void* out, long long n, cudaStream_t stream);
int colsum_bf16(const void* x, long long ld, float* out, int T, int N, cudaStream_t stream);

int attn_fwd_lse(const void* qkv, const int32_t* attn_mask, void* ctx, float* lse, int nseq, int S, int heads,
                 float dropout_p, unsigned long long site_seed, cudaStream_t stream);
int attn_bwd_lse(const void* qkv, const int32_t* attn_mask, const void* ctx, const float* lse, const void* dctx,
                 void* dqkv, float* dbias, int nseq, int S, int heads, float dropout_p,
                 unsigned long long site_seed, cudaStream_t stream);

int attn_fwd_tc(const void* qkv, const int32_t* attn_mask, void* ctx, float* lse, int nseq, int S, int heads,
                float dropout_p, unsigned long long site_seed, cudaStream_t stream);
int attn_fwd_tc2(const void* qkv, const int32_t* attn_mask, void* ctx, float* lse, int nseq, int S, int heads,
                 float dropout_p, unsigned long long site_seed, cudaStream_t stream);
int attn_bwd_tc2(const void* qkv, const int32_t* attn_mask, const void* ctx, const float* lse, const void* dctx,
                 void* dqkv, int nseq, int S, int heads, float dropout_p, unsigned long long site_seed,
                 cudaStream_t stream);
int attn_bwd_tc(const void* qkv, const int32_t* attn_mask, const float* lse, const void* dctx, void* dqkv,
                float* dbias, int nseq, int S, int heads, float dropout_p, unsigned long long site_seed,
                cudaStream_t stream);

int attn_cls_fwd(const void* qkv, const int32_t* attn_mask, void* ctx_cls, float* probs, int nseq, int S, int heads,
                 float dropout_p, unsigned long long site_seed, cudaStream_t stream);
int attn_cls_bwd(const void* qkv, const float* probs, const void* dctx_cls, void* dqkv, int nseq, int S, int heads,
                 float dropout_p, unsigned long long site_seed, cudaStream_t stream);
int add_rows_bf16(void* dst, const void* src, int nrows, int H, long long stride_rows, cudaStream_t stream);

int score_ce_fwd(const float* q, const float* c, const uint8_t* col_mask, const uint8_t* pair_mask,
                 const int64_t* labels, float inv_t, float* lse, float* loss_sum, float* logits, int Q, int C, int d,
                 cudaStream_t stream);
int score_ce_bwd(const float* q, const float* c, const float* logits, const int64_t* labels, const float* lse,
                 float grad_scale, float inv_t, float* dq, float* dc, int Q, int C, int d, int q0, int nq, int c0,
                 int nc, cudaStream_t stream);

bool score_tc_supported(int Q, int C, int d);
long long score_tc_workspace_bytes(int Q, int C, int d, int nq, int nc);
int score_tc_fwd(const float* q, const float* c, const uint8_t* col_mask, const uint8_t* pair_mask,
                 const int64_t* labels, float inv_t, float* lse, float* loss_sum, float* logits, int Q, int C, int d,
                 int nq, int nc, void* workspace, long long workspace_bytes, cudaStream_t stream);
int score_tc_bwd(const uint8_t* col_mask, const uint8_t* pair_mask, const int64_t* labels, const float* lse,
                 float grad_scale, float inv_t, float* dq, float* dc, int Q, int C, int d, int q0, int nq, int c0, int nc,
                 void* workspace, long long workspace_bytes, cudaStream_t stream);

int sumsq_f32(const float* g, long long n, float* out, cudaStream_t stream);
int adamw_step(float* p, const float* g, float* m, float* v, void* shadow, long long n, float lr, float beta1,
               float beta2, float eps, float wd, int step, float grad_scale, const float* sumsq, float max_norm,
               cudaStream_t stream);
int cast_f32_bf16(const float* src, void* dst, long long n, cudaStream_t stream);
int cast_bf16_f32(const void* src, float* dst, long long n, cudaStream_t stream);

long long search_workspace_bytes(long long Q, int k);
int search_topk(const void* queries, const void* corpus, int dtype, long long Q, long long N, int d, int k,
                long long index_offset, float* out_scores, long long* out_index, void* workspace,
                long long workspace_bytes, cudaStream_t stream);
long long topk_merge_workspace_bytes(long long Q, int total);
int topk_merge(const float* scores, const long long* index, long long Q, int total, int k, float* out_scores,
               long long* out_index, void* workspace, long long workspace_bytes, cudaStream_t stream);

long long encoder_workspace_bytes(const dprb_encoder_weights* w, int nseq, int S, int save);
int encoder_fwd(const dprb_encoder_weights* w, const dprb_encoder_batch* b, float* pooled, cudaStream_t stream);
int encoder_bwd(const dprb_encoder_weights* w, const dprb_encoder_batch* b, const float* dpooled, int layer_lo,
                int layer_hi, cudaStream_t stream);

}  // namespace dprb

// C-ABI surface of libdprb.so (declared in include/dprb.h): thin extern "C" shims over the C++
// launchers, plus the thread-local error string.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include "common.cuh"
#include "dprb_internal.h"

namespace dprb {

static thread_local char g_err[1024] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int num_sms() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    cached = n;
  }
  return cached;
}

}  // namespace dprb

using namespace dprb;
#define S(x) reinterpret_cast<cudaStream_t>(x)

extern "C" {

int dprb_version(void) { return DPRB_VERSION; }
const char* dprb_last_error(void) { return g_err; }
int dprb_num_sms(void) { return num_sms(); }
int64_t dprb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int dprb_gemm_bf16(const void* A, const void* B, void* D, int M, int N, int K, int64_t lda, int64_t ldb,
                   int64_t ldd, int a_mn_major, int b_mn_major, int epilogue, const float* bias, const void* aux,
                   int64_t ld_aux, void* out2, float alpha, int splits, float* colsum, float dropout_p,
                   uint64_t dropout_site_seed, dprb_stream_t stream) {
  return gemm_bf16(A, B, D, M, N, K, lda, ldb, ldd, a_mn_major, b_mn_major, epilogue, bias, aux, ld_aux, out2,
                   alpha, splits, colsum, dropout_p, dropout_site_seed, S(stream));
}

int dprb_gemm_profile_enable(int enable, int max_launches) { return gemm_profile_enable(enable, max_launches); }
int dprb_gemm_profile_read(double* total_ms, double* total_flops, int64_t* launches) {
  long long n = 0;
  int rc = gemm_profile_read(total_ms, total_flops, &n);
  if (launches) *launches = n;
  return rc;
}

int dprb_embed_ln_fwd(const int64_t* ids, const int64_t* type_ids, const int64_t* pos_ids, const float* word,
                      const float* pos, const float* type, const float* gamma, const float* beta, void* y,
                      float* stats, int T, int H, int vocab, int max_pos, int type_vocab, float eps,
                      float dropout_p, uint64_t dropout_seed, void* y_res, dprb_stream_t stream) {
  return embed_ln_fwd(ids, type_ids, pos_ids, word, pos, type, gamma, beta, y, stats, T, H, vocab, max_pos,
                      type_vocab, eps, dropout_p, dropout_seed, y_res, S(stream));
}
int dprb_embed_ln_bwd(const void* dy, const int64_t* ids, const int64_t* type_ids, const int64_t* pos_ids,
                      const float* word, const float* pos, const float* type, const float* gamma,
                      const float* stats, float* dword, float* dpos, float* dtype, float* dgamma, float* dbeta,
                      int T, int H, float dropout_p, uint64_t dropout_seed, dprb_stream_t stream) {
  return embed_ln_bwd(dy, ids, type_ids, pos_ids, word, pos, type, gamma, stats, dword, dpos, dtype, dgamma,
                      dbeta, T, H, dropout_p, dropout_seed, S(stream));
}
int dprb_ln_fwd(const void* z, const float* gamma, const float* beta, void* y, float* stats, float* cls_out,
                int cls_stride, int T, int H, float eps, int z_f16, void* y_res, dprb_stream_t stream) {
  return ln_fwd(z, gamma, beta, y, stats, cls_out, cls_stride, T, H, eps, z_f16, y_res, S(stream));
}
int dprb_ln_bwd(const void* dy, const float* dy_cls, int cls_stride, const void* z, const float* stats,
                const float* gamma, void* dz, float* dgamma, float* dbeta, float* dbias, int T, int H, void* dzm,
                float dropout_p, uint64_t dropout_site_seed, int z_f16, dprb_stream_t stream) {
  return ln_bwd(dy, dy_cls, cls_stride, z, stats, gamma, dz, dgamma, dbeta, dbias, T, H, dzm, dropout_p,
                dropout_site_seed, z_f16, S(stream));
}
uint64_t dprb_dropout_site_seed(uint64_t dropout_seed, int layer, int site) {
  return drop_site_seed64(dropout_seed, layer, site);
}
int dprb_dropout_mask(uint8_t* keep, int64_t rows, int cols, float dropout_p, uint64_t dropout_seed, int layer,
                      int site, dprb_stream_t stream) {
  return dropout_mask(keep, rows, cols, dropout_p, dropout_seed, layer, site, S(stream));
}
int dprb_gelu_from_pre(const void* pre, void* out, int64_t n, dprb_stream_t stream) {
  return gelu_from_pre(pre, out, n, S(stream));
}
int dprb_colsum_bf16(const void* x, int64_t ld, float* out, int T, int N, dprb_stream_t stream) {
  return colsum_bf16(x, ld, out, T, N, S(stream));
}
int dprb_attn_fwd(const void* qkv, const int32_t* attn_mask, void* ctx, float* lse, int nseq, int Sq, int heads,
                  float dropout_p, uint64_t dropout_site_seed, dprb_stream_t stream) {
  return attn_fwd_lse(qkv, attn_mask, ctx, lse, nseq, Sq, heads, dropout_p, dropout_site_seed, S(stream));
}
int dprb_attn_bwd(const void* qkv, const int32_t* attn_mask, const void* ctx, const float* lse, const void* dctx,
                  void* dqkv, float* dbias, int nseq, int Sq, int heads, float dropout_p,
                  uint64_t dropout_site_seed, dprb_stream_t stream) {
  return attn_bwd_lse(qkv, attn_mask, ctx, lse, dctx, dqkv, dbias, nseq, Sq, heads, dropout_p, dropout_site_seed,
                      S(stream));
}
int dprb_score_ce_fwd(const float* q, const float* c, const uint8_t* col_mask, const uint8_t* pair_mask,
                      const int64_t* labels, float inv_temperature, float* lse, float* loss_sum, float* logits,
                      int Q, int C, int d, dprb_stream_t stream) {
  return score_ce_fwd(q, c, col_mask, pair_mask, labels, inv_temperature, lse, loss_sum, logits, Q, C, d, S(stream));
}
int dprb_score_ce_bwd(const float* q, const float* c, const float* logits, const int64_t* labels,
                      const float* lse, float grad_scale, float inv_temperature, float* dq, float* dc, int Q,
                      int C, int d, int q0, int nq, int c0, int nc, dprb_stream_t stream) {
  return score_ce_bwd(q, c, logits, labels, lse, grad_scale, inv_temperature, dq, dc, Q, C, d, q0, nq, c0, nc,
                      S(stream));
}
int dprb_score_tc_supported(int Q, int C, int d) { return score_tc_supported(Q, C, d) ? 1 : 0; }
int64_t dprb_score_tc_workspace_bytes(int Q, int C, int d, int nq, int nc) {
  return score_tc_workspace_bytes(Q, C, d, nq, nc);
}
int dprb_score_tc_fwd(const float* q, const float* c, const uint8_t* col_mask, const uint8_t* pair_mask,
                      const int64_t* labels, float inv_temperature, float* lse, float* loss_sum, float* logits, int Q,
                      int C, int d, int nq, int nc, void* workspace, int64_t workspace_bytes, dprb_stream_t stream) {
  return score_tc_fwd(q, c, col_mask, pair_mask, labels, inv_temperature, lse, loss_sum, logits, Q, C, d, nq, nc,
                      workspace, workspace_bytes, S(stream));
}
int dprb_score_tc_bwd(const uint8_t* col_mask, const uint8_t* pair_mask, const int64_t* labels, const float* lse,
                      float grad_scale, float inv_temperature, float* dq, float* dc, int Q, int C, int d, int q0, int nq,
                      int c0, int nc, void* workspace, int64_t workspace_bytes, dprb_stream_t stream) {
  return score_tc_bwd(col_mask, pair_mask, labels, lse, grad_scale, inv_temperature, dq, dc, Q, C, d, q0, nq, c0, nc,
                      workspace, workspace_bytes, S(stream));
}
int dprb_sumsq_f32(const float* g, int64_t n, float* out, dprb_stream_t stream) {
  return sumsq_f32(g, n, out, S(stream));
}
int dprb_adamw_step(float* p, const float* g, float* m, float* v, void* shadow, int64_t n, float lr, float beta1,
                    float beta2, float eps, float weight_decay, int step, float grad_scale, const float* sumsq,
                    float max_norm, dprb_stream_t stream) {
  return adamw_step(p, g, m, v, shadow, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale, sumsq,
                    max_norm, S(stream));
}
int dprb_cast_f32_bf16(const float* src, void* dst, int64_t n, dprb_stream_t stream) {
  return cast_f32_bf16(src, dst, n, S(stream));
}
int dprb_cast_bf16_f32(const void* src, float* dst, int64_t n, dprb_stream_t stream) {
  return cast_bf16_f32(src, dst, n, S(stream));
}
int64_t dprb_encoder_workspace_bytes(const dprb_encoder_weights* w, int nseq, int Sq, int save) {
  return encoder_workspace_bytes(w, nseq, Sq, save);
}
int dprb_encoder_fwd(const dprb_encoder_weights* w, const dprb_encoder_batch* b, float* pooled,
                     dprb_stream_t stream) {
  return encoder_fwd(w, b, pooled, S(stream));
}
int dprb_encoder_bwd(const dprb_encoder_weights* w, const dprb_encoder_batch* b, const float* dpooled,
                     int layer_lo, int layer_hi, dprb_stream_t stream) {
  return encoder_bwd(w, b, dpooled, layer_lo, layer_hi, S(stream));
}
int64_t dprb_search_workspace_bytes(int64_t Q, int k) { return search_workspace_bytes(Q, k); }
int dprb_search_topk(const void* queries, const void* corpus, int dtype, int64_t Q, int64_t N, int d, int k,
                     int64_t index_offset, float* out_scores, int64_t* out_index, void* workspace,
                     int64_t workspace_bytes, dprb_stream_t stream) {
  return search_topk(queries, corpus, dtype, Q, N, d, k, index_offset, out_scores,
                     reinterpret_cast<long long*>(out_index), workspace, workspace_bytes, S(stream));
}
int64_t dprb_topk_merge_workspace_bytes(int64_t Q, int total) { return topk_merge_workspace_bytes(Q, total); }
int dprb_topk_merge(const float* scores, const int64_t* index, int64_t Q, int total, int k, float* out_scores,
                    int64_t* out_index, void* workspace, int64_t workspace_bytes, dprb_stream_t stream) {
  return topk_merge(scores, reinterpret_cast<const long long*>(index), Q, total, k, out_scores,
                    reinterpret_cast<long long*>(out_index), workspace, workspace_bytes, S(stream));
}

}  // extern "C"

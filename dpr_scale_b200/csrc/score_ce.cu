// Fused in-batch-negative scoring + softmax cross-entropy (forward: one pass over the similarity
// tiles with an online row log-sum-exp and the NLL pick; backward: dq for the rank-local query rows
// and dc for the rank-local context columns from the stored logits).
//
// Replaces /root/reference/dpr_scale/task/dpr_task.py:98-105 (sim_score), :197 (mask.repeat),
// :211 (scores /= temperature), :212 (nn.CrossEntropyLoss) and, for the multi-rank case, the gradient
// flow implied by :163-195 (remote slices are detached; only local rows/columns get gradients).
//
// All arithmetic is fp32 (the reference under AMP does this product in fp16 — SURVEY §8a7); the work
// is 2*Q*C*d FLOP = 0.2 GFLOP (cfg 2) .. 12.9 GFLOP (cfg 3), latency-bound, so it stays on the FFMA pipe.
#include "common.cuh"
#include "dprb_internal.h"

namespace dprb {
namespace {

constexpr int QB = 8;        // query rows per CTA
constexpr int CW = 4;        // columns per warp step
constexpr int FWD_WARPS = 8;

__global__ void __launch_bounds__(FWD_WARPS * 32)
score_ce_fwd_kernel(const float* __restrict__ q, const float* __restrict__ c, const uint8_t* __restrict__ col_mask,
                    const uint8_t* __restrict__ pair_mask, const int64_t* __restrict__ labels, float inv_t, float* __restrict__ lse_out,
                    float* __restrict__ loss_sum, float* __restrict__ logits, int Q, int C, int d) {
  extern __shared__ float sm[];
  float* qs = sm;                       // [QB][d]
  float* red = sm + QB * d;             // [FWD_WARPS][QB][3]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * QB;
  for (int i = threadIdx.x; i < QB * d; i += blockDim.x) {
    const int r = i / d, k = i - r * d;
    qs[i] = (row0 + r < Q) ? q[(long long)(row0 + r) * d + k] : 0.f;
  }
  __syncthreads();
  long long lab[QB];
#pragma unroll
  for (int r = 0; r < QB; ++r) lab[r] = (row0 + r < Q) ? labels[row0 + r] : -1;

  float m[QB], l[QB], pick[QB];
#pragma unroll
  for (int r = 0; r < QB; ++r) { m[r] = -INFINITY; l[r] = 0.f; pick[r] = 0.f; }

  for (int cb = warp * CW; cb < C; cb += FWD_WARPS * CW) {
    float acc[QB][CW];
#pragma unroll
    for (int r = 0; r < QB; ++r)
#pragma unroll
      for (int j = 0; j < CW; ++j) acc[r][j] = 0.f;
    for (int k = lane; k < d; k += 32) {
      float cv[CW];
#pragma unroll
      for (int j = 0; j < CW; ++j) cv[j] = (cb + j < C) ? __ldg(c + (long long)(cb + j) * d + k) : 0.f;
#pragma unroll
      for (int r = 0; r < QB; ++r) {
        const float qv = qs[r * d + k];
#pragma unroll
        for (int j = 0; j < CW; ++j) acc[r][j] = fmaf(qv, cv[j], acc[r][j]);
      }
    }
#pragma unroll
    for (int r = 0; r < QB; ++r)
#pragma unroll
      for (int j = 0; j < CW; ++j) acc[r][j] = warp_sum(acc[r][j]);
#pragma unroll
    for (int j = 0; j < CW; ++j) {
      const int col = cb + j;
      if (col < C) {
        const bool masked = col_mask != nullptr && col_mask[col] != 0;
#pragma unroll
        for (int r = 0; r < QB; ++r) {
          const bool pm = pair_mask != nullptr && row0 + r < Q && pair_mask[(long long)(row0 + r) * C + col] != 0;
          const float s = (masked || pm) ? -INFINITY : acc[r][j] * inv_t;
          if (logits != nullptr && lane == ((r * CW + j) & 31) && row0 + r < Q)
            logits[(long long)(row0 + r) * C + col] = s;
          if (s > m[r]) { l[r] = l[r] * __expf(m[r] - s) + 1.f; m[r] = s; }
          else if (s != -INFINITY) l[r] += __expf(s - m[r]);
          if ((long long)col == lab[r]) pick[r] = s;
        }
      }
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < QB; ++r) {
      red[(warp * QB + r) * 3 + 0] = m[r];
      red[(warp * QB + r) * 3 + 1] = l[r];
      red[(warp * QB + r) * 3 + 2] = pick[r];
    }
  }
  __syncthreads();
  if (warp == 0) {
    float loss = 0.f;
    if (lane < QB && row0 + lane < Q) {
      float M = -INFINITY;
      for (int w = 0; w < FWD_WARPS; ++w) M = fmaxf(M, red[(w * QB + lane) * 3]);
      float L = 0.f, P = 0.f;
      for (int w = 0; w < FWD_WARPS; ++w) {
        const float mw = red[(w * QB + lane) * 3], lw = red[(w * QB + lane) * 3 + 1];
        if (mw != -INFINITY) L += lw * __expf(mw - M);
        P += red[(w * QB + lane) * 3 + 2];  // exactly one warp saw the label column (others hold 0)
      }
      const float lse = M + logf(L);
      lse_out[row0 + lane] = lse;
      loss = lse - P;
    }
    loss = warp_sum(loss);
    if (lane == 0 && loss_sum != nullptr) atomicAdd(loss_sum, loss);
  }
}

// out[a, k] (+)= sum_b W(a, b) * X[b, k]   with W derived from the stored logits:
//   W = (exp(logit[r,c] - lse[r]) - [c == label[r]]) * scale
// MODE 0 (dq): a = query row r in [a0, a0+na), b = all columns c, X = c matrix.
// MODE 1 (dc): a = column c in [a0, a0+na), b = all query rows r, X = q matrix.
constexpr int AB = 8;    // output rows per CTA
constexpr int BT = 32;   // reduction chunk
template <int MODE>
__global__ void __launch_bounds__(256)
score_ce_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ lse,
                    const int64_t* __restrict__ labels, const float* __restrict__ X, float scale,
                    float* __restrict__ out, int Q, int C, int d, int a0, int na) {
  __shared__ float W[BT][AB + 1];
  const int ab = blockIdx.x * AB;           // first local output row
  const int k = blockIdx.y * 256 + threadIdx.x;
  const int nb = MODE == 0 ? C : Q;
  float acc[AB];
#pragma unroll
  for (int i = 0; i < AB; ++i) acc[i] = 0.f;
  for (int b0 = 0; b0 < nb; b0 += BT) {
    {
      // 256 threads fill the BT x AB weight tile
      int bi, ai;
      if (MODE == 0) { ai = threadIdx.x >> 5; bi = threadIdx.x & 31; }  // rows: a, contiguous: b (= column)
      else { bi = threadIdx.x >> 3; ai = threadIdx.x & 7; }             // rows: b (= query row), contiguous: a (= column)
      const int a = a0 + ab + ai, b = b0 + bi;
      float w = 0.f;
      if (ab + ai < na && b < nb) {
        const int r = MODE == 0 ? a : b, cc = MODE == 0 ? b : a;
        const float lg = logits[(long long)r * C + cc];
        w = (lg == -INFINITY) ? 0.f : __expf(lg - lse[r]);
        if ((long long)cc == labels[r]) w -= 1.f;
        w *= scale;
      }
      W[bi][ai] = w;
    }
    __syncthreads();
    if (k < d) {
#pragma unroll 8
      for (int bi = 0; bi < BT; ++bi) {
        const int b = b0 + bi;
        if (b < nb) {
          const float x = __ldg(X + (long long)b * d + k);
#pragma unroll
          for (int i = 0; i < AB; ++i) acc[i] = fmaf(W[bi][i], x, acc[i]);
        }
      }
    }
    __syncthreads();
  }
  if (k < d) {
#pragma unroll
    for (int i = 0; i < AB; ++i)
      if (ab + i < na) out[(long long)(ab + i) * d + k] = acc[i];
  }
}

}  // namespace

int score_ce_fwd(const float* q, const float* c, const uint8_t* col_mask, const uint8_t* pair_mask,
                 const int64_t* labels, float inv_t, float* lse, float* loss_sum, float* logits, int Q, int C, int d,
                 cudaStream_t stream) {
  DPRB_REQUIRE(Q >= 0 && C > 0 && d > 0, "score_ce_fwd: bad shape Q=%d C=%d d=%d", Q, C, d);
  DPRB_REQUIRE(lse != nullptr, "score_ce_fwd: lse output required");
  if (Q == 0) return 0;
  const size_t smem = (size_t)(QB * d + FWD_WARPS * QB * 3) * sizeof(float);
  DPRB_REQUIRE(smem <= 200 * 1024, "score_ce_fwd: embedding dim %d too large", d);
  static bool attr = false;
  if (!attr) {
    DPRB_CHECK_CUDA(cudaFuncSetAttribute(score_ce_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  score_ce_fwd_kernel<<<(Q + QB - 1) / QB, FWD_WARPS * 32, smem, stream>>>(q, c, col_mask, pair_mask, labels, inv_t, lse,
                                                                           loss_sum, logits, Q, C, d);
  DPRB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int score_ce_bwd(const float* q, const float* c, const float* logits, const int64_t* labels, const float* lse,
                 float grad_scale, float inv_t, float* dq, float* dc, int Q, int C, int d, int q0, int nq, int c0,
                 int nc, cudaStream_t stream) {
  DPRB_REQUIRE(Q > 0 && C > 0 && d > 0, "score_ce_bwd: bad shape Q=%d C=%d d=%d", Q, C, d);
  DPRB_REQUIRE(q0 >= 0 && nq >= 0 && q0 + nq <= Q && c0 >= 0 && nc >= 0 && c0 + nc <= C,
               "score_ce_bwd: local ranges out of bounds (q0=%d nq=%d c0=%d nc=%d)", q0, nq, c0, nc);
  DPRB_REQUIRE(logits != nullptr && lse != nullptr, "score_ce_bwd: logits and lse from forward required");
  const float scale = grad_scale * inv_t / (float)Q;  // d(mean CE)/d(logit) * d(logit)/d(q.c)
  if (nq > 0 && dq != nullptr) {
    dim3 grid((nq + AB - 1) / AB, (d + 255) / 256);
    score_ce_bwd_kernel<0><<<grid, 256, 0, stream>>>(logits, lse, labels, c, scale, dq, Q, C, d, q0, nq);
    DPRB_CHECK_CUDA(cudaGetLastError());
  }
  if (nc > 0 && dc != nullptr) {
    dim3 grid((nc + AB - 1) / AB, (d + 255) / 256);
    score_ce_bwd_kernel<1><<<grid, 256, 0, stream>>>(logits, lse, labels, q, scale, dc, Q, C, d, c0, nc);
    DPRB_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

}  // namespace dprb

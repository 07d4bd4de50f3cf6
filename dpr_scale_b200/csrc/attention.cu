// Self-attention core (head_dim 64, S <= 256) forward + backward.
//
// Replaces BertSelfAttention.forward's scaled_dot_product_attention
// (site-packages/transformers/models/bert/modeling_bert.py:168-207 and
// integrations/sdpa_attention.py:92-101) and its autograd backward, as reached from
// /root/reference/dpr_scale/models/hf_model.py:38.
//
// Attention is 2.7 % of the encoder FLOPs at S=128 (4*S*H of 14.16 M+0.39 M per token-layer), so this
// round-1 kernel keeps whole (sequence, head) problems in shared memory and uses warp-level
// mma.sync.m16n8k16 bf16 with register-resident softmax (flash-style, exact because it is the same
// online-softmax recurrence); the projection GEMMs around it are the tcgen05 kernel.
//
// Layout: qkv bf16 [nseq*S, 3H], row t = (seq, s); Q at column h*64, K at H + h*64, V at 2H + h*64.
#include <cstdlib>
#include "common.cuh"
#include "dprb_internal.h"

namespace dprb {
namespace {

constexpr int DH = 64;
constexpr float SCALE_LOG2 = 0.125f * 1.4426950408889634f;  // (1/sqrt(64)) * log2(e)
constexpr float LN2 = 0.6931471805599453f;

__device__ __forceinline__ uint32_t swz(int row, int chunk) {  // byte offset inside a [rows][64] bf16 tile
  return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4));
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// A fragment (16 rows x 16 k) of a row-major [rows][64] tile: rows r0.., k columns kk*16..
__device__ __forceinline__ void load_a(uint32_t tile, int r0, int kk, uint32_t (&a)[4]) {
  const int l = threadIdx.x & 31;
  ldsm_x4(tile + swz(r0 + (l & 7) + ((l >> 3) & 1) * 8, kk * 2 + (l >> 4)), a[0], a[1], a[2], a[3]);
}
// B fragments for two n-blocks (n0..n0+15) from a tile stored [n][k] (k contiguous), k columns kk*16..
__device__ __forceinline__ void load_b_nk(uint32_t tile, int n0, int kk, uint32_t (&b)[4]) {
  const int l = threadIdx.x & 31;
  ldsm_x4(tile + swz(n0 + (l & 7) + (l >> 4) * 8, kk * 2 + ((l >> 3) & 1)), b[0], b[1], b[2], b[3]);
}
// B fragments for two n-blocks (n columns n0..n0+15) from a tile stored [k][n] (n contiguous), k rows k0..k0+15
__device__ __forceinline__ void load_b_kn(uint32_t tile, int k0, int n0, uint32_t (&b)[4]) {
  const int l = threadIdx.x & 31;
  ldsm_x4_t(tile + swz(k0 + (l & 7) + ((l >> 3) & 1) * 8, (n0 >> 3) + (l >> 4)), b[0], b[1], b[2], b[3]);
}

// Cooperative load of a [rows_pad][64] bf16 tile (swizzled) from a strided global matrix; rows >= rows_valid -> 0.
__device__ __forceinline__ void load_tile(uint8_t* tile, const bf16* __restrict__ g, long long ld, int rows_valid,
                                          int rows_pad, int tid, int nthreads) {
  for (int idx = tid; idx < rows_pad * 8; idx += nthreads) {
    const int r = idx >> 3, ch = idx & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < rows_valid) v = *reinterpret_cast<const uint4*>(g + (long long)r * ld + ch * 8);
    *reinterpret_cast<uint4*>(tile + swz(r, ch)) = v;
  }
}

constexpr int STG_STRIDE = 144;  // bytes per staged row: conflict-free for the (g,t) fragment pattern

// Write a 16x64 fp32 accumulator block (8 n-blocks) as bf16 rows to global through a per-warp staging buffer.
__device__ __forceinline__ void store_rows_16x64(uint8_t* stage, const float (&acc)[8][4], float s0, float s1,
                                                 bf16* __restrict__ g, long long ld, int row0, int rows_valid) {
  const int l = threadIdx.x & 31, gq = l >> 2, t = l & 3;
  __syncwarp();
#pragma unroll
  for (int nb = 0; nb < 8; ++nb) {
    *reinterpret_cast<uint32_t*>(stage + gq * STG_STRIDE + (nb * 8 + 2 * t) * 2) = pack_bf16x2(acc[nb][0] * s0, acc[nb][1] * s0);
    *reinterpret_cast<uint32_t*>(stage + (gq + 8) * STG_STRIDE + (nb * 8 + 2 * t) * 2) = pack_bf16x2(acc[nb][2] * s1, acc[nb][3] * s1);
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = l + 32 * i, r = idx >> 3, ch = idx & 7;
    if (row0 + r < rows_valid)
      *reinterpret_cast<uint4*>(g + (long long)(row0 + r) * ld + ch * 8) =
          *reinterpret_cast<const uint4*>(stage + r * STG_STRIDE + ch * 16);
  }
  __syncwarp();
}

// ------------------------------------------------------------------ forward
// grid (heads, nseq), 256 threads: Q, K, V of one (sequence, head) are read from HBM exactly once;
// each warp owns 16-row query blocks (rb = warp, warp + 8, ...).
__global__ void __launch_bounds__(256)
attn_fwd_kernel(const bf16* __restrict__ qkv, const int32_t* __restrict__ attn_mask, bf16* __restrict__ ctx,
                float* __restrict__ lse_out, int S, int S_pad, int heads) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int H = heads * DH;
  const int h = blockIdx.x, seq = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, t = lane & 3;
  uint8_t* sQ = smem;                       // [S_pad][64]
  uint8_t* sK = sQ + S_pad * 128;           // [S_pad][64]
  uint8_t* sV = sK + S_pad * 128;           // [S_pad][64]
  float* sMask = reinterpret_cast<float*>(sV + S_pad * 128);  // [S_pad] additive (0 / -inf)
  uint8_t* sStage = reinterpret_cast<uint8_t*>(sMask + S_pad);  // [8][16*STG_STRIDE]

  const bf16* base = qkv + (long long)seq * S * (3 * H) + h * DH;
  load_tile(sQ, base, 3 * H, S, S_pad, threadIdx.x, 256);
  load_tile(sK, base + H, 3 * H, S, S_pad, threadIdx.x, 256);
  load_tile(sV, base + 2 * H, 3 * H, S, S_pad, threadIdx.x, 256);
  for (int j = threadIdx.x; j < S_pad; j += 256) {
    bool keep = j < S && (attn_mask == nullptr || attn_mask[(long long)seq * S + j] != 0);
    sMask[j] = keep ? 0.f : -INFINITY;
  }
  __syncthreads();

  const uint32_t tQ = smem_u32(sQ), tK = smem_u32(sK), tV = smem_u32(sV);
  const int nkvb = S_pad / 64, nrb = S_pad / 16;
  for (int rb = warp; rb < nrb; rb += 8) {
    if (rb * 16 >= S) break;
    uint32_t aq[4][4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) load_a(tQ, rb * 16, kk, aq[kk]);

    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }

    for (int kvb = 0; kvb < nkvb; ++kvb) {
      float s[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int nb2 = 0; nb2 < 4; ++nb2) {
          uint32_t b[4];
          load_b_nk(tK, kvb * 64 + nb2 * 16, kk, b);
          mma16816(s[2 * nb2], aq[kk], b[0], b[1]);
          mma16816(s[2 * nb2 + 1], aq[kk], b[2], b[3]);
        }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        const float2 mk = *reinterpret_cast<const float2*>(sMask + kvb * 64 + nb * 8 + 2 * t);
        s[nb][0] = fmaf(s[nb][0], SCALE_LOG2, mk.x); s[nb][1] = fmaf(s[nb][1], SCALE_LOG2, mk.y);
        s[nb][2] = fmaf(s[nb][2], SCALE_LOG2, mk.x); s[nb][3] = fmaf(s[nb][3], SCALE_LOG2, mk.y);
        mx0 = fmaxf(mx0, fmaxf(s[nb][0], s[nb][1]));
        mx1 = fmaxf(mx1, fmaxf(s[nb][2], s[nb][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
      const float e0 = (mn0 == -INFINITY) ? 0.f : mn0, e1 = (mn1 == -INFINITY) ? 0.f : mn1;  // all-masked guard
      const float al0 = exp2f(m0 - e0), al1 = exp2f(m1 - e1);
      m0 = mn0; m1 = mn1;
      float rs0 = 0.f, rs1 = 0.f;
      uint32_t pa[4][4];
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        const float p0 = exp2f(s[nb][0] - e0), p1 = exp2f(s[nb][1] - e0);
        const float p2 = exp2f(s[nb][2] - e1), p3 = exp2f(s[nb][3] - e1);
        rs0 += p0 + p1; rs1 += p2 + p3;
        pa[nb >> 1][(nb & 1) * 2 + 0] = pack_bf16x2(p0, p1);
        pa[nb >> 1][(nb & 1) * 2 + 1] = pack_bf16x2(p2, p3);
      }
      l0 = l0 * al0 + rs0; l1 = l1 * al1 + rs1;
#pragma unroll
      for (int i = 0; i < 8; ++i) { o[i][0] *= al0; o[i][1] *= al0; o[i][2] *= al1; o[i][3] *= al1; }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int db2 = 0; db2 < 4; ++db2) {
          uint32_t b[4];
          load_b_kn(tV, kvb * 64 + kk * 16, db2 * 16, b);
          mma16816(o[2 * db2], pa[kk], b[0], b[1]);
          mma16816(o[2 * db2 + 1], pa[kk], b[2], b[3]);
        }
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = l0 > 0.f ? 1.f / l0 : 0.f, inv1 = l1 > 0.f ? 1.f / l1 : 0.f;
    const int row0 = rb * 16;
    store_rows_16x64(sStage + warp * 16 * STG_STRIDE, o, inv0, inv1,
                     ctx + (long long)seq * S * H + h * DH, H, row0, S);
    if (lse_out != nullptr && t == 0) {
      float* lp = lse_out + ((long long)seq * heads + h) * S;
      if (row0 + gq < S) lp[row0 + gq] = m0 * LN2 + logf(l0);
      if (row0 + gq + 8 < S) lp[row0 + gq + 8] = m1 * LN2 + logf(l1);
    }
  }
}

// ------------------------------------------------------------------ backward
// grid (heads, nseq), 256 threads. Phase 1: each warp owns 16-row query blocks -> dQ (and D_i).
// Phase 2: each warp owns 16-row key/value blocks -> dK, dV.  P is recomputed from the saved LSE.
__global__ void __launch_bounds__(256, 1)
attn_bwd_kernel(const bf16* __restrict__ qkv, const int32_t* __restrict__ attn_mask, const bf16* __restrict__ ctx,
                const float* __restrict__ lse_in, const bf16* __restrict__ dctx, bf16* __restrict__ dqkv, int S,
                int S_pad, int heads) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int H = heads * DH;
  const int h = blockIdx.x, seq = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, t = lane & 3;
  const int tile_bytes = S_pad * 128;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + tile_bytes;
  uint8_t* sV = sK + tile_bytes;
  uint8_t* sdO = sV + tile_bytes;
  float* sMask = reinterpret_cast<float*>(sdO + tile_bytes);  // [S_pad] additive, log2 domain irrelevant (0/-inf)
  float* sLse = sMask + S_pad;                                // [S_pad] natural-log LSE, scaled to log2 below
  float* sD = sLse + S_pad;                                   // [S_pad] D_i = sum_d dO*O
  uint8_t* sStage = reinterpret_cast<uint8_t*>(sD + S_pad);   // [8][16*STG_STRIDE]

  const long long tok0 = (long long)seq * S;
  const bf16* base = qkv + tok0 * (3 * H) + h * DH;
  load_tile(sQ, base, 3 * H, S, S_pad, threadIdx.x, 256);
  load_tile(sK, base + H, 3 * H, S, S_pad, threadIdx.x, 256);
  load_tile(sV, base + 2 * H, 3 * H, S, S_pad, threadIdx.x, 256);
  load_tile(sdO, dctx + tok0 * H + h * DH, H, S, S_pad, threadIdx.x, 256);
  for (int j = threadIdx.x; j < S_pad; j += 256) {
    bool keep = j < S && (attn_mask == nullptr || attn_mask[tok0 + j] != 0);
    sMask[j] = keep ? 0.f : -INFINITY;
    // rows beyond S get lse = +inf so that P = exp(s - lse) = 0 for them
    sLse[j] = j < S ? lse_in[((long long)seq * heads + h) * S + j] * 1.4426950408889634f : INFINITY;
  }
  // D_i = sum_d dO[i,d] * O[i,d]; 8 threads per row, 8 columns each
  for (int idx = threadIdx.x; idx < S_pad * 8; idx += 256) {
    const int r = idx >> 3, ch = idx & 7;
    float part = 0.f;
    if (r < S) {
      uint4 a = *reinterpret_cast<const uint4*>(dctx + (tok0 + r) * H + h * DH + ch * 8);
      uint4 b = *reinterpret_cast<const uint4*>(ctx + (tok0 + r) * H + h * DH + ch * 8);
      const uint32_t* pa = &a.x; const uint32_t* pb = &b.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float2 x = unpack_bf16x2(pa[k]), y = unpack_bf16x2(pb[k]);
        part += x.x * y.x + x.y * y.y;
      }
    }
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    part += __shfl_xor_sync(0xffffffffu, part, 4);
    if (ch == 0) sD[r] = part;
  }
  __syncthreads();

  const uint32_t tQ = smem_u32(sQ), tK = smem_u32(sK), tV = smem_u32(sV), tdO = smem_u32(sdO);
  const int nrb = S_pad / 16, nkvb = S_pad / 64;
  uint8_t* stage = sStage + warp * 16 * STG_STRIDE;
  bf16* dq_out = dqkv + tok0 * (3 * H) + h * DH;

  // ---------------- phase 1: dQ ----------------
  for (int rb = warp; rb < nrb; rb += 8) {
    if (rb * 16 >= S) break;
    uint32_t aq[4][4], ado[4][4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) { load_a(tQ, rb * 16, kk, aq[kk]); load_a(tdO, rb * 16, kk, ado[kk]); }
    const float lse0 = sLse[rb * 16 + gq], lse1 = sLse[rb * 16 + gq + 8];
    const float D0 = sD[rb * 16 + gq], D1 = sD[rb * 16 + gq + 8];
    float dq[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f; }
    for (int kvb = 0; kvb < nkvb; ++kvb) {
      float s[8][4], dp[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int nb2 = 0; nb2 < 4; ++nb2) {
          uint32_t b[4];
          load_b_nk(tK, kvb * 64 + nb2 * 16, kk, b);
          mma16816(s[2 * nb2], aq[kk], b[0], b[1]);
          mma16816(s[2 * nb2 + 1], aq[kk], b[2], b[3]);
          load_b_nk(tV, kvb * 64 + nb2 * 16, kk, b);
          mma16816(dp[2 * nb2], ado[kk], b[0], b[1]);
          mma16816(dp[2 * nb2 + 1], ado[kk], b[2], b[3]);
        }
      }
      uint32_t ads[4][4];
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        const float2 mk = *reinterpret_cast<const float2*>(sMask + kvb * 64 + nb * 8 + 2 * t);
        const float p0 = exp2f(fmaf(s[nb][0], SCALE_LOG2, mk.x) - lse0), p1 = exp2f(fmaf(s[nb][1], SCALE_LOG2, mk.y) - lse0);
        const float p2 = exp2f(fmaf(s[nb][2], SCALE_LOG2, mk.x) - lse1), p3 = exp2f(fmaf(s[nb][3], SCALE_LOG2, mk.y) - lse1);
        ads[nb >> 1][(nb & 1) * 2 + 0] = pack_bf16x2(p0 * (dp[nb][0] - D0) * 0.125f, p1 * (dp[nb][1] - D0) * 0.125f);
        ads[nb >> 1][(nb & 1) * 2 + 1] = pack_bf16x2(p2 * (dp[nb][2] - D1) * 0.125f, p3 * (dp[nb][3] - D1) * 0.125f);
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int db2 = 0; db2 < 4; ++db2) {
          uint32_t b[4];
          load_b_kn(tK, kvb * 64 + kk * 16, db2 * 16, b);
          mma16816(dq[2 * db2], ads[kk], b[0], b[1]);
          mma16816(dq[2 * db2 + 1], ads[kk], b[2], b[3]);
        }
      }
    }
    store_rows_16x64(stage, dq, 1.f, 1.f, dq_out, 3 * H, rb * 16, S);
  }

  // ---------------- phase 2: dK, dV ----------------
  for (int rb = warp; rb < nrb; rb += 8) {
    if (rb * 16 >= S) break;
    uint32_t ak[4][4], av[4][4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) { load_a(tK, rb * 16, kk, ak[kk]); load_a(tV, rb * 16, kk, av[kk]); }
    const float mk0 = sMask[rb * 16 + gq], mk1 = sMask[rb * 16 + gq + 8];
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f; dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f; }
    for (int qb = 0; qb < nkvb; ++qb) {
      float st[8][4], dpt[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) { st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f; dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int nb2 = 0; nb2 < 4; ++nb2) {
          uint32_t b[4];
          load_b_nk(tQ, qb * 64 + nb2 * 16, kk, b);   // S^T = K Q^T
          mma16816(st[2 * nb2], ak[kk], b[0], b[1]);
          mma16816(st[2 * nb2 + 1], ak[kk], b[2], b[3]);
          load_b_nk(tdO, qb * 64 + nb2 * 16, kk, b);  // dP^T = V dO^T
          mma16816(dpt[2 * nb2], av[kk], b[0], b[1]);
          mma16816(dpt[2 * nb2 + 1], av[kk], b[2], b[3]);
        }
      }
      uint32_t apt[4][4], adst[4][4];
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        const int i0 = qb * 64 + nb * 8 + 2 * t;
        const float2 ls = *reinterpret_cast<const float2*>(sLse + i0);
        const float2 dd = *reinterpret_cast<const float2*>(sD + i0);
        const float p0 = exp2f(fmaf(st[nb][0], SCALE_LOG2, mk0) - ls.x), p1 = exp2f(fmaf(st[nb][1], SCALE_LOG2, mk0) - ls.y);
        const float p2 = exp2f(fmaf(st[nb][2], SCALE_LOG2, mk1) - ls.x), p3 = exp2f(fmaf(st[nb][3], SCALE_LOG2, mk1) - ls.y);
        apt[nb >> 1][(nb & 1) * 2 + 0] = pack_bf16x2(p0, p1);
        apt[nb >> 1][(nb & 1) * 2 + 1] = pack_bf16x2(p2, p3);
        adst[nb >> 1][(nb & 1) * 2 + 0] = pack_bf16x2(p0 * (dpt[nb][0] - dd.x) * 0.125f, p1 * (dpt[nb][1] - dd.y) * 0.125f);
        adst[nb >> 1][(nb & 1) * 2 + 1] = pack_bf16x2(p2 * (dpt[nb][2] - dd.x) * 0.125f, p3 * (dpt[nb][3] - dd.y) * 0.125f);
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int db2 = 0; db2 < 4; ++db2) {
          uint32_t b[4];
          load_b_kn(tdO, qb * 64 + kk * 16, db2 * 16, b);  // dV += P^T dO
          mma16816(dv[2 * db2], apt[kk], b[0], b[1]);
          mma16816(dv[2 * db2 + 1], apt[kk], b[2], b[3]);
          load_b_kn(tQ, qb * 64 + kk * 16, db2 * 16, b);   // dK += dS^T Q
          mma16816(dk[2 * db2], adst[kk], b[0], b[1]);
          mma16816(dk[2 * db2 + 1], adst[kk], b[2], b[3]);
        }
      }
    }
    store_rows_16x64(stage, dk, 1.f, 1.f, dq_out + H, 3 * H, rb * 16, S);
    store_rows_16x64(stage, dv, 1.f, 1.f, dq_out + 2 * H, 3 * H, rb * 16, S);
  }
}

int check_shape(int nseq, int S, int heads, const char* who) {
  DPRB_REQUIRE(nseq >= 0 && heads > 0, "%s: bad nseq=%d heads=%d", who, nseq, heads);
  DPRB_REQUIRE(S >= 1 && S <= 256, "%s: sequence length %d unsupported (1..256)", who, S);
  return 0;
}

}  // namespace

int attn_fwd_lse(const void* qkv, const int32_t* attn_mask, void* ctx, float* lse, int nseq, int S, int heads,
                 float dropout_p, unsigned long long site_seed, cudaStream_t stream) {
  if (int rc = check_shape(nseq, S, heads, "attn_fwd")) return rc;
  if (nseq == 0) return 0;
  static const bool legacy = (std::getenv("DPRB_ATTN_LEGACY") != nullptr);
  if (S <= 128 && !legacy) return attn_fwd_tc(qkv, attn_mask, ctx, lse, nseq, S, heads, dropout_p, site_seed, stream);
  if (!legacy) return attn_fwd_tc2(qkv, attn_mask, ctx, lse, nseq, S, heads, dropout_p, site_seed, stream);  // 128 < S <= 256
  DPRB_REQUIRE(dropout_p == 0.f, "attn_fwd: attention dropout is implemented on the tcgen05 paths only");
  const int S_pad = (S + 63) / 64 * 64;
  const size_t smem = 3 * (size_t)S_pad * 128 + S_pad * 4 + 8 * 16 * STG_STRIDE;
  static bool attr = false;
  if (!attr) {
    DPRB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    attr = true;
  }
  dim3 grid(heads, nseq);
  attn_fwd_kernel<<<grid, 256, smem, stream>>>((const bf16*)qkv, attn_mask, (bf16*)ctx, lse, S, S_pad, heads);
  DPRB_LAUNCH_CHECK();
  return 0;
}

int attn_bwd_lse(const void* qkv, const int32_t* attn_mask, const void* ctx, const float* lse, const void* dctx,
                 void* dqkv, float* dbias, int nseq, int S, int heads, float dropout_p,
                 unsigned long long site_seed, cudaStream_t stream) {
  if (int rc = check_shape(nseq, S, heads, "attn_bwd")) return rc;
  if (nseq == 0) return 0;
  static const bool legacy = (std::getenv("DPRB_ATTN_LEGACY") != nullptr);
  if (S <= 128 && !legacy)
    return attn_bwd_tc(qkv, attn_mask, lse, dctx, dqkv, dbias, nseq, S, heads, dropout_p, site_seed, stream);
  if (!legacy) {  // 128 < S <= 256: tcgen05 kernel; the QKV bias gradient is a separate streaming pass
    if (int rc = attn_bwd_tc2(qkv, attn_mask, ctx, lse, dctx, dqkv, nseq, S, heads, dropout_p, site_seed, stream)) return rc;
    if (dbias != nullptr) return colsum_bf16(dqkv, 3LL * heads * DH, dbias, nseq * S, 3 * heads * DH, stream);
    return 0;
  }
  DPRB_REQUIRE(dropout_p == 0.f, "attn_bwd: attention dropout is implemented on the tcgen05 paths only");
  const int S_pad = (S + 63) / 64 * 64;
  const size_t smem = 4 * (size_t)S_pad * 128 + 3 * S_pad * 4 + 8 * 16 * STG_STRIDE;
  static bool attr = false;
  if (!attr) {
    DPRB_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr = true;
  }
  dim3 grid(heads, nseq);
  attn_bwd_kernel<<<grid, 256, smem, stream>>>((const bf16*)qkv, attn_mask, (const bf16*)ctx, lse, (const bf16*)dctx,
                                              (bf16*)dqkv, S, S_pad, heads);
  DPRB_LAUNCH_CHECK();
  // legacy (mma.sync) path: the QKV bias gradient is a separate streaming pass
  if (dbias != nullptr) return colsum_bf16(dqkv, 3LL * heads * DH, dbias, nseq * S, 3 * heads * DH, stream);
  return 0;
}

}  // namespace dprb

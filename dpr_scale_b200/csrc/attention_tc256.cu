// tcgen05 / TMEM / TMA self-attention for 128 < S <= 256 (head_dim 64): two 128-row query tiles x up to 256 keys per
// (sequence, head).  Same data flow as attention_tc.cu (S <= 128); what changes is the tiling:
//
//   forward : per query tile t: S_t = Q_t K^T is ONE UMMA chain with N = 256 (TMEM cols [0,256)), the softmax runs over
//             256 columns, P_t (bf16, 4 swizzled k-blocks) feeds O_t = P_t V (16 UMMAs of K = 16 over the 256 keys).
//   backward: key tile j outer, query tile i inner.  dK_j / dV_j accumulate in TMEM across the two query tiles
//             (UMMA accumulate flag), dQ_i is summed over the two key tiles in a bf16 shared-memory tile, and
//             D_i = rowsum(dO_i * O_i) is computed up front (with j outermost no thread ever sees a full row of P dP).
//
// Replaces the same reference code as attention_tc.cu:
// site-packages/transformers/models/bert/modeling_bert.py:168-207 (BertSelfAttention.forward, SDPA) and its backward.
#include "common.cuh"
#include "dprb_internal.h"

namespace dprb {
namespace {

constexpr int TILE_BYTES = 128 * 128;  // [128 rows][64 bf16], 128B-swizzled = 16 KB
constexpr float SCALE_LOG2 = 0.125f * 1.4426950408889634f;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t addr) { return make_umma_desc_sw128(addr, 0, 1024); }
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t addr, uint32_t lbo) { return make_umma_desc_sw128(addr, lbo, 1024); }
__device__ __forceinline__ void st_chunk(uint8_t* tile, int row, int chunk, const float* v, float s) {
  uint4 q;
  q.x = pack_bf16x2(v[0] * s, v[1] * s); q.y = pack_bf16x2(v[2] * s, v[3] * s);
  q.z = pack_bf16x2(v[4] * s, v[5] * s); q.w = pack_bf16x2(v[6] * s, v[7] * s);
  *reinterpret_cast<uint4*>(tile + row * 128 + ((chunk ^ (row & 7)) << 4)) = q;
}

// ------------------------------------------------------------------------------------------ forward (2 query tiles)
// smem: sQ[2] | sK[2] | sV[2] | sP[4] | mask[256] | barriers (~161 KB, 1 CTA / SM).  TMEM: S [0,256), O [256,320).
constexpr int F2_THREADS = 32 + 128;
constexpr int F2_SMEM = 10 * TILE_BYTES + 256 * 4 + 64 + 1024;

template <bool DROP>
__global__ void __launch_bounds__(F2_THREADS, 1)
attn_fwd_tc2_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_ctx,
                    const int32_t* __restrict__ attn_mask, float* __restrict__ lse_out, int S, int heads, int nseq,
                    Drop drop) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                    // 2 tiles (query rows 0-127, 128-255)
  uint8_t* sK = sQ + 2 * TILE_BYTES;     // 2 tiles = one K-major [256][64] operand
  uint8_t* sV = sK + 2 * TILE_BYTES;     // 2 tiles = one MN-major [256 keys][64] operand
  uint8_t* sP = sV + 2 * TILE_BYTES;     // 4 k-blocks of [128][64]
  float* sMask = reinterpret_cast<float*>(sP + 4 * TILE_BYTES);  // [256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sMask + 256);
  uint64_t *b_load = bars, *b_s = bars + 1, *b_p = bars + 2, *b_o = bars + 3, *b_free = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = heads * 64;
  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_qkv);
      tma_prefetch_desc(&tm_ctx);
      mbar_init(b_load, 1); mbar_init(b_s, 1); mbar_init(b_p, 4); mbar_init(b_o, 1); mbar_init(b_free, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tO = tmem + 256;
  const int nprob = nseq * heads;

  if (warp == 0) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16_f32(128, 256, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16_f32(128, 64, 0, 1);
      uint32_t ph = 0;   // per-problem phase (b_load, b_free)
      uint32_t sp = 0;   // per-query-tile phase (b_s, b_p, b_o)
      for (int prob = blockIdx.x; prob < nprob; prob += gridDim.x) {
        const int seq = prob / heads, h = prob - seq * heads;
        mbar_wait(b_free, ph ^ 1);
        mbar_arrive_expect_tx(b_load, 6 * TILE_BYTES);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          tma_load_3d(sQ + t * TILE_BYTES, &tm_qkv, b_load, h * 64, t * 128, seq);
          tma_load_3d(sK + t * TILE_BYTES, &tm_qkv, b_load, H + h * 64, t * 128, seq);
          tma_load_3d(sV + t * TILE_BYTES, &tm_qkv, b_load, 2 * H + h * 64, t * 128, seq);
        }
        mbar_wait(b_load, ph);
        tcgen05_fence_after();
        const uint64_t dk = desc_kmajor(smem_u32(sK));
        const uint64_t dv = desc_mnmajor(smem_u32(sV), 0);
        for (int qt = 0; qt < 2; ++qt) {
          const uint64_t dq = desc_kmajor(smem_u32(sQ + qt * TILE_BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tS, dq + 2 * k, dk + 2 * k, idesc_s, k > 0);
          umma_commit(b_s);
          mbar_wait(b_p, sp);
          tcgen05_fence_after();
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const uint64_t dp = desc_kmajor(smem_u32(sP + (k >> 2) * TILE_BYTES)) + 2 * (k & 3);
            umma_f16(tO, dp, dv + 128 * k, idesc_o, k > 0);
          }
          umma_commit(b_o);
          sp ^= 1;
        }
        ph ^= 1;
      }
    }
    __syncwarp();
  } else {
    const int tid = threadIdx.x - 32;            // 0..127
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;         // row inside the current query tile
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    uint32_t sp = 0;
    for (int prob = blockIdx.x; prob < nprob; prob += gridDim.x) {
      const int seq = prob / heads, h = prob - seq * heads;
      // the mask buffer is rewritten per problem: everybody must be done with the previous problem first
      named_bar_sync(1, 128);
      for (int j = tid; j < 256; j += 128) {
        const bool keep = j < S && (attn_mask == nullptr || attn_mask[(long long)seq * S + j] != 0);
        sMask[j] = keep ? 0.f : -INFINITY;
      }
      named_bar_sync(1, 128);
      for (int qt = 0; qt < 2; ++qt) {
        const int grow = qt * 128 + row;           // query row inside the sequence
        mbar_wait(b_s, sp);
        tcgen05_fence_after();
        float m = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(tS + lane_addr + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) m = fmaxf(m, fmaf(__uint_as_float(r[j]), SCALE_LOG2, sMask[c * 32 + j]));
        }
        const float e = (m == -INFINITY) ? 0.f : m;
        float l = 0.f;
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(tS + lane_addr + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            float p[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              const int j = c4 * 8 + t;
              p[t] = ex2_approx(fmaf(__uint_as_float(r[j]), SCALE_LOG2, sMask[c * 32 + j]) - e);
              l += p[t];
            }
            if (DROP) {
#pragma unroll
              for (int t = 0; t < 8; t += 8) {
                float2 m[4];
                drop.mul8((uint32_t)(prob * S + grow), (uint32_t)(c * 32 + c4 * 8), m);
#pragma unroll
                for (int w = 0; w < 4; ++w) { p[2 * w] *= m[w].x; p[2 * w + 1] *= m[w].y; }
              }
            }
            const int chunk = c * 4 + c4;  // 16-byte chunk index over the 256 key columns
            st_chunk(sP + (chunk >> 3) * TILE_BYTES, row, chunk & 7, p, 1.f);
          }
        }
        fence_proxy_async_smem();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(b_p);
        if (lse_out != nullptr && grow < S) lse_out[((long long)seq * heads + h) * S + grow] = m * LN2 + __logf(l);
        const float inv = l > 0.f ? 1.f / l : 0.f;
        mbar_wait(b_o, sp);
        tcgen05_fence_after();
        uint8_t* stage = sQ + qt * TILE_BYTES;   // this query tile is dead after its S MMA
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(tO + lane_addr + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            float v[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) v[t] = __uint_as_float(r[c4 * 8 + t]);
            st_chunk(stage, row, c * 4 + c4, v, inv);
          }
        }
        fence_proxy_async_smem();
        tcgen05_fence_before();
        named_bar_sync(2, 128);
        if (tid == 0) {
          tma_store_3d(&tm_ctx, smem_u32(stage), h * 64, qt * 128, seq);
          tma_store_commit();
          if (qt == 1) {
            tma_store_wait_read();
            mbar_arrive(b_free);
          }
        }
        sp ^= 1;
      }
    }
    if (tid == 0) tma_store_wait_all();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------ backward (2 x 2 tiles)
// 1 control warp + 8 compute warps (two per TMEM lane quarter, splitting the 128 keys of a tile in halves).
// smem: sQ[2] | sdO[2] | sK | sV | sP[2] | sdS[2] | sdQ[2] | mask[256] | lse2[256] | D[256] | barriers  (~195 KB)
// TMEM: S [0,128) dP [128,256) dV [256,320) dK [320,384) dQ [384,448)
constexpr int B2_CT = 256;
constexpr int B2_THREADS = 32 + B2_CT;
constexpr int B2_SMEM = 12 * TILE_BYTES + 3 * 256 * 4 + 128 + 1024;

template <bool DROP>
__global__ void __launch_bounds__(B2_THREADS, 1)
attn_bwd_tc2_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_dctx,
                    const __grid_constant__ CUtensorMap tm_dqkv, const int32_t* __restrict__ attn_mask,
                    const float* __restrict__ lse_in, const bf16* __restrict__ ctx, const bf16* __restrict__ dctx,
                    int S, int heads, int nseq, Drop drop) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                      // 2 tiles
  uint8_t* sdO = sQ + 2 * TILE_BYTES;      // 2 tiles
  uint8_t* sK = sdO + 2 * TILE_BYTES;      // key tile j
  uint8_t* sV = sK + TILE_BYTES;
  uint8_t* sP = sV + TILE_BYTES;           // [i 128][j 128] as 2 blocks of 64 columns
  uint8_t* sdS = sP + 2 * TILE_BYTES;
  uint8_t* sdQ = sdS + 2 * TILE_BYTES;     // 2 tiles: dQ_i summed over the key tiles (bf16)
  float* sMask = reinterpret_cast<float*>(sdQ + 2 * TILE_BYTES);  // [256]
  float* sLse = sMask + 256;                                      // [256] log2-domain LSE (+inf beyond S)
  float* sD = sLse + 256;                                         // [256] D_i = rowsum(dO_i * O_i)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sD + 256);
  uint64_t *b_q = bars, *b_kv = bars + 1, *b_s = bars + 2, *b_p = bars + 3, *b_o = bars + 4, *b_kvfree = bars + 5,
           *b_free = bars + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = heads * 64;
  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_qkv);
      tma_prefetch_desc(&tm_dctx);
      tma_prefetch_desc(&tm_dqkv);
      mbar_init(b_q, 1); mbar_init(b_kv, 1); mbar_init(b_s, 1); mbar_init(b_p, 8); mbar_init(b_o, 1);
      mbar_init(b_kvfree, 1); mbar_init(b_free, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tdP = tmem + 128, tdV = tmem + 256, tdK = tmem + 320, tdQ = tmem + 384;
  const int nprob = nseq * heads;

  if (warp == 0) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16_f32(128, 128, 0, 0);
      constexpr uint32_t idesc_t = make_idesc_bf16_f32(128, 64, 1, 1);
      constexpr uint32_t idesc_q = make_idesc_bf16_f32(128, 64, 0, 1);
      uint32_t ph = 0;    // per problem: b_q, b_free
      uint32_t kvp = 0;   // per key tile: b_kv, b_kvfree
      uint32_t sp = 0;    // per (j, i) step: b_s, b_p, b_o
      for (int prob = blockIdx.x; prob < nprob; prob += gridDim.x) {
        const int seq = prob / heads, h = prob - seq * heads;
        mbar_wait(b_free, ph ^ 1);  // previous problem: every MMA retired, every staged tile stored
        mbar_arrive_expect_tx(b_q, 4 * TILE_BYTES);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          tma_load_3d(sQ + t * TILE_BYTES, &tm_qkv, b_q, h * 64, t * 128, seq);
          tma_load_3d(sdO + t * TILE_BYTES, &tm_dctx, b_q, h * 64, t * 128, seq);
        }
        for (int j = 0; j < 2; ++j) {
          mbar_wait(b_kvfree, kvp ^ 1);  // dK / dV of the previous key tile have left sK / sV
          mbar_arrive_expect_tx(b_kv, 2 * TILE_BYTES);
          tma_load_3d(sK, &tm_qkv, b_kv, H + h * 64, j * 128, seq);
          tma_load_3d(sV, &tm_qkv, b_kv, 2 * H + h * 64, j * 128, seq);
          if (j == 0) mbar_wait(b_q, ph);
          mbar_wait(b_kv, kvp);
          tcgen05_fence_after();
          const uint64_t dk = desc_kmajor(smem_u32(sK)), dv = desc_kmajor(smem_u32(sV));
          const uint64_t bk = desc_mnmajor(smem_u32(sK), 0);
          const uint64_t dpt = desc_mnmajor(smem_u32(sP), TILE_BYTES), dst = desc_mnmajor(smem_u32(sdS), TILE_BYTES);
          for (int i = 0; i < 2; ++i) {
            const uint64_t dq = desc_kmajor(smem_u32(sQ + i * TILE_BYTES)), ddo = desc_kmajor(smem_u32(sdO + i * TILE_BYTES));
            const uint64_t bq = desc_mnmajor(smem_u32(sQ + i * TILE_BYTES), 0), bdo = desc_mnmajor(smem_u32(sdO + i * TILE_BYTES), 0);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(tS, dq + 2 * k, dk + 2 * k, idesc_s, k > 0);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(tdP, ddo + 2 * k, dv + 2 * k, idesc_s, k > 0);
            umma_commit(b_s);
            mbar_wait(b_p, sp);
            tcgen05_fence_after();
#pragma unroll
            for (int k = 0; k < 8; ++k) umma_f16(tdV, dpt + 128 * k, bdo + 128 * k, idesc_t, (i > 0 || k > 0) ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 8; ++k) umma_f16(tdK, dst + 128 * k, bq + 128 * k, idesc_t, (i > 0 || k > 0) ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const uint64_t a = desc_kmajor(smem_u32(sdS + (k >> 2) * TILE_BYTES)) + 2 * (k & 3);
              umma_f16(tdQ, a, bk + 128 * k, idesc_q, k > 0);
            }
            umma_commit(b_o);
            if (i == 1) mbar_wait(b_o, sp);  // observer: all MMAs reading sK / sV of this key tile have retired
            sp ^= 1;
          }
          kvp ^= 1;
        }
        ph ^= 1;
      }
    }
    __syncwarp();
  } else {
    const int tid = threadIdx.x - 32;            // 0..255
    const int cw = warp - 1;
    const int quarter = warp & 3;
    const int half = cw >> 2;
    const int row = quarter * 32 + lane;         // row inside a 128-row tile
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    uint32_t sp = 0;
    for (int prob = blockIdx.x; prob < nprob; prob += gridDim.x) {
      const int seq = prob / heads, h = prob - seq * heads;
      named_bar_sync(1, B2_CT);
      {
        const int r = tid;  // one sequence row per thread
        const bool keep = r < S && (attn_mask == nullptr || attn_mask[(long long)seq * S + r] != 0);
        sMask[r] = keep ? 0.f : -INFINITY;
        sLse[r] = r < S ? lse_in[((long long)seq * heads + h) * S + r] * LOG2E : INFINITY;
        float d = 0.f;
        if (r < S) {
          const long long off = ((long long)seq * S + r) * H + h * 64;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 a = *reinterpret_cast<const uint4*>(dctx + off + c * 8), b = *reinterpret_cast<const uint4*>(ctx + off + c * 8);
            const uint32_t* pa = &a.x; const uint32_t* pb = &b.x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 x = unpack_bf16x2(pa[k]), y = unpack_bf16x2(pb[k]);
              d = fmaf(x.x, y.x, d); d = fmaf(x.y, y.y, d);
            }
          }
        }
        sD[r] = d;
      }
      named_bar_sync(1, B2_CT);
      for (int j = 0; j < 2; ++j) {
        for (int i = 0; i < 2; ++i) {
          const int grow = i * 128 + row;
          const float lse2 = sLse[grow], D = sD[grow];
          mbar_wait(b_s, sp);
          tcgen05_fence_after();
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t rs[32], rd[32];
            tmem_ld_32x32(tS + lane_addr + half * 64 + c * 32, rs);
            tmem_ld_32x32(tdP + lane_addr + half * 64 + c * 32, rd);
            tmem_ld_wait();
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              float p[8], pd[8], ds[8];
#pragma unroll
              for (int t = 0; t < 8; ++t) {
                const int jj = c4 * 8 + t;
                p[t] = ex2_approx(fmaf(__uint_as_float(rs[jj]), SCALE_LOG2, sMask[j * 128 + half * 64 + c * 32 + jj]) - lse2);
                pd[t] = p[t];
                ds[t] = __uint_as_float(rd[jj]);
              }
              if (DROP) {
#pragma unroll
                for (int t = 0; t < 8; t += 8) {
                  float2 m[4];
                  drop.mul8((uint32_t)(prob * S + grow), (uint32_t)(j * 128 + half * 64 + c * 32 + c4 * 8), m);
#pragma unroll
                  for (int w = 0; w < 4; ++w) {
                    pd[2 * w] *= m[w].x; pd[2 * w + 1] *= m[w].y;
                    ds[2 * w] *= m[w].x; ds[2 * w + 1] *= m[w].y;
                  }
                }
              }
#pragma unroll
              for (int t = 0; t < 8; ++t) ds[t] = p[t] * (ds[t] - D);
              const int chunk = c * 4 + c4;
              st_chunk(sP + half * TILE_BYTES, row, chunk, pd, 1.f);
              st_chunk(sdS + half * TILE_BYTES, row, chunk, ds, 0.125f);
            }
          }
          fence_proxy_async_smem();
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(b_p);
          mbar_wait(b_o, sp);
          tcgen05_fence_after();
          {
            // dQ_i += dS_ij K_j : this thread owns columns [32*half, +32) of its row; partial sums live in smem (bf16)
            uint32_t r[32];
            tmem_ld_32x32(tdQ + lane_addr + half * 32, r);
            tmem_ld_wait();
            uint8_t* acc = sdQ + i * TILE_BYTES;
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              float v[8];
#pragma unroll
              for (int t = 0; t < 8; ++t) v[t] = __uint_as_float(r[c4 * 8 + t]);
              const int chunk = half * 4 + c4;
              if (j > 0) {
                const uint4 q = *reinterpret_cast<const uint4*>(acc + row * 128 + ((chunk ^ (row & 7)) << 4));
                const uint32_t* pq = &q.x;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  const float2 f = unpack_bf16x2(pq[t]);
                  v[2 * t] += f.x; v[2 * t + 1] += f.y;
                }
              }
              st_chunk(acc, row, chunk, v, 1.f);
            }
          }
          if (i == 1) {
            // dK_j, dV_j are complete (accumulated over both query tiles): stage into sK / sV (dead: the control thread
            // observed b_o of this step before it may reload them) and store
#pragma unroll
            for (int which = 0; which < 2; ++which) {
              uint32_t r[32];
              tmem_ld_32x32((which == 0 ? tdK : tdV) + lane_addr + half * 32, r);
              tmem_ld_wait();
              uint8_t* dst = which == 0 ? sK : sV;
#pragma unroll
              for (int c4 = 0; c4 < 4; ++c4) {
                float v[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) v[t] = __uint_as_float(r[c4 * 8 + t]);
                st_chunk(dst, row, half * 4 + c4, v, 1.f);
              }
            }
            fence_proxy_async_smem();
            tcgen05_fence_before();
            named_bar_sync(2, B2_CT);
            if (tid == 0) {
              tma_store_3d(&tm_dqkv, smem_u32(sK), H + h * 64, j * 128, seq);
              tma_store_3d(&tm_dqkv, smem_u32(sV), 2 * H + h * 64, j * 128, seq);
              tma_store_commit();
              tma_store_wait_read();
              mbar_arrive(b_kvfree);
            }
          }
          tcgen05_fence_before();
          sp ^= 1;
        }
      }
      // dQ of both query tiles
      fence_proxy_async_smem();
      named_bar_sync(2, B2_CT);
      if (tid == 0) {
        tma_store_3d(&tm_dqkv, smem_u32(sdQ), h * 64, 0, seq);
        tma_store_3d(&tm_dqkv, smem_u32(sdQ + TILE_BYTES), h * 64, 128, seq);
        tma_store_commit();
        tma_store_wait_read();
        mbar_arrive(b_free);
      }
    }
    if (tid == 0) tma_store_wait_all();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn2() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}
int make_tmap3b(CUtensorMap* out, const void* base, int nseq, int S, long long cols) {
  EncodeTiledFn fn = encode_fn2();
  DPRB_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  DPRB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && cols % 8 == 0, "attention operand misaligned");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)S, (cuuint64_t)nseq};
  cuuint64_t strides[2] = {(cuuint64_t)cols * 2, (cuuint64_t)S * cols * 2};
  cuuint32_t box[3] = {64u, 128u, 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DPRB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(3d) failed with CUresult %d", (int)r);
  return 0;
}

}  // namespace

int attn_fwd_tc2(const void* qkv, const int32_t* attn_mask, void* ctx, float* lse, int nseq, int S, int heads,
                 float dropout_p, unsigned long long site_seed, cudaStream_t stream) {
  DPRB_REQUIRE(S > 128 && S <= 256, "attn_fwd_tc2: S=%d outside (128, 256]", S);
  const Drop drop = drop_from_site(dropout_p, site_seed);
  const int H = heads * 64;
  CUtensorMap tq, tc;
  if (int rc = make_tmap3b(&tq, qkv, nseq, S, 3LL * H)) return rc;
  if (int rc = make_tmap3b(&tc, ctx, nseq, S, H)) return rc;
  static bool attr = false;
  if (!attr) {
    DPRB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, F2_SMEM));
    DPRB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, F2_SMEM));
    attr = true;
  }
  int sms = num_sms();
  if (sms <= 0) sms = 148;
  const int nprob = nseq * heads;
  const int grid = nprob < sms ? nprob : sms;
  if (drop.on()) attn_fwd_tc2_kernel<true><<<grid, F2_THREADS, F2_SMEM, stream>>>(tq, tc, attn_mask, lse, S, heads, nseq, drop);
  else attn_fwd_tc2_kernel<false><<<grid, F2_THREADS, F2_SMEM, stream>>>(tq, tc, attn_mask, lse, S, heads, nseq, drop);
  DPRB_LAUNCH_CHECK();
  return 0;
}

int attn_bwd_tc2(const void* qkv, const int32_t* attn_mask, const void* ctx, const float* lse, const void* dctx,
                 void* dqkv, int nseq, int S, int heads, float dropout_p, unsigned long long site_seed,
                 cudaStream_t stream) {
  DPRB_REQUIRE(S > 128 && S <= 256, "attn_bwd_tc2: S=%d outside (128, 256]", S);
  const Drop drop = drop_from_site(dropout_p, site_seed);
  const int H = heads * 64;
  CUtensorMap tq, tdo, tdq;
  if (int rc = make_tmap3b(&tq, qkv, nseq, S, 3LL * H)) return rc;
  if (int rc = make_tmap3b(&tdo, dctx, nseq, S, H)) return rc;
  if (int rc = make_tmap3b(&tdq, dqkv, nseq, S, 3LL * H)) return rc;
  static bool attr = false;
  if (!attr) {
    DPRB_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, B2_SMEM));
    DPRB_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, B2_SMEM));
    attr = true;
  }
  int sms = num_sms();
  if (sms <= 0) sms = 148;
  const int nprob = nseq * heads;
  const int grid = nprob < sms ? nprob : sms;
  if (drop.on())
    attn_bwd_tc2_kernel<true><<<grid, B2_THREADS, B2_SMEM, stream>>>(tq, tdo, tdq, attn_mask, lse, (const bf16*)ctx, (const bf16*)dctx, S, heads, nseq, drop);
  else
    attn_bwd_tc2_kernel<false><<<grid, B2_THREADS, B2_SMEM, stream>>>(tq, tdo, tdq, attn_mask, lse, (const bf16*)ctx, (const bf16*)dctx, S, heads, nseq, drop);
  DPRB_LAUNCH_CHECK();
  return 0;
}

}  // namespace dprb

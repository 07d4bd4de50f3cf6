// Self-attention core on the 5th-gen tensor cores (tcgen05 / TMEM / TMA) for head_dim 64 and S <= 128:
// one (sequence, head) problem = one 128-row UMMA tile; persistent CTAs loop over problems.
//
//   forward : S = Q K^T (UMMA 128x128x16 x4) -> TMEM -> row softmax in registers (thread i = query row i)
//             -> P (bf16) to 128B-swizzled smem -> O = P V (UMMA 128x64x16 x8, V read in place as an MN-major
//             operand) -> TMEM -> scale by 1/l -> smem -> TMA store.  LSE saved for backward.
//   backward: S = Q K^T and dP = dO V^T (UMMA) -> P = exp2(S - lse), D_i = sum_j P dP, dS = P (dP - D) / 8
//             in registers -> P, dS (bf16) to smem -> dV = P^T dO, dK = dS^T Q, dQ = dS K (UMMA; the transposed
//             operands are the SAME smem tiles read through MN-major descriptors) -> TMEM -> smem -> TMA store.
//
// All tiles are moved by TMA through 3-D tensor maps [nseq, S, columns]: rows >= S of a short sequence are
// zero-filled on load and clipped on store by the hardware, so no per-row predication is needed.
//
// Replaces BertSelfAttention.forward's scaled_dot_product_attention
// (site-packages/transformers/models/bert/modeling_bert.py:168-207, integrations/sdpa_attention.py:92-101)
// and its autograd backward, reached from /root/reference/dpr_scale/models/hf_model.py:38.
#include "common.cuh"
#include "dprb_internal.h"

namespace dprb {
namespace {

constexpr int TILE_BYTES = 128 * 128;  // [128 rows][64 bf16], 128B-swizzled = 16 KB
constexpr float SCALE_LOG2 = 0.125f * 1.4426950408889634f;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr int SM_THREADS = 128;        // softmax / epilogue threads (4 warps = 128 TMEM lanes)
constexpr int NTHREADS = 32 + SM_THREADS;

__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// K-major A/B tile [128 rows][64 K] (one swizzle atom wide): 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t addr) { return make_umma_desc_sw128(addr, 0, 1024); }
// MN-major operand read from a [K rows][64 MN] tile; further 64-wide MN atoms are `lbo` bytes apart.
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t addr, uint32_t lbo) { return make_umma_desc_sw128(addr, lbo, 1024); }

// pack 8 floats (scaled) into one 16-byte chunk of a swizzled [128][64] tile row
__device__ __forceinline__ void st_chunk(uint8_t* tile, int row, int chunk, const float* v, float s) {
  uint4 q;
  q.x = pack_bf16x2(v[0] * s, v[1] * s); q.y = pack_bf16x2(v[2] * s, v[3] * s);
  q.z = pack_bf16x2(v[4] * s, v[5] * s); q.w = pack_bf16x2(v[6] * s, v[7] * s);
  *reinterpret_cast<uint4*>(tile + row * 128 + ((chunk ^ (row & 7)) << 4)) = q;
}

// ------------------------------------------------------------------------------------------ forward
// smem: sQ | sK | sV | mask[2][128] | barriers (51 KB) - the two P k-blocks ALIAS sQ | sK (both are dead once the S MMA
// has completed, which is what the softmax threads wait for before they write P), and the staged output aliases sQ again
// once the P V MMA has completed.  TMEM: S cols [0,128); O reuses cols [0,64) (S is dead once P has been written).
// => 128 TMEM columns and 51 KB per CTA: FOUR co-resident CTAs per SM (was two at 83 KB / 256 columns).  Each CTA is a
// serial load -> MMA -> softmax -> MMA -> store chain of ~6 us per (sequence, head); with two in flight the SM moved
// 2.84 TB/s chip-wide (0.44 of the HBM peak, ncu r1), bounded by that chain's latency, not by bandwidth.
constexpr int FWD_SMEM = 3 * TILE_BYTES + 2 * 128 * 4 + 64 + 1024;
constexpr int FWD_CTAS_PER_SM = 4;

template <bool DROP>
__global__ void __launch_bounds__(NTHREADS, FWD_CTAS_PER_SM)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_ctx,
                   const int32_t* __restrict__ attn_mask, float* __restrict__ lse_out, int S, int heads, int nseq,
                   Drop drop) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + TILE_BYTES;
  uint8_t* sV = sK + TILE_BYTES;
  uint8_t* sP = sQ;               // 2 x 16 KB over sQ | sK (see above)
  float* sMask = reinterpret_cast<float*>(sV + TILE_BYTES);  // [2][128]
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
    tmem_alloc(tmem_slot, 128);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tO = tmem;
  const int nprob = nseq * heads;

  if (warp == 0) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16_f32(128, 128, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16_f32(128, 64, 0, 1);
      uint32_t ph = 0;
      for (int prob = blockIdx.x; prob < nprob; prob += gridDim.x) {
        const int seq = prob / heads, h = prob - seq * heads;
        mbar_wait(b_free, ph ^ 1);
        mbar_arrive_expect_tx(b_load, 3 * TILE_BYTES);
        tma_load_3d(sQ, &tm_qkv, b_load, h * 64, 0, seq);
        tma_load_3d(sK, &tm_qkv, b_load, H + h * 64, 0, seq);
        tma_load_3d(sV, &tm_qkv, b_load, 2 * H + h * 64, 0, seq);
        mbar_wait(b_load, ph);
        tcgen05_fence_after();
        const uint64_t dq = desc_kmajor(smem_u32(sQ)), dk = desc_kmajor(smem_u32(sK));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tS, dq + 2 * k, dk + 2 * k, idesc_s, k > 0);
        umma_commit(b_s);
        mbar_wait(b_p, ph);
        tcgen05_fence_after();
        const uint64_t dv = desc_mnmajor(smem_u32(sV), 0);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t dp = desc_kmajor(smem_u32(sP + (k >> 2) * TILE_BYTES)) + 2 * (k & 3);
          umma_f16(tO, dp, dv + 128 * k, idesc_o, k > 0);  // V: 16 key rows = 2048 B per K step
        }
        umma_commit(b_o);
        ph ^= 1;
      }
    }
    __syncwarp();
  } else {
    const int tid = threadIdx.x - 32;            // 0..127
    const int quarter = warp & 3;                // TMEM lane quarter of this warp
    const int row = quarter * 32 + lane;         // query row owned by this thread
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    uint32_t ph = 0;
    for (int prob = blockIdx.x; prob < nprob; prob += gridDim.x) {
      const int seq = prob / heads, h = prob - seq * heads;
      float* mk = sMask + ph * 128;
      {
        const bool keep = tid < S && (attn_mask == nullptr || attn_mask[(long long)seq * S + tid] != 0);
        mk[tid] = keep ? 0.f : -INFINITY;
      }
      named_bar_sync(1, SM_THREADS);
      mbar_wait(b_s, ph);
      tcgen05_fence_after();
      // pass 1: row max (TMEM reads are cheap: re-reading S beats holding 128 scores in registers)
      float m = -INFINITY;
      const float2 sc2 = make_float2(SCALE_LOG2, SCALE_LOG2);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tS + lane_addr + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 2) {   // packed fp32 (FFMA2): the softmax passes are issue-bound, not MUFU-bound
          const float2 a = __ffma2_rn(make_float2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), sc2,
                                      *reinterpret_cast<const float2*>(mk + c * 32 + j));
          m = fmaxf(m, fmaxf(a.x, a.y));
        }
      }
      const float e = (m == -INFINITY) ? 0.f : m;  // fully masked row guard
      const float2 ne2 = make_float2(-e, -e);
      // pass 2: P = exp2(s - max) -> bf16 -> swizzled smem (A operand of P V), row sum
      float2 l2 = make_float2(0.f, 0.f);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tS + lane_addr + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          float p[8];
          float2 mm4[4];
          if (DROP) drop.mul8((uint32_t)(prob * S + row), (uint32_t)(c * 32 + c4 * 8), mm4);
#pragma unroll
          for (int t = 0; t < 8; t += 2) {
            const int j = c4 * 8 + t;
            const float2 a = __fadd2_rn(__ffma2_rn(make_float2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), sc2,
                                                   *reinterpret_cast<const float2*>(mk + c * 32 + j)), ne2);
            float2 pv = make_float2(ex2_approx(a.x), ex2_approx(a.y));
            l2 = __fadd2_rn(l2, pv);
            if (DROP) {
              // attention-probability dropout (modeling_bert.py: dropout on the softmax output): the row sum keeps
              // the un-dropped value, only the P V operand is masked and rescaled
              pv = __fmul2_rn(pv, mm4[t >> 1]);
            }
            p[t] = pv.x; p[t + 1] = pv.y;
          }
          const int chunk = c * 4 + c4;
          st_chunk(sP + (chunk >> 3) * TILE_BYTES, row, chunk & 7, p, 1.f);
        }
      }
      const float l = l2.x + l2.y;
      fence_proxy_async_smem();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_p);
      if (lse_out != nullptr && row < S) lse_out[((long long)seq * heads + h) * S + row] = m * LN2 + __logf(l);
      const float inv = l > 0.f ? 1.f / l : 0.f;
      mbar_wait(b_o, ph);
      tcgen05_fence_after();
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tO + lane_addr + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          float v[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) v[t] = __uint_as_float(r[c4 * 8 + t]);
          st_chunk(sQ, row, c * 4 + c4, v, inv);  // Q tile is dead after the S MMA: reuse as the store staging tile
        }
      }
      fence_proxy_async_smem();
      tcgen05_fence_before();
      named_bar_sync(2, SM_THREADS);
      if (tid == 0) {
        tma_store_3d(&tm_ctx, smem_u32(sQ), h * 64, 0, seq);
        tma_store_commit();
        tma_store_wait_read();
        mbar_arrive(b_free);
      }
      ph ^= 1;
    }
    if (tid == 0) tma_store_wait_all();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

// ------------------------------------------------------------------------------------------ backward
// 1 control warp + 8 compute warps.  Two warps share each TMEM lane quarter and split the 128 key columns in halves.
// Inputs (Q, K, V, dO) are double-buffered: the tiles of problem n+1 are requested while problem n is processed.
// smem: 2 x (sQ | sK | sV | sdO) | sP (2 blocks) | sdS (2 blocks) | mask[2][128] | Dpart[2][128] | barriers (~195 KB)
// TMEM: S [0,128) dP [128,256) dV [256,320) dK [320,384) dQ [384,448)
constexpr int BWD_CW = 8;                       // compute warps
constexpr int BWD_CT = BWD_CW * 32;             // compute threads
constexpr int BWD_THREADS = 32 + BWD_CT;
constexpr int BWD_SMEM = 8 * TILE_BYTES + 4 * TILE_BYTES + 4 * 128 * 4 + 128 + 1024;

template <bool DROP>
__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_dctx,
                   const __grid_constant__ CUtensorMap tm_dqkv, const int32_t* __restrict__ attn_mask,
                   const float* __restrict__ lse_in, float* __restrict__ dbias, int S, int heads, int nseq, Drop drop) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sIn = smem;                       // [2][4][TILE_BYTES]: Q, K, V, dO
  uint8_t* sP = sIn + 8 * TILE_BYTES;        // 2 x 16 KB  [i][j]
  uint8_t* sdS = sP + 2 * TILE_BYTES;        // 2 x 16 KB  [i][j]
  float* sMask = reinterpret_cast<float*>(sdS + 2 * TILE_BYTES);  // [2][128]
  float* sDp = sMask + 256;                                       // [2][128] partial D per column half
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDp + 256);
  uint64_t *b_load = bars /*[2]*/, *b_s = bars + 2, *b_p = bars + 3, *b_o = bars + 4, *b_free = bars + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = heads * 64;
  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_qkv);
      tma_prefetch_desc(&tm_dctx);
      tma_prefetch_desc(&tm_dqkv);
      mbar_init(&b_load[0], 1); mbar_init(&b_load[1], 1);
      mbar_init(b_s, 1); mbar_init(b_p, BWD_CW); mbar_init(b_o, 1); mbar_init(b_free, 4);
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
      constexpr uint32_t idesc_s = make_idesc_bf16_f32(128, 128, 0, 0);    // Q K^T, dO V^T
      constexpr uint32_t idesc_t = make_idesc_bf16_f32(128, 64, 1, 1);     // P^T dO, dS^T Q
      constexpr uint32_t idesc_q = make_idesc_bf16_f32(128, 64, 0, 1);     // dS K
      auto issue_loads = [&](int prob, int buf) {
        const int seq = prob / heads, h = prob - seq * heads;
        uint8_t* base = sIn + buf * 4 * TILE_BYTES;
        mbar_arrive_expect_tx(&b_load[buf], 4 * TILE_BYTES);
        tma_load_3d(base, &tm_qkv, &b_load[buf], h * 64, 0, seq);
        tma_load_3d(base + TILE_BYTES, &tm_qkv, &b_load[buf], H + h * 64, 0, seq);
        tma_load_3d(base + 2 * TILE_BYTES, &tm_qkv, &b_load[buf], 2 * H + h * 64, 0, seq);
        tma_load_3d(base + 3 * TILE_BYTES, &tm_dctx, &b_load[buf], h * 64, 0, seq);
      };
      if ((int)blockIdx.x < nprob) issue_loads(blockIdx.x, 0);
      int it = 0;
      for (int prob = blockIdx.x; prob < nprob; prob += gridDim.x, ++it) {
        const int buf = it & 1;
        // This problem's inputs were requested one iteration ago, and the S / dP columns of TMEM are free since the
        // previous problem's dS pass (b_p): issue S = Q K^T and dP = dO V^T NOW, so they run under the previous
        // problem's epilogue (which reads only the dV / dK / dQ columns and writes the OTHER input buffer) instead of
        // after it.
        mbar_wait(&b_load[buf], (it >> 1) & 1);
        tcgen05_fence_after();
        uint8_t* sQ = sIn + buf * 4 * TILE_BYTES;
        uint8_t *sK = sQ + TILE_BYTES, *sV = sQ + 2 * TILE_BYTES, *sdO = sQ + 3 * TILE_BYTES;
        const uint64_t dq = desc_kmajor(smem_u32(sQ)), dk = desc_kmajor(smem_u32(sK));
        const uint64_t dv = desc_kmajor(smem_u32(sV)), ddo = desc_kmajor(smem_u32(sdO));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tS, dq + 2 * k, dk + 2 * k, idesc_s, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tdP, ddo + 2 * k, dv + 2 * k, idesc_s, k > 0);
        umma_commit(b_s);
        // the previous problem's epilogue (staged in the other buffer) must have drained before that buffer is reloaded
        if (it > 0) mbar_wait(b_free, (it - 1) & 1);
        if (prob + (int)gridDim.x < nprob) issue_loads(prob + gridDim.x, buf ^ 1);
        mbar_wait(b_p, it & 1);
        tcgen05_fence_after();
        // transposed A operands: the [i][j] tiles read MN-major (M = j: two 64-wide atoms 16 KB apart; K = i)
        const uint64_t dpt = desc_mnmajor(smem_u32(sP), TILE_BYTES), dst = desc_mnmajor(smem_u32(sdS), TILE_BYTES);
        const uint64_t bdo = desc_mnmajor(smem_u32(sdO), 0), bq = desc_mnmajor(smem_u32(sQ), 0), bk = desc_mnmajor(smem_u32(sK), 0);
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_f16(tdV, dpt + 128 * k, bdo + 128 * k, idesc_t, k > 0);
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_f16(tdK, dst + 128 * k, bq + 128 * k, idesc_t, k > 0);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t a = desc_kmajor(smem_u32(sdS + (k >> 2) * TILE_BYTES)) + 2 * (k & 3);
          umma_f16(tdQ, a, bk + 128 * k, idesc_q, k > 0);
        }
        umma_commit(b_o);
      }
    }
    __syncwarp();
  } else {
    const int tid = threadIdx.x - 32;            // 0..255
    const int cw = warp - 1;                     // 0..7
    const int quarter = warp & 3;                // TMEM lane quarter
    const int half = cw >> 2;                    // which 64 of the 128 key columns / which 32 of the 64 output columns
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    int it = 0;
    for (int prob = blockIdx.x; prob < nprob; prob += gridDim.x, ++it) {
      const int seq = prob / heads, h = prob - seq * heads;
      const int buf = it & 1;
      float* mk = sMask + buf * 128;
      float* dpart = sDp + buf * 128 * 0;  // single buffer: guarded by the named barriers below
      if (tid < 128) {
        const bool keep = tid < S && (attn_mask == nullptr || attn_mask[(long long)seq * S + tid] != 0);
        mk[tid] = keep ? 0.f : -INFINITY;
      }
      // rows beyond S: lse = +inf  =>  P = 0
      const float lse2 = row < S ? lse_in[((long long)seq * heads + h) * S + row] * LOG2E : INFINITY;
      named_bar_sync(1, BWD_CT);
      mbar_wait(b_s, it & 1);
      tcgen05_fence_after();
      // pass 1 (own 64 columns): P (packed in registers + smem block `half`), partial D_i = sum_j P_ij dP_ij
      // With attention dropout: P_d = P * mask/(1-p) feeds dV; dP arrives w.r.t. P_d, so dP_m = dP * mask/(1-p),
      // D_i = sum_j P_ij dP_m_ij and dS = P (dP_m - D).  The keep bits of this thread's 64 columns live in one register.
      // (packed fp32 throughout: these passes are bound by instruction issue at the power-capped clock)
      uint32_t pk[32];
      uint64_t keep = ~0ull;
      float2 D2 = make_float2(0.f, 0.f);
      const float2 sc2 = make_float2(SCALE_LOG2, SCALE_LOG2), nl2 = make_float2(-lse2, -lse2);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t rs[32], rd[32];
        tmem_ld_32x32(tS + lane_addr + half * 64 + c * 32, rs);
        tmem_ld_32x32(tdP + lane_addr + half * 64 + c * 32, rd);
        tmem_ld_wait();
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          uint32_t qd[4];
          float2 mm4[4];
          if (DROP) drop.mul8((uint32_t)(prob * S + row), (uint32_t)(half * 64 + c * 32 + c4 * 8), mm4);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int j = c4 * 8 + 2 * t;
            const float2 a = __fadd2_rn(__ffma2_rn(make_float2(__uint_as_float(rs[j]), __uint_as_float(rs[j + 1])), sc2,
                                                   *reinterpret_cast<const float2*>(mk + half * 64 + c * 32 + j)), nl2);
            const float2 pv = make_float2(ex2_approx(a.x), ex2_approx(a.y));
            float2 pdv = pv;
            if (DROP) {
              const float2 mm = mm4[t];
              if (mm.x == 0.f) keep &= ~(1ull << (c * 32 + j));
              if (mm.y == 0.f) keep &= ~(1ull << (c * 32 + j + 1));
              pdv = __fmul2_rn(pv, mm);
            }
            D2 = __ffma2_rn(pdv, make_float2(__uint_as_float(rd[j]), __uint_as_float(rd[j + 1])), D2);
            pk[c * 16 + c4 * 4 + t] = pack_bf16x2(pv.x, pv.y);
            qd[t] = DROP ? pack_bf16x2(pdv.x, pdv.y) : pk[c * 16 + c4 * 4 + t];
          }
          const int chunk = c * 4 + c4;  // 16-byte chunk inside this half's 64-column block
          *reinterpret_cast<uint4*>(sP + half * TILE_BYTES + row * 128 + ((chunk ^ (row & 7)) << 4)) =
              make_uint4(qd[0], qd[1], qd[2], qd[3]);
        }
      }
      float D = D2.x + D2.y;
      dpart[half * 128 + row] = D;
      named_bar_sync(2, BWD_CT);
      D = dpart[row] + dpart[128 + row];
      const float2 nD2 = make_float2(-D, -D);
      // pass 2: dS = P (dP - D) / sqrt(dh)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t rd[32];
        tmem_ld_32x32(tdP + lane_addr + half * 64 + c * 32, rd);
        tmem_ld_wait();
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          float ds[8];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 pp = unpack_bf16x2(pk[c * 16 + c4 * 4 + t]);
            const float2 dv = make_float2(__uint_as_float(rd[c4 * 8 + 2 * t]), __uint_as_float(rd[c4 * 8 + 2 * t + 1]));
            float2 inner;
            if (DROP) {
              const int j0 = c * 32 + c4 * 8 + 2 * t;
              const float2 mm = make_float2(((keep >> j0) & 1ull) ? drop.scale : 0.f,
                                            ((keep >> (j0 + 1)) & 1ull) ? drop.scale : 0.f);
              inner = __ffma2_rn(dv, mm, nD2);
            } else {
              inner = __fadd2_rn(dv, nD2);
            }
            const float2 o = __fmul2_rn(pp, inner);
            ds[2 * t] = o.x; ds[2 * t + 1] = o.y;
          }
          st_chunk(sdS + half * TILE_BYTES, row, c * 4 + c4, ds, 0.125f);
        }
      }
      fence_proxy_async_smem();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_p);
      mbar_wait(b_o, it & 1);
      tcgen05_fence_after();
      // epilogue: dQ -> sQ, dK -> sK, dV -> sV of this problem's input buffer (dead once b_o has fired);
      // this warp converts columns [32*half, +32) of its 32 rows for each of the three tiles
      uint8_t* sQ = sIn + buf * 4 * TILE_BYTES;
#pragma unroll
      for (int which = 0; which < 3; ++which) {
        const uint32_t t0 = which == 0 ? tdQ : (which == 1 ? tdK : tdV);
        uint8_t* dst = sQ + which * TILE_BYTES;
        uint32_t r[32];
        tmem_ld_32x32(t0 + lane_addr + half * 32, r);
        tmem_ld_wait();
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
      named_bar_sync(3, BWD_CT);
      if (cw < 3) {
        // bias gradient of the fused QKV projection: column sums of the staged (bf16) dQ / dK / dV tiles.
        // warp cw sums tile cw; lane l owns columns 2l, 2l+1 (a 128-byte row per read: conflict-free)
        if (dbias != nullptr) {
          const uint8_t* tile = sQ + cw * TILE_BYTES;
          float s0 = 0.f, s1 = 0.f;
          const int nrows = S < 128 ? S : 128;
          for (int r = 0; r < nrows; ++r) {
            const float2 f = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(
                tile + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + (lane & 3) * 4));
            s0 += f.x; s1 += f.y;
          }
          atomicAdd(dbias + cw * H + h * 64 + 2 * lane, s0);
          atomicAdd(dbias + cw * H + h * 64 + 2 * lane + 1, s1);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(b_free);
      } else if (cw == 3) {
        if (lane == 0) {
          tma_store_3d(&tm_dqkv, smem_u32(sQ), h * 64, 0, seq);
          tma_store_3d(&tm_dqkv, smem_u32(sQ + TILE_BYTES), H + h * 64, 0, seq);
          tma_store_3d(&tm_dqkv, smem_u32(sQ + 2 * TILE_BYTES), 2 * H + h * 64, 0, seq);
          tma_store_commit();
          tma_store_wait_read();
          mbar_arrive(b_free);
        }
      }
    }
    if (cw == 3 && lane == 0) tma_store_wait_all();
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
EncodeTiledFn encode_fn() {
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

// bf16 [nseq, S, cols] (row stride `cols` elements), box = [1, 128 rows, 64 cols], 128B swizzle
int make_tmap3(CUtensorMap* out, const void* base, int nseq, int S, long long cols) {
  EncodeTiledFn fn = encode_fn();
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

int attn_fwd_tc(const void* qkv, const int32_t* attn_mask, void* ctx, float* lse, int nseq, int S, int heads,
                float dropout_p, unsigned long long site_seed, cudaStream_t stream) {
  const Drop drop = drop_from_site(dropout_p, site_seed);
  const int H = heads * 64;
  CUtensorMap tq, tc;
  if (int rc = make_tmap3(&tq, qkv, nseq, S, 3LL * H)) return rc;
  if (int rc = make_tmap3(&tc, ctx, nseq, S, H)) return rc;
  static bool attr = false;
  if (!attr) {
    DPRB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    DPRB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    attr = true;
  }
  int sms = num_sms();
  if (sms <= 0) sms = 148;
  const int nprob = nseq * heads;
  const int grid = nprob < FWD_CTAS_PER_SM * sms ? nprob : FWD_CTAS_PER_SM * sms;  // co-resident CTAs interleave their serial chains
  if (drop.on()) attn_fwd_tc_kernel<true><<<grid, NTHREADS, FWD_SMEM, stream>>>(tq, tc, attn_mask, lse, S, heads, nseq, drop);
  else attn_fwd_tc_kernel<false><<<grid, NTHREADS, FWD_SMEM, stream>>>(tq, tc, attn_mask, lse, S, heads, nseq, drop);
  DPRB_LAUNCH_CHECK();
  return 0;
}

int attn_bwd_tc(const void* qkv, const int32_t* attn_mask, const float* lse, const void* dctx, void* dqkv,
                float* dbias, int nseq, int S, int heads, float dropout_p, unsigned long long site_seed,
                cudaStream_t stream) {
  const Drop drop = drop_from_site(dropout_p, site_seed);
  const int H = heads * 64;
  CUtensorMap tq, tdo, tdq;
  if (int rc = make_tmap3(&tq, qkv, nseq, S, 3LL * H)) return rc;
  if (int rc = make_tmap3(&tdo, dctx, nseq, S, H)) return rc;
  if (int rc = make_tmap3(&tdq, dqkv, nseq, S, 3LL * H)) return rc;
  static bool attr = false;
  if (!attr) {
    DPRB_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    DPRB_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    attr = true;
  }
  int sms = num_sms();
  if (sms <= 0) sms = 148;
  const int nprob = nseq * heads;
  const int grid = nprob < sms ? nprob : sms;
  if (drop.on()) attn_bwd_tc_kernel<true><<<grid, BWD_THREADS, BWD_SMEM, stream>>>(tq, tdo, tdq, attn_mask, lse, dbias, S, heads, nseq, drop);
  else attn_bwd_tc_kernel<false><<<grid, BWD_THREADS, BWD_SMEM, stream>>>(tq, tdo, tdq, attn_mask, lse, dbias, S, heads, nseq, drop);
  DPRB_LAUNCH_CHECK();
  return 0;
}

}  // namespace dprb

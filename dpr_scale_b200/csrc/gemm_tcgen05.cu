// Persistent warp-specialised bf16 GEMM for sm_100a: TMA -> 128B-swizzled smem ring -> tcgen05.mma
// (UMMA 128x256x16, fp32 accumulators in TMEM, double-buffered) -> tcgen05.ld epilogue with the
// fused bias / bias+GELU / bias+residual / dGELU / fp32 split-K accumulate variants the encoder needs.
//
// Replaces, on the reference path, every torch.nn.Linear call inside HF BertLayer
// (site-packages/transformers/models/bert/modeling_bert.py:179-181 QKV, :295 attention output,
// :340 intermediate, :353 output) and their autograd backward (dgrad / wgrad), reached from
// /root/reference/dpr_scale/models/hf_model.py:38.
//
// D[M,N] = epi( sum_k A(m,k) * B(n,k) ).  Each operand may be K-major (row = MN index, K contiguous)
// or MN-major (row = K index, MN contiguous); the latter lets dgrad read W[N_out,K_in] and wgrad
// read dY[T,N_out] / X[T,K_in] in place, with no transposed copies.
#include <vector>
#include "common.cuh"
#include "dprb_internal.h"

namespace dprb {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 256;
constexpr int BLOCK_K = 64;   // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;  // 32 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = ACC_STAGES * BLOCK_N;  // 512 = all of TMEM
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 128 + NUM_EPI_WARPS * 32;  // warps 0..3: TMA, MMA, TMEM alloc, spare
constexpr int SLAB_BYTES = 32 * 128;  // 32 rows x 64 bf16, 128B-swizzled: one TMA-store box per epilogue warp
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + NUM_EPI_WARPS * SLAB_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget exceeded");

struct GemmParams {
  int M, N, K;
  int num_m_blocks, num_n_blocks;
  int k_blocks_total, k_blocks_per_split, splits;
  int epilogue;
  void* D;
  long long ldd;
  const float* bias;   // [N] fp32 or null
  const bf16* aux;     // residual (EPI_BIAS_RESIDUAL) or pre-activation (EPI_DGELU), ld = ld_aux
  long long ld_aux;
  bf16* out2;          // EPI_BIAS_GELU: pre-activation store, ld = ldd
  float alpha;         // scale applied to the accumulator before the epilogue
  float* colsum;       // optional: colsum[n] += sum_m D(m, n) of the bf16-rounded output (bias gradients)
};

template <int A_MN, int B_MN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_d, const __grid_constant__ CUtensorMap tmap_d2,
                 const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B needs 1024-byte aligned tiles.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint8_t* smem_stage = smem + STAGES * STAGE_BYTES;  // [NUM_EPI_WARPS][SLAB_BYTES] epilogue staging slabs
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stage + NUM_EPI_WARPS * SLAB_BYTES);
  uint64_t* full_bar = bars;                       // [STAGES]
  uint64_t* empty_bar = bars + STAGES;             // [STAGES]
  uint64_t* tmem_full_bar = bars + 2 * STAGES;     // [ACC_STAGES]
  uint64_t* tmem_empty_bar = tmem_full_bar + ACC_STAGES;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + ACC_STAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_d);
    tma_prefetch_desc(&tmap_d2);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < ACC_STAGES; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], NUM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_base_slot, TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  const int tiles = p.num_m_blocks * p.num_n_blocks;
  const int units = tiles * p.splits;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int tile = u % tiles, split = u / tiles;
        const int m_blk = tile / p.num_n_blocks, n_blk = tile % p.num_n_blocks;
        const int kb0 = split * p.k_blocks_per_split;
        const int kb1 = min(kb0 + p.k_blocks_per_split, p.k_blocks_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          uint8_t* sa = smem_a + stage * A_STAGE_BYTES;
          uint8_t* sb = smem_b + stage * B_STAGE_BYTES;
          if (A_MN == 0) {
            tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
          } else {
#pragma unroll
            for (int i = 0; i < BLOCK_M / 64; ++i)
              tma_load_2d(sa + i * (BLOCK_K * 128), &tmap_a, &full_bar[stage], m_blk * BLOCK_M + i * 64, kb * BLOCK_K);
          }
          if (B_MN == 0) {
            tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BLOCK_K, n_blk * BLOCK_N);
          } else {
#pragma unroll
            for (int i = 0; i < BLOCK_N / 64; ++i)
              tma_load_2d(sb + i * (BLOCK_K * 128), &tmap_b, &full_bar[stage], n_blk * BLOCK_N + i * 64, kb * BLOCK_K);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================ MMA issuer (one thread) ================================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(BLOCK_M, BLOCK_N, A_MN, B_MN);
      // K-major SW128: 8-row groups are 1024 B apart (SBO); LBO unused inside one swizzle atom.
      // MN-major SW128: 64-element MN atoms are BLOCK_K*128 B apart (LBO); 8-deep K groups 1024 B apart (SBO).
      constexpr uint32_t A_LBO = A_MN ? BLOCK_K * 128 : 0, A_SBO = 1024;
      constexpr uint32_t B_LBO = B_MN ? BLOCK_K * 128 : 0, B_SBO = 1024;
      // advancing K by UMMA_K (16): K-major -> +32 B inside the swizzle row; MN-major -> +16 rows of 128 B.
      constexpr uint32_t A_KSTEP = (A_MN ? UMMA_K * 128 : UMMA_K * 2) >> 4;
      constexpr uint32_t B_KSTEP = (B_MN ? UMMA_K * 128 : UMMA_K * 2) >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int split = u / tiles;
        const int kb0 = split * p.k_blocks_per_split;
        const int kb1 = min(kb0 + p.k_blocks_per_split, p.k_blocks_total);
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint64_t a_desc = make_umma_desc_sw128(smem_u32(smem_a + stage * A_STAGE_BYTES), A_LBO, A_SBO);
          const uint64_t b_desc = make_umma_desc_sw128(smem_u32(smem_b + stage * B_STAGE_BYTES), B_LBO, B_SBO);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            umma_f16(d_tmem, a_desc + (uint64_t)(k * A_KSTEP), b_desc + (uint64_t)(k * B_KSTEP), idesc,
                     (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ================================ epilogue warps ================================
    // Warp (quarter, col_half) owns TMEM lanes [32*quarter, +32) x columns [128*col_half, +128) of the tile.
    // bf16 outputs: registers -> 128B-swizzled smem slab (32 rows x 64 cols) -> TMA store (full-line writes;
    // partial tiles are clipped by the tensor map).  aux (residual / pre-activation) is read with coalesced
    // 16-byte loads through the same slab.  fp32 outputs (split-K wgrad) go out as RED/ST from registers.
    const int ew = warp - 4;
    const int quarter = warp & 3;
    const int col_half = ew >> 2;
    uint8_t* slab = smem_stage + ew * SLAB_BYTES;
    const uint32_t slab_u32 = smem_u32(slab);
    const bool f32_out = (p.epilogue == DPRB_EPI_F32_ATOMIC_ADD || p.epilogue == DPRB_EPI_F32_STORE);
    const bool has_aux = (p.epilogue == DPRB_EPI_BIAS_RESIDUAL || p.epilogue == DPRB_EPI_DGELU);
    const bool is_gelu = (p.epilogue == DPRB_EPI_BIAS_GELU);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      const int tile = u % tiles;
      const int m_blk = tile / p.num_n_blocks, n_blk = tile % p.num_n_blocks;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();
      const int row0 = m_blk * BLOCK_M + quarter * 32;
      const int row = row0 + lane;
      const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BLOCK_N + col_half * 128);
      if (f32_out) {
        const bool row_ok = row < p.M;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const int col0 = n_blk * BLOCK_N + col_half * 128 + c * 32;
          uint32_t r[32];
          tmem_ld_32x32(tbase + c * 32, r);
          tmem_ld_wait();
          if (col0 < p.N && row_ok) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
            const bool full = (col0 + 32 <= p.N);
            float* dst = reinterpret_cast<float*>(p.D) + (long long)row * p.ldd + col0;
            if (p.epilogue == DPRB_EPI_F32_ATOMIC_ADD) {
              if (full) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) red_add_v4_f32(dst + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) if (col0 + j < p.N) atomicAdd(dst + j, v[j]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) dst[j] = v[j] + (p.bias != nullptr ? __ldg(p.bias + col0 + j) : 0.f);
            }
          }
          __syncwarp();  // reconverge before the next .sync.aligned tcgen05.ld
        }
      } else if (is_gelu) {
        // single pass, two 32x32 slabs (64B swizzle): pre-activation -> out2 (optional), GELU -> D
        uint8_t* slab_pre = slab;
        uint8_t* slab_act = slab + SLAB_BYTES / 2;
#pragma unroll 1
        for (int h = 0; h < 4; ++h) {
          const int col0 = n_blk * BLOCK_N + col_half * 128 + h * 32;
          if (col0 >= p.N) break;  // warp-uniform
          uint32_t r[32];
          tmem_ld_32x32(tbase + h * 32, r);
          tmem_ld_wait();
          float v[32];
          if (p.bias != nullptr && col0 + 32 <= p.N) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
              v[j] = fmaf(__uint_as_float(r[j]), p.alpha, b.x); v[j + 1] = fmaf(__uint_as_float(r[j + 1]), p.alpha, b.y);
              v[j + 2] = fmaf(__uint_as_float(r[j + 2]), p.alpha, b.z); v[j + 3] = fmaf(__uint_as_float(r[j + 3]), p.alpha, b.w);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              v[j] = fmaf(__uint_as_float(r[j]), p.alpha, (p.bias != nullptr && col0 + j < p.N) ? __ldg(p.bias + col0 + j) : 0.f);
          }
          if (lane == 0) tma_store_wait_read();
          __syncwarp();
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            uint4 qp, qa;
            uint32_t* pp = &qp.x;
            uint32_t* pa = &qa.x;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const uint32_t pre2 = pack_bf16x2(v[c4 * 8 + 2 * t], v[c4 * 8 + 2 * t + 1]);
              // GELU acts on the bf16-rounded pre-activation: backward re-reads exactly that value
              const float2 f = unpack_bf16x2(pre2);
              pp[t] = pre2;
              pa[t] = pack_bf16x2(gelu_erf(f.x), gelu_erf(f.y));
            }
            const uint32_t off = lane * 64 + ((c4 ^ ((lane >> 1) & 3)) << 4);
            if (p.out2 != nullptr) *reinterpret_cast<uint4*>(slab_pre + off) = qp;
            *reinterpret_cast<uint4*>(slab_act + off) = qa;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (p.out2 != nullptr) tma_store_2d(&tmap_d2, slab_u32, col0, row0);
            tma_store_2d(&tmap_d, slab_u32 + SLAB_BYTES / 2, col0, row0);
            tma_store_commit();
          }
        }
        __syncwarp();
      } else {
        const int n_pass = 1;
#pragma unroll 1
        for (int sl = 0; sl < 2; ++sl) {          // two 64-column slabs per warp
          const int colbase = n_blk * BLOCK_N + col_half * 128 + sl * 64;
          if (colbase >= p.N) break;              // warp-uniform
#pragma unroll 1
          for (int pass = 0; pass < n_pass; ++pass) {
            const bool store_pre = false;
            // the previous TMA store must have finished READING the slab before it is overwritten
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
            if (has_aux) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int idx = lane + 32 * i, r = idx >> 3, ch = idx & 7;
                uint4 q = make_uint4(0, 0, 0, 0);
                if (row0 + r < p.M && colbase + ch * 8 < p.N)
                  q = ldg_nc_v4(p.aux + (long long)(row0 + r) * p.ld_aux + colbase + ch * 8);
                *reinterpret_cast<uint4*>(slab + r * 128 + ((ch ^ (r & 7)) << 4)) = q;
              }
              __syncwarp();
            }
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {         // 32 accumulator columns at a time
              const int col0 = colbase + h * 32;
              uint32_t r[32];
              tmem_ld_32x32(tbase + sl * 64 + h * 32, r);
              tmem_ld_wait();
              float v[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
              if (p.bias != nullptr) {
                if (col0 + 32 <= p.N) {
#pragma unroll
                  for (int j = 0; j < 32; j += 4) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
                    v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < 32; ++j) if (col0 + j < p.N) v[j] += __ldg(p.bias + col0 + j);
                }
              }
              if (has_aux) {
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                  const int ch = h * 4 + c4;
                  const uint4 q = *reinterpret_cast<const uint4*>(slab + lane * 128 + ((ch ^ (lane & 7)) << 4));
                  const float2 f0 = unpack_bf16x2(q.x), f1 = unpack_bf16x2(q.y), f2 = unpack_bf16x2(q.z), f3 = unpack_bf16x2(q.w);
                  const float a[8] = {f0.x, f0.y, f1.x, f1.y, f2.x, f2.y, f3.x, f3.y};
#pragma unroll
                  for (int t = 0; t < 8; ++t) {
                    if (p.epilogue == DPRB_EPI_BIAS_RESIDUAL) v[c4 * 8 + t] += a[t];
                    else v[c4 * 8 + t] *= gelu_erf_grad(a[t]);
                  }
                }
              }
#pragma unroll
              for (int c4 = 0; c4 < 4; ++c4) {
                const int ch = h * 4 + c4;
                uint4 q;
                q.x = pack_bf16x2(v[c4 * 8 + 0], v[c4 * 8 + 1]); q.y = pack_bf16x2(v[c4 * 8 + 2], v[c4 * 8 + 3]);
                q.z = pack_bf16x2(v[c4 * 8 + 4], v[c4 * 8 + 5]); q.w = pack_bf16x2(v[c4 * 8 + 6], v[c4 * 8 + 7]);
                *reinterpret_cast<uint4*>(slab + lane * 128 + ((ch ^ (lane & 7)) << 4)) = q;
              }
            }
            if (p.colsum != nullptr) {
              // column sums of the staged (bf16-rounded) slab: lane l owns columns 2l, 2l+1; conflict-free reads
              __syncwarp();
              const int nrows = min(32, p.M - row0);
              float s0 = 0.f, s1 = 0.f;
              for (int r = 0; r < nrows; ++r) {
                const float2 f = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(
                    slab + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + (lane & 3) * 4));
                s0 += f.x; s1 += f.y;
              }
              const int col = colbase + lane * 2;
              if (col < p.N) atomicAdd(p.colsum + col, s0);
              if (col + 1 < p.N) atomicAdd(p.colsum + col + 1, s1);
            }
            fence_proxy_async_smem();   // make the generic-proxy smem writes visible to the TMA engine
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(store_pre ? &tmap_d2 : &tmap_d, slab_u32, colbase, row0);
              tma_store_commit();
            }
          }
        }
        __syncwarp();
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) tma_store_wait_all();  // smem must stay valid until the last bulk store has drained
    __syncwarp();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
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

// Row-major bf16 matrix [rows, cols] with leading dimension ld (elements); box = [box_rows, 64 cols], 128B swizzle.
int make_tmap(CUtensorMap* out, const void* base, long long rows, long long cols, long long ld, int box_rows,
              bool is_output = false, int box_cols = 64) {
  EncodeTiledFn fn = get_encode_fn();
  DPRB_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  DPRB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "gemm operand base %p not 16-byte aligned", base);
  DPRB_REQUIRE((ld * 2) % 16 == 0, "gemm operand leading dimension %lld not a multiple of 8 elements", ld);
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  is_output ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DPRB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld ld=%lld)",
               (int)r, rows, cols, ld);
  return 0;
}

// ---- optional live profiling: CUDA events around every GEMM launch (bench.py's roofline leg) ----
struct GemmProfile {
  bool enabled = false;
  std::vector<cudaEvent_t> ev;   // pairs
  std::vector<double> flops;
  size_t used = 0;
};
GemmProfile g_prof;

int choose_splits(int tiles, int k_blocks, int sms) {
  // minimise the makespan ceil(tiles*s/sms)/s over s, keeping >= 4 k-blocks per split
  int best = 1;
  double best_cost = 1e30;
  for (int s = 1; s <= 32; ++s) {
    if (k_blocks / s < 4 && s > 1) break;
    int waves = (tiles * s + sms - 1) / sms;
    double cost = (double)waves / s + 0.002 * s;  // small penalty for extra atomic traffic
    if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
  }
  return best;
}

}  // namespace

int gemm_bf16(const void* A, const void* B, void* D, int M, int N, int K, long long lda, long long ldb,
              long long ldd, int a_mn_major, int b_mn_major, int epilogue, const float* bias, const void* aux,
              long long ld_aux, void* out2, float alpha, int splits, float* colsum, cudaStream_t stream) {
  DPRB_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  DPRB_REQUIRE(epilogue >= 0 && epilogue < DPRB_EPI_COUNT, "gemm: bad epilogue %d", epilogue);
  const bool f32_out = (epilogue == DPRB_EPI_F32_ATOMIC_ADD || epilogue == DPRB_EPI_F32_STORE);
  DPRB_REQUIRE(f32_out || (ldd % 8 == 0 && (reinterpret_cast<uintptr_t>(D) & 15) == 0),
               "gemm: bf16 output must be 16-byte aligned with ldd %% 8 == 0 (ldd=%lld)", ldd);
  DPRB_REQUIRE(!f32_out || (ldd % 4 == 0 && (reinterpret_cast<uintptr_t>(D) & 15) == 0),
               "gemm: fp32 output must be 16-byte aligned with ldd %% 4 == 0 (ldd=%lld)", ldd);
  if (epilogue == DPRB_EPI_BIAS_RESIDUAL || epilogue == DPRB_EPI_DGELU)
    DPRB_REQUIRE(aux != nullptr && ld_aux % 8 == 0, "gemm: epilogue %d needs aux with ld %% 8 == 0", epilogue);
  if (bias != nullptr) DPRB_REQUIRE((reinterpret_cast<uintptr_t>(bias) & 15) == 0, "gemm: bias not 16B aligned");

  CUtensorMap ta, tb;
  int rc;
  if (!a_mn_major) rc = make_tmap(&ta, A, M, K, lda, BLOCK_M); else rc = make_tmap(&ta, A, K, M, lda, BLOCK_K);
  if (rc) return rc;
  if (!b_mn_major) rc = make_tmap(&tb, B, N, K, ldb, BLOCK_N); else rc = make_tmap(&tb, B, K, N, ldb, BLOCK_K);
  if (rc) return rc;
  CUtensorMap td, td2;
  if (!f32_out) {
    const int bc = (epilogue == DPRB_EPI_BIAS_GELU) ? 32 : 64;  // GELU epilogue stages two 32x32 slabs
    if ((rc = make_tmap(&td, D, M, N, ldd, 32, true, bc))) return rc;
    if (epilogue == DPRB_EPI_BIAS_GELU && out2 != nullptr) {
      DPRB_REQUIRE((reinterpret_cast<uintptr_t>(out2) & 15) == 0, "gemm: out2 not 16-byte aligned");
      if ((rc = make_tmap(&td2, out2, M, N, ldd, 32, true, bc))) return rc;
    } else {
      td2 = td;
    }
  } else {
    td = ta;  // unused by the fp32 epilogues; any valid descriptor
    td2 = ta;
  }

  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.num_m_blocks = (M + BLOCK_M - 1) / BLOCK_M;
  p.num_n_blocks = (N + BLOCK_N - 1) / BLOCK_N;
  p.k_blocks_total = (K + BLOCK_K - 1) / BLOCK_K;
  const int sms = num_sms();
  const int tiles = p.num_m_blocks * p.num_n_blocks;
  if (epilogue != DPRB_EPI_F32_ATOMIC_ADD) splits = 1;
  else if (splits <= 0) splits = choose_splits(tiles, p.k_blocks_total, sms);
  if (splits > p.k_blocks_total) splits = p.k_blocks_total;
  p.k_blocks_per_split = (p.k_blocks_total + splits - 1) / splits;
  p.splits = (p.k_blocks_total + p.k_blocks_per_split - 1) / p.k_blocks_per_split;
  p.epilogue = epilogue;
  p.D = D; p.ldd = ldd; p.bias = bias; p.aux = reinterpret_cast<const bf16*>(aux); p.ld_aux = ld_aux;
  p.out2 = reinterpret_cast<bf16*>(out2); p.alpha = alpha;
  p.colsum = colsum;
  DPRB_REQUIRE(colsum == nullptr || (!f32_out && epilogue != DPRB_EPI_BIAS_GELU),
               "gemm: colsum is supported for the BIAS / BIAS_RESIDUAL / DGELU epilogues only");

  const int units = tiles * p.splits;
  const int grid = units < sms ? units : sms;

  static bool attr_set = false;
  if (!attr_set) {
    DPRB_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    DPRB_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    DPRB_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    DPRB_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  auto launch = [&](auto kern) -> int {
    const bool prof = g_prof.enabled && g_prof.used + 2 <= g_prof.ev.size();
    if (prof) DPRB_CHECK_CUDA(cudaEventRecord(g_prof.ev[g_prof.used], stream));
    kern<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(ta, tb, td, td2, p);
    DPRB_CHECK_CUDA(cudaGetLastError());
    if (prof) {
      DPRB_CHECK_CUDA(cudaEventRecord(g_prof.ev[g_prof.used + 1], stream));
      g_prof.flops.push_back(2.0 * (double)M * (double)N * (double)K);
      g_prof.used += 2;
    }
    return 0;
  };
  if (!a_mn_major && !b_mn_major) return launch(gemm_bf16_kernel<0, 0>);
  if (!a_mn_major && b_mn_major) return launch(gemm_bf16_kernel<0, 1>);
  if (a_mn_major && !b_mn_major) return launch(gemm_bf16_kernel<1, 0>);
  return launch(gemm_bf16_kernel<1, 1>);
}

int gemm_profile_enable(int enable, int max_launches) {
  if (enable) {
    const size_t want = (size_t)max_launches * 2;
    while (g_prof.ev.size() < want) {
      cudaEvent_t e;
      DPRB_CHECK_CUDA(cudaEventCreate(&e));
      g_prof.ev.push_back(e);
    }
    g_prof.used = 0;
    g_prof.flops.clear();
  }
  g_prof.enabled = enable != 0;
  return 0;
}

// Sums the recorded launch durations (synchronises on each end event). Outputs: total ms, total FLOPs, launches.
int gemm_profile_read(double* total_ms, double* total_flops, long long* launches) {
  double ms = 0.0, fl = 0.0;
  for (size_t i = 0; i + 1 < g_prof.used; i += 2) {
    DPRB_CHECK_CUDA(cudaEventSynchronize(g_prof.ev[i + 1]));
    float t = 0.f;
    DPRB_CHECK_CUDA(cudaEventElapsedTime(&t, g_prof.ev[i], g_prof.ev[i + 1]));
    ms += t;
    fl += g_prof.flops[i / 2];
  }
  if (total_ms) *total_ms = ms;
  if (total_flops) *total_flops = fl;
  if (launches) *launches = (long long)(g_prof.used / 2);
  return 0;
}

}  // namespace dprb

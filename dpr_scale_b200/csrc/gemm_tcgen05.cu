// Persistent warp-specialised bf16 GEMM for sm_100a: TMA -> 128B-swizzled smem ring -> tcgen05.mma
// (fp32 accumulators in TMEM, double-buffered) -> tcgen05.ld epilogue -> swizzled smem slabs -> TMA store,
// with the fused bias / bias+GELU / bias+residual / dGELU / fp32 split-K accumulate variants the encoder needs.
//
// Two instantiations of one kernel:
//   CG = 2 (default): a CTA PAIR (cluster 2x1x1) computes a 256x256 tile with tcgen05.mma.cta_group::2
//           (UMMA 256x256x16): each CTA stages its own 128 rows of A and HALF of B (128 of the 256 N rows), so
//           shared-memory fill traffic per SM drops from 96 to 64 B/clk at full tensor rate (the 128 B/clk smem
//           port is what caps the 1-CTA version near 65 % of the MMA peak).  5-stage ring of 32 KB, double-buffered
//           epilogue slabs.
//   CG = 1: one CTA per 128x256 tile, UMMA 128x256x16, 4-stage ring of 48 KB (kept for A/B measurements:
//           DPRB_GEMM_1CTA=1).
//
// Replaces, on the reference path, every torch.nn.Linear call inside HF BertLayer
// (site-packages/transformers/models/bert/modeling_bert.py:179-181 QKV, :295 attention output,
// :340 intermediate, :353 output) and their autograd backward (dgrad / wgrad), reached from
// /root/reference/dpr_scale/models/hf_model.py:38.
//
// D[M,N] = epi( sum_k A(m,k) * B(n,k) ).  Each operand may be K-major (row = MN index, K contiguous)
// or MN-major (row = K index, MN contiguous); the latter lets dgrad read W[N_out,K_in] and wgrad
// read dY[T,N_out] / X[T,K_in] in place, with no transposed copies.
#include <cstdlib>
#include <vector>
#include "common.cuh"
#include "dprb_internal.h"

namespace dprb {

namespace {

constexpr int CTA_M = 128;    // accumulator rows per CTA (TMEM lanes)
constexpr int BLOCK_N = 256;
constexpr int BLOCK_K = 64;   // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = ACC_STAGES * BLOCK_N;  // 512 = all of TMEM
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 128 + NUM_EPI_WARPS * 32;  // warps 0..3: TMA, MMA, TMEM alloc, spare
constexpr int SLAB_BYTES = 32 * 128;  // 32 rows x 64 bf16, 128B-swizzled: one TMA-store box

template <int CG>
struct Cfg {
  static constexpr int STAGES = CG == 2 ? 5 : 4;
  static constexpr int A_STAGE_BYTES = CTA_M * BLOCK_K * 2;                // 16 KB
  static constexpr int B_ROWS = BLOCK_N / CG;                              // N rows staged per CTA
  static constexpr int B_STAGE_BYTES = B_ROWS * BLOCK_K * 2;               // 32 / 16 KB
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int SLABS_PER_WARP = CG == 2 ? 2 : 1;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + NUM_EPI_WARPS * SLABS_PER_WARP * SLAB_BYTES +
                                    1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget exceeded");
};

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
  int aux_prefetch;    // 1: request aux with cp.async before waiting for the accumulator (CG = 2 only)
  Drop drop;           // EPI_BIAS_RESIDUAL only: D = dropout(acc + bias) + aux  (hidden dropout before the residual)
  uint32_t idesc;      // UMMA instruction descriptor (operand formats: bf16 or fp16 per operand)
  int dynamic;         // 1: work units are handed out by cluster launch control (see below); 0: static striding
  int aux_f16, out_f16;  // aux / D hold fp16 instead of bf16 (the encoder's fp16 residual stream)
  int save_pre;          // EPI_BIAS_GELU: out2 receives the pre-activation itself instead of gelu'(pre) (lean activations)
};

// ---- cluster / 2-CTA helpers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same smem offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
// TMA load issued by either CTA of a pair; completion bytes are credited to the LEADER's barrier (peer bit cleared)
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive (once the issued MMAs retire) on the barrier at this smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tma_store_wait_read_n(int slabs_in_flight) {
  if (slabs_in_flight == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}

// ---- dynamic persistent scheduling with cluster launch control (sm_100 CLC) ----------------------------------
// The grid has ONE cluster per work unit.  The clusters that get SMs process their own unit and then keep stealing
// units from clusters that have not been launched yet (clusterlaunchcontrol.try_cancel): a cancelled cluster never
// runs, its unit index (its first ctaid.x / 2) is executed by the thief.  Unlike a static `u += num_workers` loop this
// stays efficient when some SMs are busy with other kernels - NCCL's all-reduce CTAs during backward, the query
// encoder's kernels on the side stream - because nobody waits for a CTA that could not become resident: the measured
// cost of that wait was +2.8 ms of GEMM time per step at N = 2 (gemm_ms 61.0 -> 64.4).
// Protocol (2 response slots, so the next unit is requested while the current one is computed): warp 3 of EVERY CTA
// arms its CTA's clc_full[slot] (expect_tx 16) once its local consumers have released the slot and signals the
// leader; the leader's warp 3 then issues ONE multicast try_cancel whose 16-byte response lands in both CTAs.  The
// consumers (TMA thread, MMA thread, lane 0 of each epilogue warp) wait on clc_full, decode, release clc_empty.
__device__ __forceinline__ void clc_try_cancel_multicast(void* resp, uint64_t* bar) {
  asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.multicast::cluster::all.b128 [%0], [%1];"
               ::"r"(smem_u32(resp)), "r"(smem_u32(bar)) : "memory");
}
// returns the stolen cluster's first ctaid.x, or -1 when nothing was left to cancel
__device__ __forceinline__ int clc_decode(const void* resp) {
  uint64_t lo, hi;
  asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "r"(smem_u32(resp)) : "memory");
  uint32_t ok, x;
  asm volatile(
      "{\n\t.reg .b128 r;\n\t.reg .pred p;\n\t"
      "mov.b128 r, {%2, %3};\n\t"
      "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p, r;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "@p clusterlaunchcontrol.query_cancel.get_first_ctaid::x.b32.b128 %1, r;\n\t"
      "@!p mov.u32 %1, 0;\n\t}"
      : "=r"(ok), "=r"(x) : "l"(lo), "l"(hi));
  return ok ? (int)x : -1;
}
struct UnitCursor {       // one per consumer thread
  int slot;
  uint32_t phase;
};
// next work unit of this cluster (or -1): static striding, or the response of the pending CLC request
__device__ __forceinline__ int next_unit(int dynamic, int u, int num_workers, int units, UnitCursor& c, uint64_t* clc_full,
                                         uint64_t* clc_empty, const uint8_t* clc_resp) {
  if (!dynamic) {
    u += num_workers;
    return u < units ? u : -1;
  }
  mbar_wait(&clc_full[c.slot], c.phase);
  const int x = clc_decode(clc_resp + c.slot * 16);
  mbar_arrive(&clc_empty[c.slot]);
  if (++c.slot == 2) { c.slot = 0; c.phase ^= 1; }
  return x < 0 ? -1 : (x >> 1);
}

template <int A_MN, int B_MN, int CG>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_d, const __grid_constant__ CUtensorMap tmap_d2,
                 const GemmParams p) {
  using C = Cfg<CG>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B needs 1024-byte aligned tiles (the dynamic smem base is identical in both CTAs of a pair).
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * C::A_STAGE_BYTES;
  uint8_t* smem_stage = smem + STAGES * C::STAGE_BYTES;  // [NUM_EPI_WARPS][SLABS_PER_WARP][SLAB_BYTES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stage + NUM_EPI_WARPS * C::SLABS_PER_WARP * SLAB_BYTES);
  uint64_t* full_bar = bars;                       // [STAGES]   (CG=2: only the leader's are used)
  uint64_t* empty_bar = bars + STAGES;             // [STAGES]
  uint64_t* tmem_full_bar = bars + 2 * STAGES;     // [ACC_STAGES]
  uint64_t* tmem_empty_bar = tmem_full_bar + ACC_STAGES;  // (CG=2: only the leader's are used)
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + ACC_STAGES);
  uint64_t* clc_full = tmem_empty_bar + ACC_STAGES + 1;   // [2]
  uint64_t* clc_empty = clc_full + 2;                      // [2]
  uint64_t* clc_ready = clc_empty + 2;                     // [2] (leader's are used)
  uint8_t* clc_resp = reinterpret_cast<uint8_t*>(bars) + 208;   // [2][16], 16-byte aligned

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool is_leader = cta_rank == 0;
  const int worker = (CG == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;        // CTA or CTA pair index
  const int num_workers = (CG == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_d);
    tma_prefetch_desc(&tmap_d2);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], CG);   // one arrive per producer CTA (+ transaction bytes)
      mbar_init(&empty_bar[i], 1);   // tcgen05.commit (multicast to both CTAs when CG = 2)
    }
    for (int i = 0; i < ACC_STAGES; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], NUM_EPI_WARPS * CG);  // epilogue warps of both CTAs release the leader's MMA
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&clc_full[i], 1);                                         // the local arming arrive (+ 16 response bytes)
      mbar_init(&clc_empty[i], NUM_EPI_WARPS + 1 + (is_leader ? 1 : 0));  // local consumers: TMA, (MMA), epilogue warps
      mbar_init(&clc_ready[i], CG);                                       // both CTAs have armed their clc_full
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (CG == 2) tmem_alloc_2sm(tmem_base_slot, TMEM_COLS); else tmem_alloc(tmem_base_slot, TMEM_COLS);
  }
  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  const int tiles = p.num_m_blocks * p.num_n_blocks;
  const int units = tiles * p.splits;
  const int m_rows_per_unit = CTA_M * CG;

  if (warp == 0) {
    // ================================ TMA producer (one per CTA) ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      UnitCursor cur = {0, 0};
      for (int u = worker; u >= 0; u = next_unit(p.dynamic, u, num_workers, units, cur, clc_full, clc_empty, clc_resp)) {
        const int tile = u % tiles, split = u / tiles;
        const int m_blk = tile / p.num_n_blocks, n_blk = tile % p.num_n_blocks;
        const int m0 = m_blk * m_rows_per_unit + (int)cta_rank * CTA_M;        // this CTA's A rows
        const int n0 = n_blk * BLOCK_N + (int)cta_rank * C::B_ROWS;            // this CTA's share of B
        const int kb0 = split * p.k_blocks_per_split;
        const int kb1 = min(kb0 + p.k_blocks_per_split, p.k_blocks_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (is_leader) mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES * CG);
          uint8_t* sa = smem_a + stage * C::A_STAGE_BYTES;
          uint8_t* sb = smem_b + stage * C::B_STAGE_BYTES;
          auto load = [&](void* dst, const CUtensorMap* tm, int c0, int c1) {
            if (CG == 2) tma_load_2d_2sm(dst, tm, &full_bar[stage], c0, c1);
            else tma_load_2d(dst, tm, &full_bar[stage], c0, c1);
          };
          if (A_MN == 0) {
            load(sa, &tmap_a, kb * BLOCK_K, m0);
          } else {
#pragma unroll
            for (int i = 0; i < CTA_M / 64; ++i) load(sa + i * (BLOCK_K * 128), &tmap_a, m0 + i * 64, kb * BLOCK_K);
          }
          if (B_MN == 0) {
            load(sb, &tmap_b, kb * BLOCK_K, n0);
          } else {
#pragma unroll
            for (int i = 0; i < C::B_ROWS / 64; ++i) load(sb + i * (BLOCK_K * 128), &tmap_b, n0 + i * 64, kb * BLOCK_K);
          }
          if (CG == 2 && !is_leader) mbar_arrive_remote(&full_bar[stage], 0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================ MMA issuer (one thread of the leader CTA) ================================
    if (lane == 0 && is_leader) {
      const uint32_t idesc = p.idesc;
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
      UnitCursor cur = {0, 0};
      for (int u = worker; u >= 0; u = next_unit(p.dynamic, u, num_workers, units, cur, clc_full, clc_empty, clc_resp)) {
        const int split = u / tiles;
        const int kb0 = split * p.k_blocks_per_split;
        const int kb1 = min(kb0 + p.k_blocks_per_split, p.k_blocks_total);
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint64_t a_desc = make_umma_desc_sw128(smem_u32(smem_a + stage * C::A_STAGE_BYTES), A_LBO, A_SBO);
          const uint64_t b_desc = make_umma_desc_sw128(smem_u32(smem_b + stage * C::B_STAGE_BYTES), B_LBO, B_SBO);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint32_t accum = (kb > kb0 || k > 0) ? 1u : 0u;
            if (CG == 2) umma_f16_2sm(d_tmem, a_desc + (uint64_t)(k * A_KSTEP), b_desc + (uint64_t)(k * B_KSTEP), idesc, accum);
            else umma_f16(d_tmem, a_desc + (uint64_t)(k * A_KSTEP), b_desc + (uint64_t)(k * B_KSTEP), idesc, accum);
          }
          // frees the smem slot (in both CTAs) when these MMAs retire
          if (CG == 2) umma_commit_2sm(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue warps (of both CTAs)
        if (CG == 2) umma_commit_2sm(&tmem_full_bar[acc]); else umma_commit(&tmem_full_bar[acc]);
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // ================================ CLC scheduler (one thread per CTA) ================================
    if (lane == 0 && p.dynamic) {
      int slot = 0;
      uint32_t ph = 0;
      for (;;) {
        mbar_wait(&clc_empty[slot], ph ^ 1);                 // local consumers are done with this slot's old response
        mbar_arrive_expect_tx(&clc_full[slot], 16);          // arm: the response is 16 bytes
        if (CG == 2 && !is_leader) mbar_arrive_remote(&clc_ready[slot], 0);
        else mbar_arrive(&clc_ready[slot]);
        if (is_leader) {
          mbar_wait(&clc_ready[slot], ph);                   // both CTAs armed
          clc_try_cancel_multicast(clc_resp + slot * 16, &clc_full[slot]);
        }
        mbar_wait(&clc_full[slot], ph);
        const int x = clc_decode(clc_resp + slot * 16);
        if (x < 0) break;                                    // nothing left to steal: no further requests
        if (++slot == 2) { slot = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ================================ epilogue warps ================================
    // Warp (quarter, col_half) owns TMEM lanes [32*quarter, +32) x columns [128*col_half, +128) of this CTA's
    // 128x256 accumulator.  bf16 outputs: registers -> 128B-swizzled smem slab (32 rows x 64 cols) -> TMA store
    // (full-line writes; partial tiles are clipped by the tensor map).  aux (residual / pre-activation) is read with
    // coalesced 16-byte loads through the same slab.  fp32 outputs (split-K wgrad) go out as RED/ST from registers.
    const int ew = warp - 4;
    const int quarter = warp & 3;
    const int col_half = ew >> 2;
    constexpr int NSLAB = C::SLABS_PER_WARP;
    uint8_t* slab_base = smem_stage + ew * NSLAB * SLAB_BYTES;
    int slab_sel = 0;
    const bool f32_out = (p.epilogue == DPRB_EPI_F32_ATOMIC_ADD || p.epilogue == DPRB_EPI_F32_STORE);
    const bool has_aux = (p.epilogue == DPRB_EPI_BIAS_RESIDUAL || p.epilogue == DPRB_EPI_DGELU || p.epilogue == DPRB_EPI_DGELU_PRE);
    const bool is_gelu = (p.epilogue == DPRB_EPI_BIAS_GELU);
    int acc = 0;
    uint32_t acc_phase = 0;
    UnitCursor cur = {0, 0};
    for (int u = worker; u >= 0;) {
      const int tile = u % tiles;
      const int m_blk = tile / p.num_n_blocks, n_blk = tile % p.num_n_blocks;
      const int row0 = m_blk * m_rows_per_unit + (int)cta_rank * CTA_M + quarter * 32;
      const int row = row0 + lane;
      // aux (residual / pre-activation) does not depend on the accumulator: request both of this warp's slabs with
      // cp.async BEFORE waiting for the MMA, and pull the next tile's aux lines into L2, so the DRAM latency of the
      // epilogue operand hides behind the tensor-core work instead of sitting on the epilogue's critical path.
      const bool aux_pf = has_aux && NSLAB == 2 && p.aux_prefetch;
      if (aux_pf) {
        if (lane == 0) tma_store_wait_read_n(0);
        __syncwarp();
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
          const int colbase = n_blk * BLOCK_N + col_half * 128 + sl * 64;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int idx = lane + 32 * i, r = idx >> 3, ch = idx & 7;
            const bool ok = row0 + r < p.M && colbase + ch * 8 < p.N;
            const bf16* src = ok ? p.aux + (long long)(row0 + r) * p.ld_aux + colbase + ch * 8 : p.aux;
            cp_async_16_zfill(slab_base + sl * SLAB_BYTES + r * 128 + ((ch ^ (r & 7)) << 4), src, ok);
          }
        }
        cp_async_commit();
        const int un = p.dynamic ? units : u + num_workers;   // (dynamic: the next unit is not known yet)
        if (un < units) {
          const int tn = un % tiles;
          const int rn = (tn / p.num_n_blocks) * m_rows_per_unit + (int)cta_rank * CTA_M + quarter * 32 + lane;
          const int cn = (tn % p.num_n_blocks) * BLOCK_N + col_half * 128;
          if (rn < p.M && cn < p.N) {
            prefetch_l2(p.aux + (long long)rn * p.ld_aux + cn);
            if (cn + 64 < p.N) prefetch_l2(p.aux + (long long)rn * p.ld_aux + cn + 64);
          }
        }
      }
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();
      if (aux_pf) {
        cp_async_wait<0>();
        __syncwarp();
      }
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
        // single pass, two 32x32 slabs (64B swizzle) per 32-column chunk: gelu'(pre) -> out2 (optional), gelu(pre) -> D
#pragma unroll 1
        for (int h = 0; h < 4; ++h) {
          const int col0 = n_blk * BLOCK_N + col_half * 128 + h * 32;
          if (col0 >= p.N) break;  // warp-uniform
          uint32_t r[32];
          tmem_ld_32x32(tbase + h * 32, r);
          tmem_ld_wait();
          // pairs of adjacent accumulator columns go through the packed-fp32 pipe (FFMA2): bias, GELU and its derivative
          float2 v2[16];
          const float2 al = make_float2(p.alpha, p.alpha);
          if (p.bias != nullptr && col0 + 32 <= p.N) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
              v2[j >> 1] = __ffma2_rn(make_float2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), al, make_float2(b.x, b.y));
              v2[(j >> 1) + 1] = __ffma2_rn(make_float2(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])), al, make_float2(b.z, b.w));
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float b0 = (p.bias != nullptr && col0 + j < p.N) ? __ldg(p.bias + col0 + j) : 0.f;
              const float b1 = (p.bias != nullptr && col0 + j + 1 < p.N) ? __ldg(p.bias + col0 + j + 1) : 0.f;
              v2[j >> 1] = __ffma2_rn(make_float2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), al, make_float2(b0, b1));
            }
          }
          uint8_t* slab_pre = slab_base + slab_sel * SLAB_BYTES;
          uint8_t* slab_act = slab_pre + SLAB_BYTES / 2;
          if (lane == 0) tma_store_wait_read_n(NSLAB - 1);
          __syncwarp();
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            uint4 qp, qa;
            uint32_t* pp = &qp.x;
            uint32_t* pa = &qa.x;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              // out2 receives gelu'(pre): backward (EPI_DGELU) only ever needs the derivative, never pre itself
              float2 g, d;
              gelu_and_grad2(v2[c4 * 4 + t], g, d);
              if (p.save_pre) d = v2[c4 * 4 + t];        // lean activations: keep pre, rebuild gelu / gelu' in backward
              pp[t] = pack_bf16x2(d.x, d.y);
              pa[t] = pack_bf16x2(g.x, g.y);
            }
            const uint32_t off = lane * 64 + ((c4 ^ ((lane >> 1) & 3)) << 4);
            if (p.out2 != nullptr) *reinterpret_cast<uint4*>(slab_pre + off) = qp;
            *reinterpret_cast<uint4*>(slab_act + off) = qa;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (p.out2 != nullptr) tma_store_2d(&tmap_d2, smem_u32(slab_pre), col0, row0);
            tma_store_2d(&tmap_d, smem_u32(slab_act), col0, row0);
            tma_store_commit();
          }
          slab_sel = (slab_sel + 1) % NSLAB;
        }
        __syncwarp();
      } else {
#pragma unroll 1
        for (int sl = 0; sl < 2; ++sl) {          // two 64-column slabs per warp
          const int colbase = n_blk * BLOCK_N + col_half * 128 + sl * 64;
          if (colbase >= p.N) break;              // warp-uniform
          uint8_t* slab = slab_base + (aux_pf ? sl : slab_sel) * SLAB_BYTES;
          if (!aux_pf) {
            // the TMA store that last used this slab must have finished READING it before it is overwritten
            if (lane == 0) tma_store_wait_read_n(NSLAB - 1);
            __syncwarp();
          }
          if (has_aux && !aux_pf) {
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
            if (p.drop.on()) {
              // hidden dropout between the dense layer and the residual add (training only; kept out of the hot
              // loops below so the p = 0 path compiles exactly as before)
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                float2 m[4];
                p.drop.mul8((uint32_t)row, (uint32_t)(col0 + j), m);
#pragma unroll
                for (int w = 0; w < 4; ++w) { v[j + 2 * w] *= m[w].x; v[j + 2 * w + 1] *= m[w].y; }
              }
            }
            if (has_aux) {
#pragma unroll
              for (int c4 = 0; c4 < 4; ++c4) {
                const int ch = h * 4 + c4;
                const uint4 q = *reinterpret_cast<const uint4*>(slab + lane * 128 + ((ch ^ (lane & 7)) << 4));
                float2 f0, f1, f2, f3;
                if (p.aux_f16) { f0 = unpack_f16x2(q.x); f1 = unpack_f16x2(q.y); f2 = unpack_f16x2(q.z); f3 = unpack_f16x2(q.w); }
                else { f0 = unpack_bf16x2(q.x); f1 = unpack_bf16x2(q.y); f2 = unpack_bf16x2(q.z); f3 = unpack_bf16x2(q.w); }
                const float a[8] = {f0.x, f0.y, f1.x, f1.y, f2.x, f2.y, f3.x, f3.y};
                if (p.epilogue == DPRB_EPI_DGELU_PRE) {
                  // aux holds the pre-activation: derivative rebuilt here (same fitted function as the forward)
#pragma unroll
                  for (int t = 0; t < 8; t += 2) {
                    float2 g, d;
                    gelu_and_grad2(make_float2(a[t], a[t + 1]), g, d);
                    v[c4 * 8 + t] *= d.x; v[c4 * 8 + t + 1] *= d.y;
                  }
                } else {
#pragma unroll
                  for (int t = 0; t < 8; ++t) {
                    if (p.epilogue == DPRB_EPI_BIAS_RESIDUAL) v[c4 * 8 + t] += a[t];
                    else v[c4 * 8 + t] *= a[t];  // EPI_DGELU: aux holds gelu'(pre) written by the forward epilogue
                  }
                }
              }
            }
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              const int ch = h * 4 + c4;
              uint4 q;
              if (p.out_f16) {
                q.x = pack_f16x2(v[c4 * 8 + 0], v[c4 * 8 + 1]); q.y = pack_f16x2(v[c4 * 8 + 2], v[c4 * 8 + 3]);
                q.z = pack_f16x2(v[c4 * 8 + 4], v[c4 * 8 + 5]); q.w = pack_f16x2(v[c4 * 8 + 6], v[c4 * 8 + 7]);
              } else {
                q.x = pack_bf16x2(v[c4 * 8 + 0], v[c4 * 8 + 1]); q.y = pack_bf16x2(v[c4 * 8 + 2], v[c4 * 8 + 3]);
                q.z = pack_bf16x2(v[c4 * 8 + 4], v[c4 * 8 + 5]); q.w = pack_bf16x2(v[c4 * 8 + 6], v[c4 * 8 + 7]);
              }
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
            tma_store_2d(&tmap_d, smem_u32(slab), colbase, row0);
            tma_store_commit();
          }
          slab_sel = (slab_sel + 1) % NSLAB;
        }
        __syncwarp();
      }
      // all tcgen05.ld of this accumulator stage have completed (tmem_ld_wait): release it to the MMA warp
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2 && !is_leader) mbar_arrive_remote(&tmem_empty_bar[acc], 0);
        else mbar_arrive(&tmem_empty_bar[acc]);
      }
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      int nu = 0;
      if (lane == 0) nu = next_unit(p.dynamic, u, num_workers, units, cur, clc_full, clc_empty, clc_resp);
      u = __shfl_sync(0xFFFFFFFFu, nu, 0);
    }
    if (lane == 0) tma_store_wait_all();  // smem must stay valid until the last bulk store has drained
    __syncwarp();
  }

  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    if (CG == 2) tmem_dealloc_2sm(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
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
  static thread_local bool ctx_bound = false;   // driver entry point: needs a current context on THIS thread
  if (!ctx_bound) {
    DPRB_CHECK_CUDA(cudaFree(nullptr));
    ctx_bound = true;
  }
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
              long long ld_aux, void* out2, float alpha, int splits, float* colsum, float dropout_p,
              unsigned long long drop_site_seed, cudaStream_t stream) {
  DPRB_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  const int dt_flags = epilogue & ~0xFF;   // DPRB_GEMM_{A,B,AUX,OUT}_F16, DPRB_GEMM_SAVE_PRE
  epilogue &= 0xFF;
  const int a_f16 = (dt_flags & DPRB_GEMM_A_F16) != 0, b_f16 = (dt_flags & DPRB_GEMM_B_F16) != 0;
  const int aux_f16 = (dt_flags & DPRB_GEMM_AUX_F16) != 0, out_f16 = (dt_flags & DPRB_GEMM_OUT_F16) != 0;
  DPRB_REQUIRE(epilogue >= 0 && epilogue < DPRB_EPI_COUNT, "gemm: bad epilogue %d", epilogue);
  DPRB_REQUIRE(!(aux_f16 || out_f16) || epilogue == DPRB_EPI_BIAS || epilogue == DPRB_EPI_BIAS_RESIDUAL,
               "gemm: fp16 aux / output is implemented for the BIAS and BIAS_RESIDUAL epilogues only");
  DPRB_REQUIRE(a_f16 == b_f16, "gemm: tcgen05 kind::f16 rejects an fp16 x bf16 operand pair (illegal instruction): give "
               "DPRB_GEMM_A_F16 and DPRB_GEMM_B_F16 together or not at all");
  DPRB_REQUIRE(!out_f16 || colsum == nullptr, "gemm: colsum reads a bf16 slab (not available with fp16 output)");
  const bool f32_out = (epilogue == DPRB_EPI_F32_ATOMIC_ADD || epilogue == DPRB_EPI_F32_STORE);
  DPRB_REQUIRE(f32_out || (ldd % 8 == 0 && (reinterpret_cast<uintptr_t>(D) & 15) == 0),
               "gemm: bf16 output must be 16-byte aligned with ldd %% 8 == 0 (ldd=%lld)", ldd);
  DPRB_REQUIRE(!f32_out || (ldd % 4 == 0 && (reinterpret_cast<uintptr_t>(D) & 15) == 0),
               "gemm: fp32 output must be 16-byte aligned with ldd %% 4 == 0 (ldd=%lld)", ldd);
  if (epilogue == DPRB_EPI_BIAS_RESIDUAL || epilogue == DPRB_EPI_DGELU || epilogue == DPRB_EPI_DGELU_PRE)
    DPRB_REQUIRE(aux != nullptr && ld_aux % 8 == 0, "gemm: epilogue %d needs aux with ld %% 8 == 0", epilogue);
  if (bias != nullptr) DPRB_REQUIRE((reinterpret_cast<uintptr_t>(bias) & 15) == 0, "gemm: bias not 16B aligned");

  CUtensorMap ta, tb;
  int rc;
  static const bool one_cta = (std::getenv("DPRB_GEMM_1CTA") != nullptr);
  const int CG = one_cta ? 1 : 2;
  if (!a_mn_major) rc = make_tmap(&ta, A, M, K, lda, CTA_M); else rc = make_tmap(&ta, A, K, M, lda, BLOCK_K);
  if (rc) return rc;
  if (!b_mn_major) rc = make_tmap(&tb, B, N, K, ldb, BLOCK_N / CG); else rc = make_tmap(&tb, B, K, N, ldb, BLOCK_K);
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
  p.num_m_blocks = (M + CTA_M * CG - 1) / (CTA_M * CG);
  p.num_n_blocks = (N + BLOCK_N - 1) / BLOCK_N;
  p.k_blocks_total = (K + BLOCK_K - 1) / BLOCK_K;
  const int sms = num_sms();
  const int workers = CG == 2 ? sms / 2 : sms;  // CTAs or CTA pairs
  const int tiles = p.num_m_blocks * p.num_n_blocks;
  if (epilogue != DPRB_EPI_F32_ATOMIC_ADD) splits = 1;
  else if (splits <= 0) splits = choose_splits(tiles, p.k_blocks_total, workers);
  if (splits > p.k_blocks_total) splits = p.k_blocks_total;
  p.k_blocks_per_split = (p.k_blocks_total + splits - 1) / splits;
  p.splits = (p.k_blocks_total + p.k_blocks_per_split - 1) / p.k_blocks_per_split;
  p.epilogue = epilogue;
  p.D = D; p.ldd = ldd; p.bias = bias; p.aux = reinterpret_cast<const bf16*>(aux); p.ld_aux = ld_aux;
  p.out2 = reinterpret_cast<bf16*>(out2); p.alpha = alpha;
  p.colsum = colsum;
  p.aux_f16 = aux_f16; p.out_f16 = out_f16;
  p.save_pre = (dt_flags & DPRB_GEMM_SAVE_PRE) != 0;
  p.idesc = make_idesc_16_f32(CTA_M * CG, BLOCK_N, a_mn_major ? 1 : 0, b_mn_major ? 1 : 0, a_f16, b_f16);
  p.drop = drop_from_site(epilogue == DPRB_EPI_BIAS_RESIDUAL ? dropout_p : 0.f, drop_site_seed);
  static const bool no_aux_pf = (std::getenv("DPRB_NO_AUX_PF") != nullptr);
  // measured (same box, cfg-2 shapes): prefetching aux before the accumulator wait gains 10-14 % where the epilogue
  // is the critical path (K <= 1024: attention-out, dGELU) and costs ~1.5 % on the K >= 2304 GEMMs
  p.aux_prefetch = (no_aux_pf || K > 1024) ? 0 : 1;
  DPRB_REQUIRE(colsum == nullptr || (!f32_out && epilogue != DPRB_EPI_BIAS_GELU),
               "gemm: colsum is supported for the BIAS / BIAS_RESIDUAL / DGELU epilogues only");

  const int units = tiles * p.splits;
  static const bool static_sched = (std::getenv("DPRB_GEMM_STATIC") != nullptr);
  p.dynamic = (CG == 2 && !static_sched && units > workers) ? 1 : 0;
  const int grid = p.dynamic ? units * CG : (units < workers ? units : workers) * CG;

  static bool attr_set = false;
  if (!attr_set) {
#define DPRB_SET_ATTR(K, CGV)                                                                                           \
  DPRB_CHECK_CUDA(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<CGV>::SMEM_BYTES));
    DPRB_SET_ATTR((gemm_bf16_kernel<0, 0, 1>), 1) DPRB_SET_ATTR((gemm_bf16_kernel<0, 1, 1>), 1)
    DPRB_SET_ATTR((gemm_bf16_kernel<1, 0, 1>), 1) DPRB_SET_ATTR((gemm_bf16_kernel<1, 1, 1>), 1)
    DPRB_SET_ATTR((gemm_bf16_kernel<0, 0, 2>), 2) DPRB_SET_ATTR((gemm_bf16_kernel<0, 1, 2>), 2)
    DPRB_SET_ATTR((gemm_bf16_kernel<1, 0, 2>), 2) DPRB_SET_ATTR((gemm_bf16_kernel<1, 1, 2>), 2)
#undef DPRB_SET_ATTR
    attr_set = true;
  }
  auto launch = [&](auto kern, int smem_bytes) -> int {
    const bool prof = g_prof.enabled && g_prof.used + 2 <= g_prof.ev.size();
    if (prof) DPRB_CHECK_CUDA(cudaEventRecord(g_prof.ev[g_prof.used], stream));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    count_launch();
    DPRB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, td, td2, p));
    if (prof) {
      DPRB_CHECK_CUDA(cudaEventRecord(g_prof.ev[g_prof.used + 1], stream));
      g_prof.flops.push_back(2.0 * (double)M * (double)N * (double)K);
      g_prof.used += 2;
    }
    return 0;
  };
  const int key = (a_mn_major ? 2 : 0) | (b_mn_major ? 1 : 0);
  if (CG == 2) {
    switch (key) {
      case 0: return launch(gemm_bf16_kernel<0, 0, 2>, Cfg<2>::SMEM_BYTES);
      case 1: return launch(gemm_bf16_kernel<0, 1, 2>, Cfg<2>::SMEM_BYTES);
      case 2: return launch(gemm_bf16_kernel<1, 0, 2>, Cfg<2>::SMEM_BYTES);
      default: return launch(gemm_bf16_kernel<1, 1, 2>, Cfg<2>::SMEM_BYTES);
    }
  }
  switch (key) {
    case 0: return launch(gemm_bf16_kernel<0, 0, 1>, Cfg<1>::SMEM_BYTES);
    case 1: return launch(gemm_bf16_kernel<0, 1, 1>, Cfg<1>::SMEM_BYTES);
    case 2: return launch(gemm_bf16_kernel<1, 0, 1>, Cfg<1>::SMEM_BYTES);
    default: return launch(gemm_bf16_kernel<1, 1, 1>, Cfg<1>::SMEM_BYTES);
  }
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

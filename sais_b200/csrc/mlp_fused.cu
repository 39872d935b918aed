// Fused ViT MLP for sm_100a:   x += fc2( GELU( fc1(xn) + b1 ) ) + b2      (hidden activations never leave the SM)
//
// Reference: Mlp.forward + the residual add of Block.forward, SAIS/scripts/dino-main/vision_transformer.py:56-65,107.
//
// Why fused: at dim 384 the two GEMMs are bound by the memory hierarchy, not by the tensor pipe — the unfused pair
// writes and re-reads the [rows,1536] hidden matrix through HBM (6,144 B/row) and pulls 128x256-tile operands
// through L2 at ~2x the ~45 B/clk/SM the L2 can deliver.  Here a CTA PAIR (2-CTA cluster, tcgen05 cta_group::2) owns
// 256 rows: the LayerNorm'd rows A[256,384] stay resident in shared memory, the hidden dimension is walked in
// chunks of 64, and per chunk j
//     G1(j):  S_j[256,64]   = A · W1[j]ᵀ                 (24 MMAs 256x64x16,  accumulator S in TMEM, double buffered)
//     E1(j):  H_j           = bf16(GELU(S_j + b1[j]))     (epilogue warps, TMEM -> registers -> swizzled smem tile)
//     G2(j):  acc[256,384] += H_j · W2[:, j]ᵀ             (8 MMAs 256x192x16, accumulator acc in TMEM)
// with the tensor pipe running G1(j+1) while the epilogue warps do E1(j).  Only the weights stream (each CTA loads
// half of every W1/W2 chunk; the pair exchanges operand halves in hardware), 48 KB per chunk per CTA.  The finished
// accumulator (+ b2) is added to the fp32 residual stream IN L2 by a TMA reduce-add store, so the residual is never
// loaded into the SM at all.
//
// With ln_stats / ln_colsum the first epilogue also applies the block's norm2 (folded LayerNorm, gemm_tcgen05.cu): A then
// holds the RAW bf16 rows of the residual stream and W1 / b1 the gamma-scaled weights / folded bias.
//
//   warps 0..7  : E1 and the output epilogue (two warps per TMEM lane quarter)
//   warps 8..9  : cast warps (optional outputs xb_out / stats_out, below)
//   warp 10     : TMA producer (A tile, then the W1/W2 chunk ring in exactly the order the MMAs consume it)
//   warp 11     : TMEM allocator + MMA issuer (leader CTA of the pair only)
// Both role warps walk their loops with all 32 lanes (warp-uniform control flow) and elect one lane only around the TMA /
// tcgen05 instructions: issued from a divergent single lane every MMA cost ~200 cycles of issue overhead (161 -> 115 us).
// TMEM columns: acc [0,384) | S0 [384,448) | S1 [448,512).
// Work units: 256-row tiles, dealt round-robin to the 74 CTA pairs.  Every output element receives exactly one reduce-add,
// so the result is deterministic (an earlier version split the tiles of the last partial round along the hidden
// dimension: +1.4 % throughput, but partial sums then met in L2 in arbitrary order — run-to-run bit differences; removed).
//
// Cast warps (xb_out / stats_out): the NEXT block's LayerNorm-folded qkv GEMM consumes the bf16 copy and the per-row
// (sum, sum of squares) of the UPDATED residual stream, which a reduce-add kernel never holds in registers.  Instead of a
// separate pass over the stream (rowstats_cast384_kernel: 20 us per block at batch 256, 5.8 % of the step), two otherwise
// idle warps per CTA re-read the rows of the unit that has just been reduced into L2 (L2 hits, ld.global.cg) while the
// tensor pipe works on the next unit, and write the bf16 copy + statistics with the very same code
// (rowcast.cuh: bit-identical to the stand-alone kernel).  Ordering: an epilogue warp's reduce-adds are complete
// (cp.async.bulk.wait_group 0 + fence.proxy.async) before its arrival on cast_full; the wait is deferred to the second hidden
// chunk of the following unit so that it never stalls the E1 pipeline.  The last unit of a CTA is cast by the eight
// epilogue warps together (nothing is left to overlap it with).
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "rowcast.cuh"

namespace sais {

namespace {

constexpr int MT = 128;             // rows per CTA (256 per pair)
constexpr int DM = 384;             // model dim (K of fc1, N of fc2)
constexpr int HID = 1536;           // hidden dim
constexpr int HC = 64;              // hidden chunk (N of G1, K of G2)
constexpr int NCHUNK = HID / HC;    // 24
constexpr int KB = DM / 64;         // 6 k-blocks of A
constexpr int kEpiWarps = 8;
constexpr int kCastWarps = 2;
constexpr int kThreads = 32 * (kEpiWarps + kCastWarps + 2);
constexpr int A_BYTES = KB * MT * 128;    // 98,304: six 128-row x 128-byte swizzled k-blocks
constexpr int STAGE_BYTES = 24576;        // one W1 chunk half (6 x 32 rows x 128 B) or one W2 chunk half (2 x 96 rows x 128 B)
constexpr int NSTAGE = 4;
constexpr int H_BYTES = MT * 128;         // 16,384: H_j tile, 128 rows x 64 bf16
constexpr int kPfFirst = 6;               // hidden chunk at which the producer starts prefetching the unit's residual rows
constexpr int kTmemCols = 512;
constexpr int kSCol = DM;                 // S0 at 384, S1 at 448
constexpr int kTailBytes = 256 /*barriers*/ + DM * 4 /*b2*/;
constexpr int kSmemBytes = 1024 + A_BYTES + NSTAGE * STAGE_BYTES + 2 * H_BYTES + kTailBytes;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

struct MlpParams {
  const float* b1;
  const float* b2;
  // folded LayerNorm (gemm_tcgen05.cu, "LayerNorm folding"): xn holds the RAW bf16 rows, w1 = gamma-scaled weights, b1 = d;
  // the first epilogue applies rstd * (S - mean * c) + d per row from the producer's (sum, sum of squares) partials
  const float* ln_stats;   // [rows][4][2] or nullptr
  const float* ln_colsum;  // [1536] c_n
  float ln_eps;
  int64_t rows;
  int reverse;      // walk the row tiles from the last to the first (kernels.h g_tile_reverse)
  int num_tiles;    // 256-row tiles = work units
  // cast warps: bf16 copy + row statistics of the UPDATED stream (both or neither; may alias xn / ln_stats)
  const float* x;
  __nv_bfloat16* xb_out;
  float* stats_out;
  long long* dbg;
};

__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, uint32_t smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
// bring a box of a tensor into L2 without touching shared memory (the x tile the unit's reduce-adds will hit)
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* m, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
               :
               : "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void sts128m(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory"); }

__device__ __forceinline__ int unit_tile(const MlpParams& p, int u) { return p.reverse ? p.num_tiles - 1 - u : u; }
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w1,
                 const __grid_constant__ CUtensorMap tmap_w2, const __grid_constant__ CUtensorMap tmap_out,
                 const __grid_constant__ CUtensorMap tmap_xpf, const MlpParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_smem = smem;
  uint8_t* w_smem = smem + A_BYTES;
  uint8_t* h_smem = w_smem + NSTAGE * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(h_smem + 2 * H_BYTES);
  uint64_t* w_full = bars;            // [NSTAGE] (leader's copy is the one that counts)
  uint64_t* w_empty = bars + NSTAGE;  // [NSTAGE]
  uint64_t* a_full = bars + 8;
  uint64_t* a_empty = bars + 9;
  uint64_t* s_full = bars + 10;   // [2]
  uint64_t* s_empty = bars + 12;  // [2] leader's, 16 arrivals
  uint64_t* h_full = bars + 14;   // [2] leader's, 16 arrivals
  uint64_t* h_empty = bars + 16;  // [2]
  uint64_t* acc_full = bars + 18;
  uint64_t* acc_empty = bars + 19;  // leader's, 16 arrivals
  uint64_t* cast_full = bars + 20;  // this CTA's, kEpiWarps arrivals: the unit's reduce-adds have completed
  uint64_t* cast_done = bars + 21;  // this CTA's, kCastWarps arrivals: its rows have been cast
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 22);
  float* b2_smem = reinterpret_cast<float*>(bars + 32);  // [384]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;
  auto stamp = [&](int role, int idx, int ev) {
    if (p.dbg != nullptr && blockIdx.x == 0 && idx < 32 && ev < 4) p.dbg[(role * 32 + idx) * 4 + ev] = clock64();
  };

  constexpr int kCastWarp0 = kEpiWarps, kProducerWarp = kEpiWarps + kCastWarps, kMmaWarp = kProducerWarp + 1;
  const bool do_cast = p.xb_out != nullptr;
  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_w1);
    tma_prefetch_desc(&tmap_w2);
    tma_prefetch_desc(&tmap_out);
    tma_prefetch_desc(&tmap_xpf);
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 2 * kEpiWarps);
      mbar_init(&h_full[s], 2 * kEpiWarps);
      mbar_init(&h_empty[s], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 2 * kEpiWarps);
    mbar_init(cast_full, kEpiWarps);
    mbar_init(cast_done, kCastWarps);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc_cg2(tmem_base_smem, kTmemCols);
    tmem_relinquish_cg2();
  }
  // PDL (common.cuh): the set-up above overlapped the previous kernel's tail; no global access before this line
  pdl_wait();
  for (int i = threadIdx.x; i < DM; i += kThreads) b2_smem[i] = __ldg(p.b2 + i);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  if (warp == kProducerWarp) {
    // ===================== TMA producer (whole warp, warp-uniform; TMA instructions elected) =====================
    {
      int stage = 0;
      uint32_t wphase = 0;
      uint32_t ui = 0;
      int pg = 0, pg2 = 0;  // chunk counters (timeline only)
      auto load_w1 = [&](int j) {
        mbar_wait(&w_empty[stage], wphase ^ 1);
        if (lane == 0) stamp(3, pg, 0);
        const uint32_t lbar = leader_smem_u32(&w_full[stage]);
        uint8_t* dst = w_smem + stage * STAGE_BYTES;
        if (elect_one()) {
          if (crank == 0) mbar_arrive_expect_tx(&w_full[stage], 2 * STAGE_BYTES);
#pragma unroll
          for (int kb = 0; kb < KB; ++kb)
            tma_load_2d_cg2(dst + kb * 4096, &tmap_w1, lbar, kb * 64, j * HC + int(crank) * (HC / 2));
        }
        __syncwarp();
        ++pg;
        if (++stage == NSTAGE) { stage = 0; wphase ^= 1; }
      };
      auto load_w2 = [&](int j) {
        mbar_wait(&w_empty[stage], wphase ^ 1);
        if (lane == 0) stamp(3, pg2, 1);
        const uint32_t lbar = leader_smem_u32(&w_full[stage]);
        uint8_t* dst = w_smem + stage * STAGE_BYTES;
        if (elect_one()) {
          if (crank == 0) mbar_arrive_expect_tx(&w_full[stage], 2 * STAGE_BYTES);
#pragma unroll
          for (int h = 0; h < 2; ++h)
            tma_load_2d_cg2(dst + h * 12288, &tmap_w2, lbar, j * HC, h * 192 + int(crank) * 96);
        }
        __syncwarp();
        ++pg2;
        if (++stage == NSTAGE) { stage = 0; wphase ^= 1; }
      };
      for (int u = pair; u < p.num_tiles; u += npairs, ++ui) {
        const int m0 = unit_tile(p, u) * (2 * MT) + int(crank) * MT;
        mbar_wait(a_empty, (ui & 1) ^ 1);  // the previous unit's G1s have retired
        if (elect_one()) {
          if (crank == 0) mbar_arrive_expect_tx(a_full, 2 * A_BYTES);
          const uint32_t lbar = leader_smem_u32(a_full);
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) tma_load_2d_cg2(a_smem + kb * (MT * 128), &tmap_a, lbar, kb * 64, m0);
        }
        __syncwarp();
        for (int jj = 0; jj <= NCHUNK; ++jj) {
          if (jj < NCHUNK) load_w1(jj);
          if (jj >= 1) load_w2(jj - 1);
          // L2 prefetch of this unit's 128 x 384 fp32 residual rows, one 128 x 32 box per hidden chunk in the middle of the
          // unit.  All CTAs reach their output epilogue at about the same time (equal units, lock step): without this the
          // reduce-adds of 148 CTAs (28 MB) all miss L2 at once and the epilogue runs at HBM speed (measured 15-20 k cycles
          // against ~9 k for the bytes it moves through the SM's L2 port); prefetched, the reads are spread over the hidden
          // loop and the reduce-adds hit.
          if (jj >= kPfFirst && jj < kPfFirst + DM / 32) {
            if (elect_one()) tma_prefetch_l2_2d(&tmap_xpf, (jj - kPfFirst) * 32, m0);
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (leader CTA) =====================
    // The whole warp walks the loop (warp-uniform control flow, descriptors in uniform registers); only the tcgen05
    // instructions sit under elect_one() — issued from one divergent lane every MMA cost a ~200-cycle ELECT / R2UR
    // waterfall (profiles/r01f_mma_issue.md), and this kernel issues 32 small MMAs per hidden chunk.
    if (crank == 0) {
      constexpr uint32_t idesc1 = umma_idesc_bf16(2 * MT, HC);
      constexpr uint32_t idesc2 = umma_idesc_bf16(2 * MT, 192);
      int stage = 0;
      uint32_t wphase = 0;
      uint32_t g = 0;  // chunks issued so far by this pair (selects S / H buffers and their barrier phases)
      uint32_t ui = 0;
      const uint32_t a_s = smem_u32(a_smem);
      const uint32_t h_s = smem_u32(h_smem);
      for (int u = pair; u < p.num_tiles; u += npairs, ++ui) {
        mbar_wait(a_full, ui & 1);
        tc_fence_after();
        if (lane == 0) stamp(2, ui, 0);
        for (int jj = 0; jj <= NCHUNK; ++jj) {
          if (jj < NCHUNK) {  // ---- G1: S[b] = A · W1[j]ᵀ
            const uint32_t gg = g + jj, b = gg & 1;
            mbar_wait(&s_empty[b], ((gg >> 1) & 1) ^ 1);
            if (lane == 0) stamp(0, gg, 0);
            mbar_wait(&w_full[stage], wphase);
            tc_fence_after();
            if (lane == 0) stamp(0, gg, 1);
            const uint32_t w_s = smem_u32(w_smem + stage * STAGE_BYTES);
            const uint32_t d = tmem_base + kSCol + b * HC;
            if (elect_one()) {
#pragma unroll
              for (int kb = 0; kb < KB; ++kb) {
                const uint64_t da = umma_desc_sw128_kmajor(a_s + kb * (MT * 128));
                const uint64_t db = umma_desc_sw128_kmajor(w_s + kb * 4096);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_cg2(d, da + 2 * k, db + 2 * k, idesc1, (kb | k) != 0);
              }
              umma_commit_cg2_mcast(&w_empty[stage], uint16_t(0b11));
              if (jj == NCHUNK - 1) umma_commit_cg2_mcast(a_empty, uint16_t(0b11));
              umma_commit_cg2_mcast(&s_full[b], uint16_t(0b11));
            }
            __syncwarp();
            if (++stage == NSTAGE) { stage = 0; wphase ^= 1; }
          }
          if (jj >= 1) {  // ---- G2: acc += H[b] · W2[:, j]ᵀ
            const uint32_t gg = g + jj - 1, b = gg & 1;
            if (jj == 1) {
              mbar_wait(acc_empty, (ui & 1) ^ 1);  // the previous unit's output epilogue has drained acc
              if (lane == 0) stamp(2, ui, 1);
            }
            mbar_wait(&h_full[b], (gg >> 1) & 1);
            if (lane == 0) stamp(0, gg, 2);
            mbar_wait(&w_full[stage], wphase);
            tc_fence_after();
            if (lane == 0) stamp(0, gg, 3);
            const uint32_t w_s = smem_u32(w_smem + stage * STAGE_BYTES);
            const uint64_t da = umma_desc_sw128_kmajor(h_s + b * H_BYTES);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const uint64_t db = umma_desc_sw128_kmajor(w_s + h * 12288);
                  umma_f16_cg2(tmem_base + h * 192, da + 2 * k, db + 2 * k, idesc2, (jj > 1) || (k > 0));
                }
              }
              umma_commit_cg2_mcast(&w_empty[stage], uint16_t(0b11));
              umma_commit_cg2_mcast(&h_empty[b], uint16_t(0b11));
              if (jj == NCHUNK) umma_commit_cg2_mcast(acc_full, uint16_t(0b11));
            }
            __syncwarp();
            if (++stage == NSTAGE) { stage = 0; wphase ^= 1; }
          }
        }
        g += NCHUNK;
      }
    }
  } else if (warp >= kCastWarp0) {
    // ===================== cast warps: bf16 copy + row statistics of the unit reduced one unit ago =====================
    if (do_cast) {
      const int cw = warp - kCastWarp0;
      constexpr int kRowsPerCastWarp = MT / kCastWarps;
      uint32_t ui = 0;
      for (int u = pair; u < p.num_tiles; u += npairs, ++ui) {
        if (u + npairs >= p.num_tiles) break;  // the last unit is cast by the epilogue warps themselves
        const int64_t r0 = int64_t(unit_tile(p, u)) * (2 * MT) + int(crank) * MT + cw * kRowsPerCastWarp;
        mbar_wait(cast_full, ui & 1);
#pragma unroll 1
        for (int i = 0; i < kRowsPerCastWarp; i += 4) {
          const int64_t left = p.rows - (r0 + i);
          if (left <= 0) break;
          rowcast_rows<4, true>(p.x, p.xb_out, p.stats_out, r0 + i, 1, left < 4 ? int(left) : 4, lane);
          // paced: 16 batches of 4 rows spread over most of a unit (~28 us).  Back to back they pull 192 KB + push 96 KB
          // through the SM's L2 port within a few microseconds, on top of the operand stream — measured as a 15 % slower
          // hidden loop while the cast ran.
          __nanosleep(1000);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(cast_done);
      }
    }
  } else {
    // ===================== epilogue warps 0..7 =====================
    const int ew = warp;
    const int q = ew & 3;      // TMEM lane quarter
    const int half = ew >> 2;  // column half (E1: 32 of 64; output: 192 of 384)
    const int row = q * 32 + lane;
    const uint32_t t_lane = tmem_base + (uint32_t(q * 32) << 16);
    const uint32_t h_row = smem_u32(h_smem) + row * 128;
    const int sw = row & 7;
    const uint32_t my_stage = smem_u32(h_smem) + ew * 4096;  // output staging: 32 rows x 128 B (swizzled), aliasing H
    uint32_t g = 0, ui = 0;
    bool cast_pending = false;  // the previous unit's reduce-adds have not been confirmed complete yet
    for (int u = pair; u < p.num_tiles; u += npairs, ++ui) {
      const int m0 = unit_tile(p, u) * (2 * MT) + int(crank) * MT;
      // h = x / 2 is produced directly (bias, rstd and -mean * rstd halved: exact scalings) for gelu_erf_fast2_half;
      // without the folded LayerNorm rstd = 1, mean = 0 and the column-sum pointer aliases the bias
      float rs_h = 0.5f, nm_h = 0.0f;
      if (p.ln_stats != nullptr) {
        const int64_t r = int64_t(m0) + row;
        const float4* sp = reinterpret_cast<const float4*>(p.ln_stats + (r < p.rows ? r : p.rows - 1) * 8);
        const float4 s0 = __ldg(sp), s1 = __ldg(sp + 1);
        const float mean = ((s0.x + s0.z) + (s1.x + s1.z)) * (1.0f / float(DM));
        const float var = fmaxf(((s0.y + s0.w) + (s1.y + s1.w)) * (1.0f / float(DM)) - mean * mean, 0.0f);
        const float rstd = rsqrtf(var + p.ln_eps);
        rs_h = 0.5f * rstd;
        nm_h = -mean * rstd * 0.5f;
      }
      const uint64_t rs2 = pack2(rs_h, rs_h), nm2 = pack2(nm_h, nm_h);
      const float* cs_base = p.ln_colsum != nullptr ? p.ln_colsum : p.b1;
      for (int jj = 0; jj < NCHUNK; ++jj) {
        const uint32_t gg = g + jj, b = gg & 1;
        const int j = jj;
        if (jj == 2 && cast_pending) {
          // hand the PREVIOUS unit's rows to the cast warps: by now (two hidden chunks later) this warp's reduce-adds have
          // long completed, so the wait does not stall E1
          if (lane == 0) {
            tma_store_wait<0>();
            fence_proxy_async_global();
            if (ui >= 2) mbar_wait(cast_done, ui & 1);  // (never blocks in practice: the cast of unit ui - 2 is a unit old)
            mbar_arrive(cast_full);
          }
          __syncwarp();
          cast_pending = false;
        }
        float4 bias[8], cs[8];
        {
          const float4* bp = reinterpret_cast<const float4*>(p.b1 + j * HC + half * 32);
          const float4* cp = reinterpret_cast<const float4*>(cs_base + j * HC + half * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            bias[i] = __ldg(bp + i);
            cs[i] = __ldg(cp + i);
          }
        }
        mbar_wait(&s_full[b], (gg >> 1) & 1);
        tc_fence_after();
        if (ew == 0 && lane == 0) stamp(1, gg, 0);
        uint32_t v[32];
        tmem_ld_32x32(t_lane + kSCol + b * HC + half * 32, v);
        tmem_ld_wait_dep(v);
        tc_fence_before();
        if (lane == 0) mbar_arrive_cluster(leader_smem_u32(&s_empty[b]));  // S[b] may be overwritten by G1(gg + 2)
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float f0, f1, f2, f3;
          unpack2(fma2(pack2u(v[4 * i], v[4 * i + 1]), rs2,
                       fma2(nm2, pack2(cs[i].x, cs[i].y), pack2(0.5f * bias[i].x, 0.5f * bias[i].y))), f0, f1);
          unpack2(fma2(pack2u(v[4 * i + 2], v[4 * i + 3]), rs2,
                       fma2(nm2, pack2(cs[i].z, cs[i].w), pack2(0.5f * bias[i].z, 0.5f * bias[i].w))), f2, f3);
          gelu_erf_fast2_half(f0, f1);
          gelu_erf_fast2_half(f2, f3);
          pk[2 * i] = pack_bf16x2(f0, f1);
          pk[2 * i + 1] = pack_bf16x2(f2, f3);
        }
        mbar_wait(&h_empty[b], ((gg >> 1) & 1) ^ 1);  // G2(gg - 2) has finished reading H[b]
        if (ew == 0 && lane == 0) stamp(1, gg, 1);
        const uint32_t hb = h_row + b * H_BYTES;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          sts128m(hb + (((half * 4 + i) ^ sw) << 4), pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(leader_smem_u32(&h_full[b]));
        if (ew == 0 && lane == 0) stamp(1, gg, 2);
      }
      g += NCHUNK;

      // ---- output epilogue: acc (+ b2) -> swizzled staging -> TMA reduce-add into the residual stream
      mbar_wait(acc_full, ui & 1);
      tc_fence_after();
      if (ew == 0 && lane == 0) stamp(2, ui, 2);
      constexpr bool add_bias = true;
      constexpr int NOC = (DM / 2) / 16;  // 12 chunks of 16 columns per warp, handled two at a time:
      // both accumulator chunks are moved to registers and biased BEFORE waiting for the previous iteration's two stores to
      // have read the staging tiles, so that wait (the TMA queue is busy with the next unit's operand loads) overlaps the
      // math; one fence / commit per 32 columns.  The output epilogue is fully exposed.
      uint32_t v[16], w[16];
      tmem_ld_32x16(t_lane + half * 192, v);
      tmem_ld_32x16(t_lane + half * 192 + 16, w);
#pragma unroll 1
      for (int c = 0; c < NOC; c += 2) {
        const int col = half * 192 + c * 16;
        tmem_ld_wait_dep(v);
        tmem_ld_wait_dep(w);
        float f[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          f[i] = __uint_as_float(v[i]);
          f[16 + i] = __uint_as_float(w[i]);
        }
        if (c + 2 < NOC) {
          tmem_ld_32x16(t_lane + col + 32, v);
          tmem_ld_32x16(t_lane + col + 48, w);
        } else {
          tc_fence_before();
          if (lane == 0) mbar_arrive_cluster(leader_smem_u32(acc_empty));  // acc fully read by this warp
        }
        if (add_bias) {
          const float4* bp = reinterpret_cast<const float4*>(b2_smem + col);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 b4 = bp[i];
            f[4 * i] += b4.x; f[4 * i + 1] += b4.y; f[4 * i + 2] += b4.z; f[4 * i + 3] += b4.w;
          }
        }
        if (lane == 0) tma_store_wait_read<0>();  // the previous iteration's store has read the staging tile
        __syncwarp();
        // one 32-row x 128-byte box (128-byte swizzle) per 32 columns: half as many TMA requests as two 64-byte-wide boxes
        // — the drain was paced by the TMA unit's request rate (192 KB in ~16 k cycles = 12 B/clk), not by the L2 port
#pragma unroll
        for (int i = 0; i < 8; ++i)
          sts128m(my_stage + lane * 128 + ((i ^ (lane & 7)) << 4), __float_as_uint(f[4 * i]), __float_as_uint(f[4 * i + 1]),
                  __float_as_uint(f[4 * i + 2]), __float_as_uint(f[4 * i + 3]));
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_reduce_add_2d(&tmap_out, my_stage, col, m0 + q * 32);
          tma_store_commit();
        }
      }
      const bool last_unit = (u + npairs >= p.num_tiles);
      if (lane == 0) {
        if (do_cast && last_unit) {
          tma_store_wait<0>();  // complete, not just read: the rows are re-read below
          fence_proxy_async_global();
        } else {
          tma_store_wait_read<0>();
        }
      }
      if (ew == 0 && lane == 0) stamp(2, ui, 3);
      epi_bar_sync();  // every warp's staging (which aliases H) has been read before any warp writes H again
      cast_pending = do_cast && !last_unit;
      if (do_cast && last_unit) {
        // last unit of this CTA: nothing left to hide the cast under — all eight epilogue warps share its 128 rows
        // (every warp's reduce-adds completed before the barrier above)
        constexpr int kRowsPerWarp = MT / kEpiWarps;
        const int64_t r0 = int64_t(m0) + ew * kRowsPerWarp;
#pragma unroll 1
        for (int i = 0; i < kRowsPerWarp; i += 4) {
          const int64_t left = p.rows - (r0 + i);
          if (left <= 0) break;
          rowcast_rows<4, true>(p.x, p.xb_out, p.stats_out, r0 + i, 1, left < 4 ? int(left) : 4, lane);
        }
      }
    }
    if (lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, kTmemCols);
  }
}

}  // namespace

int vit_mlp_fused(const sais_bf16* xn, const sais_bf16* w1, const float* b1, const sais_bf16* w2, const float* b2,
                  float* x, int64_t rows, cudaStream_t stream, const float* ln_stats, const float* ln_colsum, float ln_eps,
                  sais_bf16* xb_out, float* stats_out) {
  if (rows == 0) return kOk;
  if (!xn || !w1 || !b1 || !w2 || !b2 || !x || rows < 0) {
    set_last_error("vit_mlp: bad arguments");
    return kErrInvalidArg;
  }
  if ((reinterpret_cast<uintptr_t>(xn) | reinterpret_cast<uintptr_t>(w1) | reinterpret_cast<uintptr_t>(w2) |
       reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(b1)) & 15) {
    set_last_error("vit_mlp: operands must be 16-byte aligned");
    return kErrInvalidArg;
  }
  if ((xb_out == nullptr) != (stats_out == nullptr) ||
      ((reinterpret_cast<uintptr_t>(xb_out) | reinterpret_cast<uintptr_t>(stats_out)) & 15)) {
    set_last_error("vit_mlp: xb_out and stats_out go together (16-byte aligned)");
    return kErrInvalidArg;
  }
  CUtensorMap ta, tw1, tw2, tout;
  int rc = make_tmap_2d(&ta, xn, kTmapBf16, uint64_t(rows), DM, DM, MT, 64, 128);
  if (rc) return rc;
  if ((rc = make_tmap_2d(&tw1, w1, kTmapBf16, HID, DM, DM, HC / 2, 64, 128))) return rc;
  if ((rc = make_tmap_2d(&tw2, w2, kTmapBf16, DM, HID, HID, 96, 64, 128))) return rc;
  if ((rc = make_tmap_2d(&tout, x, kTmapF32, uint64_t(rows), DM, DM, 32, 32, 128))) return rc;
  CUtensorMap txpf;  // L2-prefetch view of x: 128-row x 32-column boxes
  if ((rc = make_tmap_2d(&txpf, x, kTmapF32, uint64_t(rows), DM, DM, MT, 32, 128))) return rc;

  if ((rc = ensure_dynamic_smem(reinterpret_cast<const void*>(mlp_fused_kernel), kSmemBytes, "mlp_fused"))) return rc;
  MlpParams p;
  p.b1 = b1;
  p.b2 = b2;
  if ((ln_stats == nullptr) != (ln_colsum == nullptr) ||
      ((reinterpret_cast<uintptr_t>(ln_stats) | reinterpret_cast<uintptr_t>(ln_colsum)) & 15)) {
    set_last_error("vit_mlp: ln_stats and ln_colsum go together (16-byte aligned)");
    return kErrInvalidArg;
  }
  p.ln_stats = ln_stats;
  p.ln_colsum = ln_colsum;
  p.ln_eps = ln_eps;
  p.rows = rows;
  p.reverse = g_tile_reverse;
  p.num_tiles = int((rows + 2 * MT - 1) / (2 * MT));
  p.x = x;
  p.xb_out = reinterpret_cast<__nv_bfloat16*>(xb_out);
  p.stats_out = stats_out;
  const int pairs = balanced_ctas(p.num_tiles, num_sms() / 2);  // 197 row tiles at batch 256: 66 pairs, three rounds

  static const char* timeline = getenv("SAIS_MLP_TIMELINE");
  p.dbg = nullptr;
  constexpr int kDbgN = 4 * 32 * 4;
  if (timeline) {
    if (cudaMalloc(&p.dbg, kDbgN * sizeof(long long)) != cudaSuccess) p.dbg = nullptr;
    if (p.dbg) cudaMemsetAsync(p.dbg, 0, kDbgN * sizeof(long long), stream);
  }
  {
    LaunchScope ls(kClsMlpFused, stream, 4.0 * double(rows) * DM * HID);
    // (cluster = 1 here: the kernel carries its own __cluster_dims__(2, 1, 1))
    rc = check_cuda(launch_pdl(mlp_fused_kernel, dim3(2 * pairs), dim3(kThreads), size_t(kSmemBytes), stream, 1, ta, tw1, tw2,
                               tout, txpf, p),
                    "mlp_fused_kernel launch");
  }
  if (p.dbg) {
    static long long h[kDbgN];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, p.dbg, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(p.dbg);
    long long t0 = 0;
    for (long long v : h) if (v && (!t0 || v < t0)) t0 = v;
    if (FILE* f = fopen(timeline, "w")) {
      fprintf(f, "# rows=%lld tiles=%d pairs=%d cast=%d\n", (long long)rows, p.num_tiles, pairs, int(xb_out != nullptr));
      const char* names[4] = {"mma(chunk: s_empty ok, w_full(W1) ok, h_full ok, w_full(W2) ok)",
                              "e1(chunk: s_full, h_empty, h_full arrive)",
                              "unit(mma a_full, mma acc_empty, out acc_full, out done)",
                              "producer(chunk: w_empty for W1 ok, w_empty for W2 ok)"};
      for (int r = 0; r < 4; ++r) {
        fprintf(f, "%s\n", names[r]);
        for (int i = 0; i < 32; ++i) {
          fprintf(f, "  %2d:", i);
          for (int e = 0; e < 4; ++e) fprintf(f, " %8lld", h[(r * 32 + i) * 4 + e] ? h[(r * 32 + i) * 4 + e] - t0 : -1);
          fprintf(f, "\n");
        }
      }
      fclose(f);
    }
  }
  return rc;
}

}  // namespace sais

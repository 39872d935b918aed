// Persistent warp-specialised bf16 GEMM for sm_100a:  out = act(A · Wᵀ + bias) [+ residual].
//
//   warps 0..EW-1 : epilogue (EW = 8: two warps per TMEM lane quarter, each taking every other 32-column chunk;
//                   EW = 16: the lean GELU epilogue of fc1, four warps per quarter on 16-column register blocks):
//                   tcgen05.ld -> bias / folded LayerNorm / GELU / ReLU / residual -> swizzled smem staging -> TMA store.
//                   fp32 residual tiles are TMA-loaded into a ring of staging buffers nbuf - 1 chunks ahead, so
//                   the epilogue issues no scattered global accesses at all (row-per-thread accumulator
//                   layouts would otherwise turn every 16-byte store into its own 32-byte sector write).
//   warp EW       : TMA producer   (cp.async.bulk.tensor, 128-byte swizzle, multi-stage smem ring; in the A-stationary
//                   variant the A k-blocks of an m-tile group stay resident and the ring carries W only)
//   warp EW + 1   : TMEM allocator + tcgen05.mma issuer (128 x BLOCK_N x 16 per instruction, CTA pairs: 256 x BLOCK_N)
// The two role warps sit in the HIGHEST warp ids on purpose: the SM's warp arbiter favours higher
// warp ids, and the MMA issuer must never wait behind ALU-heavy epilogue warps of its sub-partition.
//
// Accumulators live in TMEM and are double buffered (2 x BLOCK_N fp32 columns), so the epilogue of
// tile i overlaps the MMAs of tile i+1.  Tiles are walked n-fastest so that CTAs running concurrently
// share the same A rows through L2 while the (small) weight matrix stays L2 resident.
//
// Split-precision mode (split3): A and W carry [hi | lo] bf16 halves and three passes hi·hi + lo·hi + hi·lo
// accumulate into the same TMEM tile — an fp32-equivalent product on the bf16 tensor pipe.
//
// LayerNorm folding (fast path): LayerNorm is affine per row, so  LN(x) · Wᵀ + b  =  rstd · (x · W'ᵀ − mean · c) + d  with
// W' = gamma ∘ W, c_n = Σ_k W'_nk, d_n = Σ_k beta_k W_nk + b_n.  A CONSUMER GEMM (qkv, fc1) therefore reads the raw
// bf16 residual stream and applies mean / rstd per row in its epilogue (ln_stats_in); the PRODUCER GEMM before it
// (proj, fc2) writes that bf16 copy next to the fp32 stream and per-row (sum, sum of squares) partials, one slot
// per (n-tile, epilogue-warp half), so no LayerNorm kernel and no atomics are needed (ln_stats_out / out2_bf16).
//
// Reference call sites this kernel replaces: nn.Linear / conv-as-GEMM in
// SAIS/scripts/dino-main/vision_transformer.py:60-63,82,90,126-130 and the in/out/FF projections of
// nn.TransformerEncoderLayer reached via SAIS/scripts/prepare_model.py:213.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "kernels.h"

namespace sais {

thread_local int g_tile_reverse = 0;

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
// Epilogue warps per CTA (template parameter EW): 8 (two per TMEM lane quarter) or 16 (four per quarter).  The epilogue is
// latency-bound with two warps per scheduler (ncu: 22% issue-active), so the bf16-output GEMMs of the big ViT shapes
// run with 16; the fp32-output modes stay at 8 (their staging tiles are twice as large).
constexpr int CW = 32;               // epilogue chunk width (columns)
constexpr bool kAStatDefault = true;   // A-stationary K = 384 GEMMs by default (SAIS_GEMM_ASTAT overrides)
constexpr int kStageBufBytes = 4096;  // one staging buffer: 32 rows x 128 B (fp32) or 2 x (32 rows x 64 B) (bf16 hi, lo)

template <int BLOCK_N, int CG, int EW>
struct GemmCfg {
  static constexpr int kEpiWarps = EW;
  static constexpr int kSub = EW / 4;                                    // epilogue warps per TMEM lane quarter
  static constexpr int kChunksPerWarp = (BLOCK_N / CW + kSub - 1) / kSub;  // at most this many 32-column chunks per warp
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = (BLOCK_N / CG) * BLOCK_K * 2;  // a CTA pair splits the W tile's rows
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = (2 * BLOCK_N <= 256) ? 256 : 512;
  static constexpr int kMaxStages = 6;
  // tail: barriers + per-warp bias slices (+ the folded-LayerNorm column-sum slices, consumer GEMMs only)
  static int tail_bytes(bool csum) { return 512 + (csum ? 2 : 1) * kEpiWarps * kChunksPerWarp * CW * 4; }
  static int budget(bool csum) { return 227 * 1024 - 1024 /*align slack*/ - tail_bytes(csum); }
  // staging per epilogue warp: two buffers of 32 rows x 128 B (fp32 / bf16 hi+lo) or 32 rows x 64 B (plain bf16)
  // + xb: one extra 2 KB tile per warp for the bf16 copy a LayerNorm-producer GEMM writes next to its fp32 output
  static int epi_bytes(bool wide, int nbuf, int xb = 0) {
    return kEpiWarps * (nbuf * (wide ? kStageBufBytes : kStageBufBytes / 2) + xb);
  }
  static int stages(bool wide, int nbuf, int xb = 0, bool csum = true) {
    const int s = (budget(csum) - epi_bytes(wide, nbuf, xb)) / kStageBytes;
    return s > kMaxStages ? kMaxStages : s;
  }
  static int smem_bytes(bool wide, int nbuf, int xb = 0, bool csum = true) {
    return stages(wide, nbuf, xb, csum) * kStageBytes + epi_bytes(wide, nbuf, xb) + 1024 + tail_bytes(csum);
  }
  // A-stationary variant (K = kAStatKB * 64): the A rows of an m-tile group stay resident for all of its n-tiles, the ring
  // carries W only
  static constexpr int kAStatKB = 6;
  static constexpr int kAResident = kAStatKB * kABytes;
  static int stages_astat(bool wide, int nbuf, bool csum) {
    const int s = (budget(csum) - epi_bytes(wide, nbuf, 0) - kAResident) / kBBytes;
    return s > kMaxStages ? kMaxStages : s;
  }
  static int smem_bytes_astat(bool wide, int nbuf, bool csum) {
    return kAResident + stages_astat(wide, nbuf, csum) * kBBytes + epi_bytes(wide, nbuf, 0) + 1024 + tail_bytes(csum);
  }
};

struct GemmParams {
  const float* bias;
  const float* residual;
  float* out_f32;
  __nv_bfloat16* out_bf16;
  const float* row_add;
  int64_t ldr, ldo32, ldo16;
  int M, N, K;
  int act;
  int remap_group;
  __nv_bfloat16* out2_bf16;  // LayerNorm-producer mode: bf16 copy of the fp32 output (pitch ldo2)
  int64_t ldo2;
  int split3;         // 1: A and W hold [hi | lo] bf16 halves (2K columns); accumulate hi*hi + lo*hi + hi*lo
  int split_out;      // 1: out_bf16 has 2N columns: hi = bf16(v) at n, lo = bf16(v - hi) at N + n
  int exact_gelu;     // 1: erff-based GELU (precise mode); 0: tanh.approx form fitted to the erf definition
  long long* dbg;     // dev knob (SAIS_GEMM_TIMELINE=<file>): CTA 0 records clock64() per role / tile / event
  int stages;         // depth of the operand ring
  int stage_buf;      // bytes per epilogue staging buffer (4096 or 2048)
  int reverse;        // walk the m-tiles from the last to the first (kernels.h g_tile_reverse)
  int remap_tma;      // patch-embed row remap through TMA: tmap_out / tmap_res are 3-D views [group][remap_group tokens][N] of
                      // the output with 32-row / 4-row boxes (opt-in, SAIS_PATCH_TMA=1)
  int nbuf;           // staging buffers per epilogue warp (2..4)
  // LayerNorm folding (see the file comment)
  const float* ln_stats_in;  // consumer: [M][4][2] (sum, sumsq) partials of the K-wide input rows
  const float* ln_colsum;    // consumer: c_n = sum_k W'_nk
  float ln_eps;
  float ln_inv_k;            // 1 / row length of the normalised input (= 1 / K)
  float* ln_stats_out;       // producer: [M][4][2], slot = n_tile * 2 + warp half (needs N / BLOCK_N == 2)
  int k_slices;              // >= 1; > 1 only in accumulate mode
  int accumulate;            // fp32 output is ADDED to what out_f32 holds (TMA reduce-add), no bias / residual
  int xb_buf;                // producer: bytes of the extra bf16 staging tile per warp (2048) or 0
};

// byte offset of 16-byte chunk `j` of row `r` inside a staging tile written/read by TMA
__device__ __forceinline__ uint32_t stage_off_f32(int r, int j) { return uint32_t(r * 128 + ((j ^ (r & 7)) << 4)); }        // SWIZZLE_128B
__device__ __forceinline__ uint32_t stage_off_bf16(int r, int j) { return uint32_t(r * 64 + ((j ^ ((r >> 1) & 3)) << 4)); }  // SWIZZLE_64B

// epilogue specialisations (keeps the hot loop small enough for the instruction cache)
enum : int {
  kModeBf16 = 0,      // bf16 out, act none / ReLU, no residual           (qkv, temporal FF1)
  kModeBf16Gelu = 1,  // bf16 out, fast erf-GELU                          (fc1)
  kModeF32 = 2,       // fp32 out, optional fp32 residual, no activation  (proj, fc2, split-precision projections)
  kModeGeneric = 3,   // everything else: [hi|lo] bf16 out, exact GELU, patch-embed row remap (direct stores)
};

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ ulonglong2 lds128_pairs(uint32_t addr) {  // two packed fp32x2 operands
  ulonglong2 v;
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void tma_load_2d_s(uint32_t smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0,
                                              int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d_s(const CUtensorMap* m, uint32_t smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d_s(const CUtensorMap* m, uint32_t smem_src, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d_s(const CUtensorMap* m, uint32_t smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}

// CG = 1: one CTA per 128 x BLOCK_N tile.  CG = 2: a CTA pair (2-CTA cluster, tcgen05 cta_group::2) per
// 256 x BLOCK_N tile — each CTA loads its own 128 rows of A but only HALF of the W tile, so the bytes every SM
// pulls from L2 per flop drop by 1/3 (the measured limiter of the CG = 1 kernel at K = 384).
// ASTAT (A-stationary, K = 384, CTA pairs): every CTA pair owns a CONTIGUOUS range of the (m-group, n-tile) sequence and
// keeps the six A k-blocks of the current m-group resident in shared memory for all n-tiles it covers there; the ring then
// streams W only.  The bytes every SM pulls through L2 — what paces the K = 384 mainloops — drop by
// ~45 % (qkv: 168 -> 94 KB per tile and CTA on average).
template <int BLOCK_N, int MODE, int CG, int EW, bool ASTAT = false>
__global__ void __launch_bounds__(32 * (2 + EW), 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_res,
                    const __grid_constant__ CUtensorMap tmap_out2, const GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N, CG, EW>;
  static_assert(!ASTAT || (CG == 2 && (MODE == kModeBf16 || MODE == kModeBf16Gelu)), "A-stationary: CTA pairs, bf16 outputs");
  constexpr int kEpiWarps = EW;
  constexpr int kSub = Cfg::kSub;
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled UMMA/TMA tiles need 1024-byte aligned bases
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int kStages = p.stages;
  constexpr int kRingStageBytes = ASTAT ? Cfg::kBBytes : Cfg::kStageBytes;
  uint8_t* a_res = smem;                                          // ASTAT: resident A k-blocks [kAStatKB][kABytes]
  uint8_t* ring = smem + (ASTAT ? Cfg::kAResident : 0);
  uint8_t* epi_smem = ring + kStages * kRingStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + kEpiWarps * (p.nbuf * p.stage_buf + p.xb_buf));
  uint64_t* full_bar = bars;                        // [kMaxStages]
  uint64_t* empty_bar = bars + Cfg::kMaxStages;     // [kMaxStages]
  uint64_t* tfull_bar = bars + 2 * Cfg::kMaxStages; // [2]
  uint64_t* tempty_bar = tfull_bar + 2;           // [2]
  constexpr int kResBars = (EW == 8) ? 4 : 2;     // residual-tile ring slots per epilogue warp (fp32 modes run on 8 warps)
  uint64_t* res_bar = tempty_bar + 2;             // [kEpiWarps][kResBars]
  uint64_t* afull_bar = res_bar + kResBars * kEpiWarps;  // ASTAT: [kAStatKB] A k-block resident / [kAStatKB] free again
  uint64_t* aempty_bar = afull_bar + Cfg::kAStatKB;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(aempty_bar + Cfg::kAStatKB);
  static_assert(16 + kResBars * kEpiWarps + 2 * Cfg::kAStatKB + 1 <= 64, "barrier area is 64 slots");
  float* bias_smem = reinterpret_cast<float*>(bars + 64);                  // [kEpiWarps][kChunksPerWarp * CW]
  float* csum_smem = bias_smem + kEpiWarps * (Cfg::kChunksPerWarp * CW);  // [kEpiWarps][kChunksPerWarp * CW]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // timeline rows: 0 producer, 1 MMA issuer, 2 epilogue warp 0, 3 epilogue warp 4; 16 tiles x 16 events each
  auto stamp = [&](int role, int idx, int ev) {
    if (p.dbg != nullptr && blockIdx.x == 0 && idx < 16 && ev < 16) p.dbg[(role * 16 + idx) * 16 + ev] = clock64();
  };

  // Work units: (m-tile group, n-tile); a group is `cluster` vertically adjacent m-tiles, one per CTA of the
  // cluster.  Units are walked n-fastest; a CTA whose m-tile lies beyond M just computes on zero-filled rows.
  constexpr int csize = CG;
  const uint32_t crank = CG > 1 ? cluster_ctarank() : 0;
  const int m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int n_tiles = p.N / BLOCK_N;
  const int num_tiles = ((m_tiles + csize - 1) / csize) * n_tiles;  // number of work units
  // ASTAT: CTA (pair) c walks the contiguous unit range [num_tiles * c / C, num_tiles * (c + 1) / C) one by one
  const int n_ctas = gridDim.x / csize, cta = blockIdx.x / csize;
  const int unit0 = ASTAT ? int(int64_t(num_tiles) * cta / n_ctas) : cta;
  const int unit_stride = ASTAT ? 1 : n_ctas;
  const int astat_end = int(int64_t(num_tiles) * (cta + 1) / n_ctas);
  const int kb_per_pass = p.K / BLOCK_K;
  const int k_blocks = p.split3 ? 3 * kb_per_pass : kb_per_pass;
  // split-K (accumulate mode, small M): a work unit is (tile, K slice); partial products meet in L2 (TMA reduce-add)
  const int S = p.k_slices;
  const int num_units = ASTAT ? astat_end : num_tiles * S;  // (loop bound of this CTA; ASTAT implies S == 1)
  const int kbs = (k_blocks + S - 1) / S;

  constexpr int kProducerWarp = kEpiWarps, kMmaWarp = kEpiWarps + 1;
  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], kEpiWarps * CG);  // (leader's copy) the epilogue warps of BOTH CTAs drain the tile
    }
    for (int s = 0; s < kResBars * kEpiWarps; ++s) mbar_init(&res_bar[s], 1);
    if (ASTAT)
      for (int s = 0; s < 2 * Cfg::kAStatKB; ++s) mbar_init(&afull_bar[s], 1);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    if (CG == 1) {
      tmem_alloc(tmem_base_smem, Cfg::kTmemCols);
      tmem_relinquish();
    } else {
      tmem_alloc_cg2(tmem_base_smem, Cfg::kTmemCols);
      tmem_relinquish_cg2();
    }
  }
  tc_fence_before();
  if (csize > 1) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  // PDL: the set-up above overlapped the previous kernel's tail; nothing before this line touched global memory
  pdl_wait();
  if (p.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {  // effective SM clock of this launch: cycles vs nanoseconds
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    p.dbg[4 * 16 * 16] = (long long)ns;
    p.dbg[4 * 16 * 16 + 1] = clock64();
  }
  const int m_groups = (m_tiles + csize - 1) / csize;
  auto grp_m0 = [&](int g) { return ((p.reverse ? m_groups - 1 - g : g) * csize + int(crank)) * BLOCK_M; };
  auto tile_m0 = [&](int unit) { return grp_m0(unit / n_tiles); };
  // unit -> (m-group, n-tile) positions are advanced incrementally where it matters (the epilogue's per-tile and per-chunk
  // bookkeeping): the integer divisions of the direct mapping were ~900 cycles at the top of every tile
  const int d_mt = unit_stride / n_tiles, d_nt = unit_stride % n_tiles;
  auto pos_of = [&](int u, int& mt, int& nt) {
    const int t = u / S;
    mt = t / n_tiles;
    nt = t - mt * n_tiles;
  };
  auto pos_next = [&](int& u, int& mt, int& nt) {  // the unit this CTA processes after u
    u += unit_stride;
    if (S == 1) {
      mt += d_mt;
      nt += d_nt;
      if (nt >= n_tiles) {
        nt -= n_tiles;
        ++mt;
      }
    } else {
      pos_of(u, mt, nt);
    }
  };

  if (warp == kProducerWarp) {
    // ===================== TMA producer (whole warp, warp-uniform; only the TMA instructions are elected) =====================
    {
      int stage = 0;
      uint32_t phase = 0;
      int tidx = 0;
      int seg = 0;  // ASTAT: m-group segments started by this CTA (= A reloads)
      for (int u = unit0; u < num_units; u += unit_stride, ++tidx) {
        const int tile = u / S;
        const int kb0 = (u % S) * kbs, kb1 = (kb0 + kbs < k_blocks) ? kb0 + kbs : k_blocks;
        const int m0 = tile_m0(tile);
        const int n0 = (tile % n_tiles) * BLOCK_N;
        if constexpr (ASTAT) {
          const bool new_seg = (u == unit0) || (tile % n_tiles == 0);
          for (int kb = 0; kb < Cfg::kAStatKB; ++kb) {
            if (new_seg) {
              // the MMAs of the previous segment that read A k-block kb have retired (committed with its last tile)
              if (seg > 0) mbar_wait(&aempty_bar[kb], (seg - 1) & 1);
              if (elect_one()) {
                if (crank == 0) mbar_arrive_expect_tx(&afull_bar[kb], Cfg::kABytes * CG);
                tma_load_2d_cg2(a_res + kb * Cfg::kABytes, &tmap_a, leader_smem_u32(&afull_bar[kb]), kb * BLOCK_K, m0);
              }
              __syncwarp();
            }
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (lane == 0) stamp(0, tidx, kb);
            if (elect_one()) {
              if (crank == 0) mbar_arrive_expect_tx(&full_bar[stage], Cfg::kBBytes * CG);
              tma_load_2d_cg2(ring + stage * kRingStageBytes, &tmap_b, leader_smem_u32(&full_bar[stage]), kb * BLOCK_K,
                              n0 + int(crank) * (BLOCK_N / 2));
            }
            __syncwarp();
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
          if (new_seg) ++seg;
          continue;
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (lane == 0) stamp(0, tidx, kb - kb0);
          uint8_t* sa = ring + stage * kRingStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          if (elect_one()) {
            if (CG == 1 || crank == 0)
              mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes * CG);
            // split3 passes: (A_hi, W_hi), (A_lo, W_hi), (A_hi, W_lo); halves sit side by side along K
            int ka = kb, kw = kb;
            if (kb >= 2 * kb_per_pass) {
              ka = kb - 2 * kb_per_pass;
              kw = kb - kb_per_pass;
            } else if (kb >= kb_per_pass) {
              kw = kb - kb_per_pass;
            }
            if (CG == 1) {
              tma_load_2d(sa, &tmap_a, &full_bar[stage], ka * BLOCK_K, m0);
              tma_load_2d(sb, &tmap_b, &full_bar[stage], kw * BLOCK_K, n0);
            } else {
              // both CTAs' loads complete on the LEADER's full barrier (its MMA thread is the only consumer)
              const uint32_t lbar = leader_smem_u32(&full_bar[stage]);
              // (no evict-first hint here: the ring variant's A tile is read by every n-tile's CTA pair, and the second
              // reader then misses — proj / fc2 5 us slower in situ)
              tma_load_2d_cg2(sa, &tmap_a, lbar, ka * BLOCK_K, m0);
              tma_load_2d_cg2(sb, &tmap_b, lbar, kw * BLOCK_K, n0 + int(crank) * (BLOCK_N / 2));
            }
          }
          __syncwarp();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    // The WHOLE warp walks the loop with warp-uniform control flow and values; only the tcgen05 instructions themselves
    // sit under elect_one().  Run by a single divergent lane (if (lane == 0) ...) the descriptors live in vector
    // registers and every tcgen05.mma costs an ELECT / 5 x R2UR.BROADCAST / branch "waterfall" on top of 64-bit vector
    // address arithmetic — measured ~215 cycles per MMA *independent of N*, i.e. the issuing thread, not the tensor
    // pipe, set the pace of every GEMM (profiles/r01f_mma_issue.md).
    if (crank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BLOCK_M * CG, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int astage = 0;
      uint32_t aphase = 0;
      int tidx = 0;
      int seg = -1;  // ASTAT: index of the m-group segment being multiplied
      for (int u = unit0; u < num_units; u += unit_stride, ++tidx) {
        const int kb0 = (u % S) * kbs, kb1 = (kb0 + kbs < k_blocks) ? kb0 + kbs : k_blocks;
        // ASTAT: first / last tile this CTA computes in the current m-group (A k-blocks arrive with the first, are released
        // with the last)
        const bool new_seg = ASTAT && ((u == unit0) || (u % n_tiles == 0));
        const bool last_of_seg = ASTAT && ((u + 1 == num_units) || ((u + 1) % n_tiles == 0));
        if (new_seg) ++seg;
        if (lane == 0) stamp(1, tidx, 0);
        mbar_wait(&tempty_bar[astage], aphase ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        if (lane == 0) stamp(1, tidx, 1);
        const uint32_t d_tmem = tmem_base + astage * BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          if (new_seg) mbar_wait(&afull_bar[kb], seg & 1);
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (lane == 0) stamp(1, tidx, 2 + kb - kb0);
          const uint32_t sa = ASTAT ? smem_u32(a_res + kb * Cfg::kABytes) : smem_u32(ring + stage * kRingStageBytes);
          const uint32_t sb = ASTAT ? smem_u32(ring + stage * kRingStageBytes) : sa + Cfg::kABytes;
          const uint64_t da = umma_desc_sw128_kmajor(sa);
          const uint64_t db = umma_desc_sw128_kmajor(sb);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              // advancing K inside the 128-byte swizzle row: +32 bytes = +2 in 16-byte address units
              if (CG == 1) umma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb != kb0) || (k != 0));
              else umma_f16_cg2(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb != kb0) || (k != 0));
            }
            // frees the smem slot (in both CTAs of a pair) once these MMAs retire
            if (CG == 1) umma_commit(&empty_bar[stage]); else umma_commit_cg2_mcast(&empty_bar[stage], uint16_t(0b11));
            if constexpr (ASTAT) {
              if (last_of_seg) umma_commit_cg2_mcast(&aempty_bar[kb], uint16_t(0b11));
            }
          }
          __syncwarp();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if (elect_one()) {
          if (CG == 1) umma_commit(&tfull_bar[astage]); else umma_commit_cg2_mcast(&tfull_bar[astage], uint16_t(0b11));
        }
        __syncwarp();
        if (++astage == 2) {
          astage = 0;
          aphase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue (warps 0..7) =====================
    const int ew = warp;       // 0..7
    const int q = warp & 3;    // TMEM lane quarter this warp may access
    const int half = ew >> 2;  // which of the kSub warps of the quarter: takes chunks half, half + kSub, ...
    constexpr int NC = BLOCK_N / CW;
    constexpr int NCWmax = Cfg::kChunksPerWarp;
    const int NCW = (NC - half + kSub - 1) / kSub;  // chunks of this warp per tile
    const int kBuf = p.stage_buf;
    const int nbuf = p.nbuf;
    const uint32_t my_stage = smem_u32(epi_smem + ew * (nbuf * kBuf + p.xb_buf));
    int bufi = 0;  // staging buffer ring index (it % nbuf)
    uint64_t* my_res_bar = res_bar + kResBars * ew;
    float* my_bias = bias_smem + ew * (NCWmax * CW);
    float* my_csum = csum_smem + ew * (NCWmax * CW);
    const bool ln_in = (MODE == kModeBf16 || MODE == kModeBf16Gelu) && p.ln_stats_in != nullptr;
    const bool ln_out = (MODE == kModeF32) && p.ln_stats_out != nullptr;
    const bool tma_epi = (MODE != kModeGeneric) || (p.remap_group == 0) || (p.remap_tma != 0);
    const bool has_res = (MODE == kModeF32 || MODE == kModeGeneric) && tma_epi && (p.residual != nullptr);
    const bool f32_out = (MODE == kModeF32) || (MODE == kModeGeneric && p.out_f32 != nullptr);

    int astage = 0;
    uint32_t aphase = 0;
    uint32_t it = 0;  // chunks processed by this warp (selects staging buffer / residual barrier phase)

    // Residual pipeline (S == 1 whenever there is a residual): the fp32 residual tile of a chunk is TMA-loaded into the
    // staging buffer the result will leave from, nbuf - 1 chunks AHEAD of its use.  One chunk ahead (nbuf = 2) leaves a
    // single 4 KB load in flight per warp behind a store-read wait — a ~3.5 k-cycle latency chain per chunk that made
    // proj epilogue-bound at 11.5 k cycles per tile against 2.5 k cycles of MMA (gpurun_out/s6f timelines).
    int pf_u = unit0, pf_c = half, pf_slot = 0;  // cursor: (unit, chunk) of the next residual tile to request, its ring slot
    int pf_mt, pf_nt;                            // (m-group, n-tile) of unit pf_u
    pos_of(unit0, pf_mt, pf_nt);
    const uint32_t xb_base = my_stage + nbuf * kBuf;  // bf16-copy staging tile(s) (LayerNorm-producer mode; one or two per warp)
    uint32_t xb_tile = xb_base;
    int rslot = 0;                               // ring slot / barrier phase of the chunk being consumed
    uint32_t rphase = 0;
    auto res_request = [&]() {  // one thread
      const int rm0 = grp_m0(pf_mt), rn0 = pf_nt * BLOCK_N;
      mbar_arrive_expect_tx(&my_res_bar[pf_slot], 32 * 128);
      tma_load_2d_s(my_stage + pf_slot * kBuf, &tmap_res, &my_res_bar[pf_slot], rn0 + pf_c * CW, rm0 + q * 32);
    };
    auto res_advance = [&]() {  // whole warp (the cursor stays warp-uniform)
      if (++pf_slot == nbuf) pf_slot = 0;
      pf_c += kSub;
      if (pf_c >= NC) {
        pf_c = half;
        pos_next(pf_u, pf_mt, pf_nt);
      }
    };
    if (has_res) {
      for (int i = 0; i < nbuf - 1; ++i)
        if (pf_u < num_units) {
          if (lane == 0) res_request();
          res_advance();
        }
    }

    // Per-tile vectors (bias slice, folded-LayerNorm column sums, row statistics) are fetched one tile AHEAD into
    // registers: the epilogue is the critical path of these GEMMs, so a global-load latency at the top of every tile
    // would be paid in full (16 tiles x ~800 cycles = 7 us at batch 256).
    float pf_bias[NCWmax], pf_csum[NCWmax];
    float4 pf_s0 = make_float4(0.f, 0.f, 0.f, 0.f), pf_s1 = pf_s0;
    auto prefetch_tile = [&](int t, int t_mt, int t_nt) {
      if (t >= num_units) return;
      const int pm0 = grp_m0(t_mt), pn0 = t_nt * BLOCK_N;
#pragma unroll
      for (int ci = 0; ci < NCWmax; ++ci) {
        pf_bias[ci] = (ci < NCW && p.bias) ? __ldg(p.bias + pn0 + (half + kSub * ci) * CW + lane) : 0.0f;
        if (ln_in) pf_csum[ci] = (ci < NCW) ? __ldg(p.ln_colsum + pn0 + (half + kSub * ci) * CW + lane) : 0.0f;
      }
      if (ln_in) {
        const int r = pm0 + q * 32 + lane;
        const float4* sp = reinterpret_cast<const float4*>(p.ln_stats_in + int64_t(r < p.M ? r : p.M - 1) * 8);
        pf_s0 = __ldg(sp);
        pf_s1 = __ldg(sp + 1);
      }
    };
    int cur_mt, cur_nt;  // position of the unit being processed
    pos_of(unit0, cur_mt, cur_nt);
    prefetch_tile(unit0, cur_mt, cur_nt);

    int tidx = 0;
    const int erole = (ew == 0) ? 2 : (ew == 4 ? 3 : -1);
    if constexpr (EW == 16) {
      // ---- lean bf16 epilogue: sixteen warps (four per TMEM lane quarter = four per scheduler) working on
      // 16-column register blocks.  The 8-warp epilogue below is a ~900-cycle dependent instruction stream per 32-column
      // chunk (sum of its SASS stall counts) on two warps per scheduler: its FMA and MUFU pipes sit idle 2/3 of the
      // time and fc1 + GELU runs 7.5 k cycles of epilogue against 2.9 k cycles of MMA per tile.  More warps only fit the
      // register file (65,536 / 576 threads = 112) with half-width blocks and without the register-resident bias /
      // column-sum double buffers (four warps per scheduler hide the broadcast LDS instead).
      static_assert(MODE == kModeBf16 || MODE == kModeBf16Gelu, "16 epilogue warps: bf16-output modes only");
      int mt = cur_mt, nt = cur_nt;  // (S == 1 in the bf16 modes)
      const uint32_t bias_s = smem_u32(my_bias), csum_s = ln_in ? smem_u32(my_csum) : bias_s;
      // GELU mode: the block below produces h = x / 2 straight away (bias, rstd and -mean * rstd halved: exact scalings),
      // which is what gelu_erf_fast2_half wants
      constexpr float kHalf = (MODE == kModeBf16Gelu) ? 0.5f : 1.0f;
      for (int u = unit0; u < num_units; u += unit_stride, ++tidx) {
        const int m0 = grp_m0(mt);
        const int n0 = nt * BLOCK_N;
        int nmt = mt + d_mt, nnt = nt + d_nt;
        if (nnt >= n_tiles) {
          nnt -= n_tiles;
          ++nmt;
        }
        if (erole >= 0 && lane == 0) stamp(erole, tidx, 0);
        // accumulator ready?  (the epilogue is the slower side, so normally yes) -> start the first TMEM read right away:
        // it streams in under the per-tile vector staging below instead of after it
        mbar_wait(&tfull_bar[astage], aphase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + astage * BLOCK_N;
        uint32_t va[16], vb[16];
        tmem_ld_32x16(t_row + half * CW, va);
        if (erole >= 0 && lane == 0) stamp(erole, tidx, 1);
        __syncwarp();
#pragma unroll
        for (int ci = 0; ci < NCWmax; ++ci)
          if (ci < NCW) {
            my_bias[ci * CW + lane] = pf_bias[ci] * kHalf;
            if (ln_in) my_csum[ci * CW + lane] = pf_csum[ci];
          }
        float ln_rstd = kHalf, ln_nmr = 0.0f;
        if (ln_in) {
          const float mean = ((pf_s0.x + pf_s0.z) + (pf_s1.x + pf_s1.z)) * p.ln_inv_k;
          const float var = fmaxf(((pf_s0.y + pf_s0.w) + (pf_s1.y + pf_s1.w)) * p.ln_inv_k - mean * mean, 0.0f);
          const float rstd = rsqrtf(var + p.ln_eps);
          ln_rstd = rstd * kHalf;
          ln_nmr = -mean * rstd * kHalf;
        }
        if (u + unit_stride < num_units) {  // next tile's vectors (consumed at the top of the next iteration)
          const int pm0 = grp_m0(nmt), pn0 = nnt * BLOCK_N;
#pragma unroll
          for (int ci = 0; ci < NCWmax; ++ci) {
            pf_bias[ci] = (ci < NCW && p.bias) ? __ldg(p.bias + pn0 + (half + kSub * ci) * CW + lane) : 0.0f;
            if (ln_in) pf_csum[ci] = (ci < NCW) ? __ldg(p.ln_colsum + pn0 + (half + kSub * ci) * CW + lane) : 0.0f;
          }
          if (ln_in) {
            const int r = pm0 + q * 32 + lane;
            const float4* sp = reinterpret_cast<const float4*>(p.ln_stats_in + int64_t(r < p.M ? r : p.M - 1) * 8);
            pf_s0 = __ldg(sp);
            pf_s1 = __ldg(sp + 1);
          }
        }
        const uint64_t rs2 = pack2(ln_rstd, ln_rstd), nm2 = pack2(ln_nmr, ln_nmr);
        __syncwarp();

        // one 16-column block: bias / folded LayerNorm, activation, bf16, two 16-byte pieces of the staging row
        auto block16 = [&](uint32_t(&v)[16], int ci, int hh, uint32_t buf) {
          const uint32_t voff = uint32_t((ci * CW + hh * 16) * 4);
          float f[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            // one instruction stream for both cases (a predicated if / else would issue both): without the folded LayerNorm
            // nm2 = 0 (c2 re-reads the bias slice) and rs2 = 1 (0.5 in GELU mode), so the two FFMA2 reduce to acc + bias exactly
            const ulonglong2 b2 = lds128_pairs(bias_s + voff + j * 16);
            const ulonglong2 c2 = lds128_pairs(csum_s + voff + j * 16);
            unpack2(fma2(pack2u(v[4 * j], v[4 * j + 1]), rs2, fma2(nm2, c2.x, b2.x)), f[4 * j], f[4 * j + 1]);
            unpack2(fma2(pack2u(v[4 * j + 2], v[4 * j + 3]), rs2, fma2(nm2, c2.y, b2.y)), f[4 * j + 2], f[4 * j + 3]);
          }
          if (MODE == kModeBf16Gelu) {
#pragma unroll
            for (int j = 0; j < 16; j += 2) gelu_erf_fast2_half(f[j], f[j + 1]);
          } else if (p.act == 2) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.0f);
          }
#pragma unroll
          for (int j = 0; j < 2; ++j)
            sts128(buf + stage_off_bf16(lane, hh * 2 + j), pack_bf16x2(f[8 * j], f[8 * j + 1]),
                   pack_bf16x2(f[8 * j + 2], f[8 * j + 3]), pack_bf16x2(f[8 * j + 4], f[8 * j + 5]),
                   pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
        };

#pragma unroll 1
        for (int ci = 0; ci < NCW; ++ci) {
          const int c = half + kSub * ci;
          const uint32_t buf = my_stage + bufi * kBuf;
          tmem_ld_wait_dep(va);
          tmem_ld_32x16(t_row + c * CW + 16, vb);  // second half streams in under the first half's math
          // the store issued from this staging buffer nbuf chunks ago must have finished reading it
          __syncwarp();
          if (elect_one()) {
            if (nbuf == 1) tma_store_wait_read<0>();
            else if (nbuf == 2) tma_store_wait_read<1>();
            else tma_store_wait_read<2>();
          }
          __syncwarp();
          block16(va, ci, 0, buf);
          tmem_ld_wait_dep(vb);
          if (ci + 1 < NCW) {
            tmem_ld_32x16(t_row + (c + kSub) * CW, va);  // next chunk's first half
          } else {  // last TMEM read of this tile by this warp: hand the accumulator back
            tc_fence_before();
            if (lane == 0) {
              if (CG == 1) mbar_arrive(&tempty_bar[astage]);
              else mbar_arrive_cluster(leader_smem_u32(&tempty_bar[astage]));
            }
          }
          block16(vb, ci, 1, buf);
          fence_proxy_async_smem();
          __syncwarp();
          if (elect_one()) {  // (the same lane every time: it owns this warp's bulk-async groups)
            tma_store_2d_s(&tmap_out, buf, n0 + c * CW, m0 + q * 32);
            tma_store_commit();
          }
          if (++bufi == nbuf) bufi = 0;
          if (erole >= 0 && lane == 0) stamp(erole, tidx, 2 + ci);
        }
        if (++astage == 2) {
          astage = 0;
          aphase ^= 1;
        }
        mt = nmt;
        nt = nnt;
      }
    } else
    for (int u = unit0; u < num_units; u += unit_stride, ++tidx) {
      const int m0 = grp_m0(cur_mt);
      const int n0 = cur_nt * BLOCK_N;
      const int tile_nt = cur_nt;
      int nxt_u = u, nxt_mt = cur_mt, nxt_nt = cur_nt;
      pos_next(nxt_u, nxt_mt, nxt_nt);
      cur_mt = nxt_mt;  // (m0 / n0 / tile_nt hold this tile's position from here on)
      cur_nt = nxt_nt;
      if (erole >= 0 && lane == 0) stamp(erole, tidx, 0);
      // accumulator ready?  (the epilogue is normally the slower side, so yes) -> start the first TMEM read right away: it
      // streams in under the per-tile vector staging below instead of after it
      mbar_wait(&tfull_bar[astage], aphase);
      tc_fence_after();
      if (erole >= 0 && lane == 0) stamp(erole, tidx, 1);
      const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + astage * BLOCK_N;
      const int row = m0 + q * 32 + lane;

      const int remap_pidx = (MODE == kModeGeneric && p.remap_tma) ? row % p.remap_group : 0;
      uint32_t v[32];
      tmem_ld_32x32(t_row + half * CW, v);
      // this tile's vectors: registers -> per-warp smem (later reads are broadcast loads), then fetch the next tile's
      __syncwarp();
#pragma unroll
      for (int ci = 0; ci < NCWmax; ++ci)
        if (ci < NCW) my_bias[ci * CW + lane] = pf_bias[ci];
      float ln_rstd = 1.0f, ln_nmr = 0.0f;  // consumer: out = acc * rstd + (-mean * rstd) * c_n + d_n
      if (ln_in) {
#pragma unroll
        for (int ci = 0; ci < NCWmax; ++ci)
          if (ci < NCW) my_csum[ci * CW + lane] = pf_csum[ci];
        const float mean = ((pf_s0.x + pf_s0.z) + (pf_s1.x + pf_s1.z)) * p.ln_inv_k;
        const float var = fmaxf(((pf_s0.y + pf_s0.w) + (pf_s1.y + pf_s1.w)) * p.ln_inv_k - mean * mean, 0.0f);
        ln_rstd = rsqrtf(var + p.ln_eps);
        ln_nmr = -mean * ln_rstd;
      }
      prefetch_tile(nxt_u, nxt_mt, nxt_nt);
      uint64_t ln_sum2 = 0, ln_sq2 = 0;  // producer: this warp's share of the row statistics of the tile (packed pairs)
      __syncwarp();

      // chunk 0's bias pairs -> registers before the accumulator is ready; every chunk reloads them for the next one
      // right after its first op, so the shared-memory latency never sits in front of the epilogue math
      const uint32_t bias_s = smem_u32(my_bias), csum_s = smem_u32(my_csum);
      ulonglong2 bv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) bv[j] = lds128_pairs(bias_s + j * 16);

#pragma unroll 1
      for (int ci = 0; ci < NCW; ++ci) {
        const int c = half + kSub * ci;
        const int n = n0 + c * CW;
        float f[32];
        tmem_ld_wait_dep(v);
        {
          // bias add (or the folded LayerNorm's  acc * rstd + (-mean * rstd) * c_n + d_n) in packed fp32x2; this is
          // also what moves the accumulator out of v, so the next chunk's TMEM load can be issued right after it
          if (ln_in) {
            const uint64_t rs2 = pack2(ln_rstd, ln_rstd), nm2 = pack2(ln_nmr, ln_nmr);
            ulonglong2 cv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) cv[j] = lds128_pairs(csum_s + (ci * CW + j * 4) * 4);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const ulonglong2 b2 = bv[j], c2 = cv[j];
              unpack2(fma2(pack2u(v[4 * j], v[4 * j + 1]), rs2, fma2(nm2, c2.x, b2.x)), f[4 * j], f[4 * j + 1]);
              unpack2(fma2(pack2u(v[4 * j + 2], v[4 * j + 3]), rs2, fma2(nm2, c2.y, b2.y)), f[4 * j + 2], f[4 * j + 3]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const ulonglong2 b2 = bv[j];
              unpack2(add2(pack2u(v[4 * j], v[4 * j + 1]), b2.x), f[4 * j], f[4 * j + 1]);
              unpack2(add2(pack2u(v[4 * j + 2], v[4 * j + 3]), b2.y), f[4 * j + 2], f[4 * j + 3]);
            }
          }
          if (ci + 1 < NCW) {
#pragma unroll
            for (int j = 0; j < 8; ++j) bv[j] = lds128_pairs(bias_s + ((ci + 1) * CW + j * 4) * 4);
          }
        }
        if (ci + 1 < NCW) {
          tmem_ld_32x32(t_row + (c + kSub) * CW, v);  // next chunk's accumulator streams in under this chunk's math
        } else {  // last TMEM read of this tile by this warp: hand the accumulator back early
          tc_fence_before();
          if (lane == 0) {
            if (CG == 1) mbar_arrive(&tempty_bar[astage]);
            else mbar_arrive_cluster(leader_smem_u32(&tempty_bar[astage]));
          }
        }
        if (MODE == kModeBf16Gelu) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) gelu_erf_fast2(f[j], f[j + 1]);
        } else if (MODE == kModeBf16) {
          if (p.act == 2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
          }
        } else if (MODE == kModeGeneric) {
          if (p.act == 1) {
            if (p.exact_gelu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = gelu_erf_exact(f[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; j += 2) gelu_erf_fast2(f[j], f[j + 1]);
            }
          } else if (p.act == 2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
          }
        }

        if (MODE == kModeGeneric && p.remap_tma) {
          // patch-embed: + row_add[row % remap_group] (the position embedding of the token this GEMM row becomes)
          const float* radd = p.row_add + int64_t(remap_pidx) * p.N + n;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 a4 = __ldg(reinterpret_cast<const float4*>(radd + j));
            f[j] += a4.x; f[j + 1] += a4.y; f[j + 2] += a4.z; f[j + 3] += a4.w;
          }
        }
        if (tma_epi) {
          const uint32_t buf = my_stage + (has_res ? rslot : bufi) * kBuf;
          if (has_res) {
            mbar_wait(&my_res_bar[rslot], rphase);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 r4 = lds128(buf + stage_off_f32(lane, j));
              unpack2(add2(pack2(f[4 * j], f[4 * j + 1]), pack2(r4.x, r4.y)), f[4 * j], f[4 * j + 1]);
              unpack2(add2(pack2(f[4 * j + 2], f[4 * j + 3]), pack2(r4.z, r4.w)), f[4 * j + 2], f[4 * j + 3]);
            }
            if (ln_out) {
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                const uint64_t x2 = pack2(f[j], f[j + 1]);
                ln_sum2 = add2(ln_sum2, x2);
                ln_sq2 = fma2(x2, x2, ln_sq2);
              }
              if (p.xb_buf) {
                // bf16 copy of the tile -> its own 64B-swizzled staging tile; leaves with the fp32 tile's TMA store below
                // (full 64-byte row segments instead of 32 scattered 16-byte pieces per store instruction, which cost
                // the LSU one wavefront each: 10 us per GEMM at batch 256).  The previous chunk's stores must have
                // finished reading the tile: they were issued a whole chunk ago, so this wait is all but free.
                __syncwarp();
                if (elect_one()) {
                  if (p.xb_buf > 2048) tma_store_wait_read<1>();  // two tiles: only the store before the previous one
                  else tma_store_wait_read<0>();
                }
                __syncwarp();
                if (p.xb_buf > 2048) xb_tile = (xb_tile == xb_base) ? xb_base + 2048u : xb_base;
                const uint32_t xb = xb_tile;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  sts128(xb + stage_off_bf16(lane, j), pack_bf16x2(f[8 * j], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                         pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
              } else if (row < p.M) {
                // bf16 copy of the row segment straight from registers (64 contiguous bytes per thread)
                uint4* xp = reinterpret_cast<uint4*>(p.out2_bf16 + int64_t(row) * p.ldo2 + n);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  xp[j] = make_uint4(pack_bf16x2(f[8 * j], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                                     pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
              }
            }
          } else {
            // the store issued from this buffer nbuf chunks ago must have finished reading it
            __syncwarp();
            if (elect_one()) {
              if (nbuf == 2) tma_store_wait_read<1>();
              else if (nbuf == 3) tma_store_wait_read<2>();
              else tma_store_wait_read<3>();
            }
            __syncwarp();
          }
          if (f32_out) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              sts128(buf + stage_off_f32(lane, j), __float_as_uint(f[4 * j]), __float_as_uint(f[4 * j + 1]),
                     __float_as_uint(f[4 * j + 2]), __float_as_uint(f[4 * j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t o0 = pack_bf16x2(f[8 * j], f[8 * j + 1]), o1 = pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
              const uint32_t o2 = pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), o3 = pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
              sts128(buf + stage_off_bf16(lane, j), o0, o1, o2, o3);
              if (MODE == kModeGeneric && p.split_out) {
                sts128(buf + 2048 + stage_off_bf16(lane, j),
                       pack_bf16x2(f[8 * j] - bf16_lo(o0), f[8 * j + 1] - bf16_hi(o0)),
                       pack_bf16x2(f[8 * j + 2] - bf16_lo(o1), f[8 * j + 3] - bf16_hi(o1)),
                       pack_bf16x2(f[8 * j + 4] - bf16_lo(o2), f[8 * j + 5] - bf16_hi(o2)),
                       pack_bf16x2(f[8 * j + 6] - bf16_lo(o3), f[8 * j + 7] - bf16_hi(o3)));
              }
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (elect_one()) {  // (the same lane every time: it owns this warp's bulk-async groups)
            if (has_res && pf_u < num_units) {
              // every earlier store has finished reading smem -> the previous chunk's buffer (ring slot pf_slot) is free:
              // request the residual tile nbuf - 1 chunks ahead into it (ahead of this chunk's store in the TMA queue)
              tma_store_wait_read<0>();
              res_request();
            }
            if (MODE == kModeGeneric && p.remap_tma) {
              // GEMM rows g * G + i -> token i of group g in the 3-D view (its token axis starts at output row 1 of every
              // group).  The 32 rows of a chunk may straddle two groups: the box of the first store is clipped at token G
              // (upper-bound clipping is legal for stores; a NEGATIVE start coordinate is an illegal instruction,
              // tools/tma3d_test.cu), the rows of the next group leave as 4-row boxes from their offset in the staging tile
              // (G and the chunk start are multiples of 4).
              const int G = p.remap_group, groups = p.M / G;
              const int r0 = m0 + q * 32, g0 = r0 / G, p0 = r0 - g0 * G;
              if (g0 < groups) tma_store_3d_s(&tmap_out, buf, n, p0, g0);
              if (p0 + 32 > G && g0 + 1 < groups)
                for (int i = G - p0; i < 32; i += 4) tma_store_3d_s(&tmap_res, buf + i * 128, n, i - (G - p0), g0 + 1);
            } else {
              if (MODE == kModeF32 && p.accumulate) tma_reduce_add_2d_s(&tmap_out, buf, n, m0 + q * 32);
              else tma_store_2d_s(&tmap_out, buf, n, m0 + q * 32);
              if (MODE == kModeGeneric && p.split_out) tma_store_2d_s(&tmap_out, buf + 2048, p.N + n, m0 + q * 32);
              if (MODE == kModeF32 && ln_out && p.xb_buf) tma_store_2d_s(&tmap_out2, xb_tile, n, m0 + q * 32);
            }
            tma_store_commit();
          }
          if (has_res) {
            if (pf_u < num_units) res_advance();
            if (++rslot == nbuf) {
              rslot = 0;
              rphase ^= 1;
            }
          }
          ++it;
          if (++bufi == nbuf) bufi = 0;
          if (erole >= 0 && lane == 0) stamp(erole, tidx, 2 + ci);
        } else if (MODE == kModeGeneric) {
          if (row < p.M) {
            // ---- direct path (patch-embed row remap: GEMM row g*G + i -> token row g*(G+1) + 1 + i, + row_add[i]) ----
            const int g = row / p.remap_group;
            const int pidx = row - g * p.remap_group;
            const int64_t orow = int64_t(g) * (p.remap_group + 1) + 1 + pidx;
            const float* radd = p.row_add + int64_t(pidx) * p.N + n;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 a4 = __ldg(reinterpret_cast<const float4*>(radd + j));
              f[j] += a4.x; f[j + 1] += a4.y; f[j + 2] += a4.z; f[j + 3] += a4.w;
            }
            if (p.residual != nullptr) {
              const float* rp = p.residual + orow * p.ldr + n;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 r4 = *reinterpret_cast<const float4*>(rp + j);
                f[j] += r4.x; f[j + 1] += r4.y; f[j + 2] += r4.z; f[j + 3] += r4.w;
              }
            }
            if (f32_out) {
              float* op = p.out_f32 + orow * p.ldo32 + n;
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(op + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
            } else {
              __nv_bfloat16* op = p.out_bf16 + orow * p.ldo16 + n;
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 o;
                o.x = pack_bf16x2(f[j], f[j + 1]);
                o.y = pack_bf16x2(f[j + 2], f[j + 3]);
                o.z = pack_bf16x2(f[j + 4], f[j + 5]);
                o.w = pack_bf16x2(f[j + 6], f[j + 7]);
                *reinterpret_cast<uint4*>(op + j) = o;
              }
            }
          }
        }
      }
      if (ln_out && row < p.M) {
        const int slot = tile_nt * 2 + half;
        float s0, s1, q0, q1;
        unpack2(ln_sum2, s0, s1);
        unpack2(ln_sq2, q0, q1);
        *reinterpret_cast<float2*>(p.ln_stats_out + int64_t(row) * 8 + slot * 2) = make_float2(s0 + s1, q0 + q1);
      }
      if (++astage == 2) {
        astage = 0;
        aphase ^= 1;
      }
    }
    __syncwarp();
    if (elect_one()) tma_store_wait<0>();  // smem must stay valid until the bulk stores have drained
  }

  if (p.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    p.dbg[4 * 16 * 16 + 2] = (long long)ns;
    p.dbg[4 * 16 * 16 + 3] = clock64();
  }
  tc_fence_before();
  // (cluster) no CTA may exit while its peer can still multicast into its smem or arrive on its barriers
  if (csize > 1) cluster_sync_all(); else __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    if (CG == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols); else tmem_dealloc_cg2(tmem_base, Cfg::kTmemCols);
  }
}

template <int BLOCK_N, int MODE, int CG, int EW, bool ASTAT = false>
int launch_gemm(const SaisGemmArgs& a, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N, CG, EW>;
  CUtensorMap ta, tb, tout, tres, tout2;
  const uint64_t kcols = uint64_t(a.K) * (a.split3 ? 2 : 1);
  int rc = make_tmap_2d(&ta, a.a, kTmapBf16, uint64_t(a.M), kcols, uint64_t(a.lda), BLOCK_M, BLOCK_K, 128);
  if (rc) return rc;
  const int m_tiles = int((a.M + BLOCK_M - 1) / BLOCK_M);
  constexpr int cluster = CG;
  rc = make_tmap_2d(&tb, a.w, kTmapBf16, uint64_t(a.N), kcols, uint64_t(a.ldw), BLOCK_N / cluster, BLOCK_K, 128);
  if (rc) return rc;
  tout = ta;
  tres = ta;
  tout2 = ta;
  // patch-embed row remap through TMA stores instead of per-thread 16-byte stores (80 us in situ at batch 256 against 41 us
  // for the shape through the TMA epilogue; +1.2 % frames/s on the whole step, full parity suite green).
  // SAIS_PATCH_TMA=0 restores the direct stores (A/B).
  static const int env_patch_tma = getenv("SAIS_PATCH_TMA") ? atoi(getenv("SAIS_PATCH_TMA")) : 1;
  const bool remap_tma = MODE == kModeGeneric && a.remap_group > 0 && a.out_f32 && !a.residual && !a.split_out && env_patch_tma &&
                         a.M % a.remap_group == 0 && a.remap_group % 4 == 0;
  if (remap_tma) {
    const uint64_t groups = uint64_t(a.M / a.remap_group), gpitch = uint64_t(a.remap_group + 1) * uint64_t(a.ldo32);
    rc = make_tmap_f32_3d(&tout, a.out_f32 + a.ldo32, uint64_t(a.N), uint64_t(a.remap_group), groups, uint64_t(a.ldo32), gpitch, CW, 32);
    if (rc) return rc;
    rc = make_tmap_f32_3d(&tres, a.out_f32 + a.ldo32, uint64_t(a.N), uint64_t(a.remap_group), groups, uint64_t(a.ldo32), gpitch, CW, 4);
    if (rc) return rc;
  }
  if (a.remap_group == 0) {
    if (a.out_f32)
      rc = make_tmap_2d(&tout, a.out_f32, kTmapF32, uint64_t(a.M), uint64_t(a.N), uint64_t(a.ldo32), 32, CW, 128);
    else
      rc = make_tmap_2d(&tout, a.out_bf16, kTmapBf16, uint64_t(a.M), uint64_t(a.N) * (a.split_out ? 2 : 1),
                        uint64_t(a.ldo16), 32, CW, 64);
    if (rc) return rc;
    if (a.residual) {
      rc = make_tmap_2d(&tres, a.residual, kTmapF32, uint64_t(a.M), uint64_t(a.N), uint64_t(a.ldr), 32, CW, 128);
      if (rc) return rc;
    }
    if (a.out2_bf16) {
      rc = make_tmap_2d(&tout2, a.out2_bf16, kTmapBf16, uint64_t(a.M), uint64_t(a.N), uint64_t(a.ldo2), 32, CW, 64);
      if (rc) return rc;
    }
  }

  if ((rc = ensure_dynamic_smem(reinterpret_cast<const void*>(gemm_tcgen05_kernel<BLOCK_N, MODE, CG, EW, ASTAT>), 227 * 1024,
                                "gemm_tcgen05")))
    return rc;
  const bool wide = a.out_f32 != nullptr || a.split_out;  // 128-byte staging rows
  GemmParams p;
  p.bias = a.bias;
  p.residual = a.residual;
  p.out_f32 = a.out_f32;
  p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(a.out_bf16);
  p.row_add = a.row_add;
  p.ldr = a.ldr;
  p.ldo32 = a.ldo32;
  p.ldo16 = a.ldo16;
  p.M = int(a.M);
  p.N = int(a.N);
  p.K = int(a.K);
  p.act = a.act;
  p.remap_group = a.remap_group;
  p.split3 = a.split3;
  p.split_out = a.split_out;
  p.exact_gelu = a.split3;  // the fp32-equivalent mode keeps the erff form
  static const char* timeline = getenv("SAIS_GEMM_TIMELINE");
  p.dbg = nullptr;
  constexpr int kDbgN = 4 * 16 * 16 + 4;
  if (timeline) {
    if (cudaMalloc(&p.dbg, kDbgN * sizeof(long long)) != cudaSuccess) p.dbg = nullptr;
    if (p.dbg) cudaMemsetAsync(p.dbg, 0, kDbgN * sizeof(long long), stream);
  }
  // staging buffers per epilogue warp; the residual epilogue runs a ring of them: residual tiles are requested nbuf - 1
  // chunks ahead
  int nbuf = (EW == 16) ? 1 : 2;  // 16 warps x one 2 KB tile = the 8-warp epilogue's staging footprint (keeps the operand ring depth)
  p.ln_stats_in = a.ln_stats_in;
  p.ln_colsum = a.ln_colsum;
  p.ln_eps = a.ln_eps;
  p.ln_inv_k = 1.0f / float(a.K);
  p.ln_stats_out = a.ln_stats_out;
  p.out2_bf16 = reinterpret_cast<__nv_bfloat16*>(a.out2_bf16);
  p.ldo2 = a.ldo2;
  p.xb_buf = a.out2_bf16 ? 2048 : 0;
  // Short-K residual GEMMs (proj: six k-blocks per tile) are bound by their epilogue's residual / store chain, not by the
  // operand ring: trade two ring stages for a third staging buffer (residual tiles requested two chunks ahead) and a second
  // bf16-copy tile.  Long-K ones (fc2) need the deep ring more (3 stages: 69 -> 77 us).
  if (a.residual && a.remap_group == 0 && !a.split3 && a.K <= 512 && a.out_f32 && !a.k_slices) {
    nbuf = 3;
    if (p.xb_buf) p.xb_buf = 4096;
  }
  p.nbuf = nbuf;
  const bool csum = a.ln_stats_in != nullptr;  // only consumer GEMMs keep column-sum slices in the tail
  p.accumulate = a.k_slices > 0;
  p.k_slices = 1;
  if (a.k_slices > 1) {  // every slice must own at least one k-block
    const int kb_total = int(a.K / BLOCK_K) * (a.split3 ? 3 : 1);
    int want = a.k_slices < kb_total ? a.k_slices : kb_total;
    const int per = (kb_total + want - 1) / want;
    p.k_slices = (kb_total + per - 1) / per;
  }
  p.stages = ASTAT ? Cfg::stages_astat(wide, nbuf, csum) : Cfg::stages(wide, nbuf, p.xb_buf, csum);
  p.stage_buf = wide ? kStageBufBytes : kStageBufBytes / 2;
  p.reverse = g_tile_reverse;
  p.remap_tma = remap_tma ? 1 : 0;
  const int units = ((m_tiles + cluster - 1) / cluster) * (p.N / BLOCK_N) * p.k_slices;
  int grid = balanced_ctas(units, num_sms() / cluster) * cluster;
  const int cls = a.split3 ? kClsGemmSplit
                  : (a.M >= 4096 && a.N == 1152 && a.K == 384 && a.out_bf16) ? kClsGemmQkv
                  : (a.M >= 4096 && a.N == 384 && a.K == 384 && a.out_f32) ? kClsGemmProj
                                                                          : kClsGemm;
  LaunchScope ls(cls, stream, 2.0 * double(a.M) * double(a.N) * double(a.K) * (a.split3 ? 3 : 1));
  const size_t smem_bytes = ASTAT ? size_t(Cfg::smem_bytes_astat(wide, nbuf, csum)) : size_t(Cfg::smem_bytes(wide, nbuf, p.xb_buf, csum));
  rc = check_cuda(launch_pdl(gemm_tcgen05_kernel<BLOCK_N, MODE, CG, EW, ASTAT>, dim3(grid), dim3(32 * (2 + EW)), smem_bytes, stream,
                             cluster, ta, tb, tout, tres, tout2, p),
                  "gemm_tcgen05_kernel launch");
  if (p.dbg) {  // dev knob: dump CTA 0's timeline (cycles relative to the first stamp), last call wins
    static long long h[kDbgN];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, p.dbg, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(p.dbg);
    long long t0 = 0;
    for (int i = 0; i < 4 * 16 * 16; ++i) if (h[i] && (!t0 || h[i] < t0)) t0 = h[i];
    if (FILE* f = fopen(timeline, "w")) {
      fprintf(f, "# M=%d N=%d K=%d BN=%d mode=%d CG=%d grid=%d stages=%d\n", p.M, p.N, p.K, BLOCK_N, MODE, CG, grid, p.stages);
      const long long dns = h[4 * 16 * 16 + 2] - h[4 * 16 * 16], dcy = h[4 * 16 * 16 + 3] - h[4 * 16 * 16 + 1];
      fprintf(f, "# CTA 0 thread 0: %lld cycles in %lld ns = %.3f GHz effective SM clock\n", dcy, dns, dns > 0 ? double(dcy) / double(dns) : 0.0);
      const char* names[4] = {"producer", "mma", "epi_w0", "epi_w4"};
      for (int r = 0; r < 4; ++r)
        for (int i = 0; i < 16; ++i) {
          fprintf(f, "%-8s tile %2d:", names[r], i);
          for (int e = 0; e < 16; ++e) fprintf(f, " %7lld", h[(r * 16 + i) * 16 + e] ? h[(r * 16 + i) * 16 + e] - t0 : -1);
          fprintf(f, "\n");
        }
      fclose(f);
    }
  }
  return rc;
}

}  // namespace

int pick_block_n(int64_t M, int64_t N) {
  // Prefer the widest tile that divides N while still giving every SM work; wide tiles lower the
  // per-flop shared-memory traffic of the single-CTA MMA.
  const int64_t m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  const int sms = num_sms();
  const int cands[3] = {256, 192, 128};
  int best = 0;
  for (int c : cands) {
    if (N % c) continue;
    if (best == 0) best = c;
    if (m_tiles * (N / c) >= sms) return c;
  }
  // small problem: take the narrowest divisor for more parallelism
  for (int i = 2; i >= 0; --i)
    if (N % cands[i] == 0) return cands[i];
  return best;
}

int gemm_bias_act(const SaisGemmArgs& a, cudaStream_t stream, int force_block_n) {
  if (!a.a || !a.w || (!a.out_f32 && !a.out_bf16) || (a.out_f32 && a.out_bf16)) {
    set_last_error("gemm: need A, W and exactly one of out_f32 / out_bf16");
    return kErrInvalidArg;
  }
  if (a.M <= 0 || a.N <= 0 || a.K <= 0 || a.K % BLOCK_K != 0 || (a.N % 128 != 0 && a.N % 192 != 0)) {
    set_last_error("gemm: unsupported shape M=%lld N=%lld K=%lld (need K%%64==0, N%%128==0 or N%%192==0)",
                   (long long)a.M, (long long)a.N, (long long)a.K);
    return kErrShape;
  }
  if (a.lda % 8 || a.ldw % 8 || (a.out_bf16 && a.ldo16 % 8) || (a.out_f32 && a.ldo32 % 4) ||
      (a.residual && a.ldr % 4)) {
    set_last_error("gemm: row pitches must keep 16-byte alignment");
    return kErrInvalidArg;
  }
  if ((reinterpret_cast<uintptr_t>(a.a) | reinterpret_cast<uintptr_t>(a.w) | reinterpret_cast<uintptr_t>(a.out_f32) |
       reinterpret_cast<uintptr_t>(a.out_bf16) | reinterpret_cast<uintptr_t>(a.residual)) & 15) {
    set_last_error("gemm: operands must be 16-byte aligned");
    return kErrInvalidArg;
  }
  if (a.remap_group > 0 && (a.row_add == nullptr || a.split_out)) {
    set_last_error("gemm: remap_group needs row_add and does not support split_out");
    return kErrInvalidArg;
  }
  if (a.split_out && !a.out_bf16) {
    set_last_error("gemm: split_out needs a bf16 output");
    return kErrInvalidArg;
  }
  if (a.act < 0 || a.act > 2) {
    set_last_error("gemm: bad activation %d", a.act);
    return kErrInvalidArg;
  }
  if (a.k_slices < 0 || (a.k_slices > 0 && (!a.out_f32 || a.bias || a.residual || a.act != 0 || a.remap_group ||
                                            a.split_out || a.ln_stats_out || a.ln_stats_in))) {
    set_last_error("gemm: accumulate mode (k_slices >= 1) needs an fp32 output and no bias / residual / activation");
    return kErrInvalidArg;
  }
  if (a.ln_stats_in && (!a.ln_colsum || !a.bias || !a.out_bf16 || a.residual || a.split3 || a.split_out || a.remap_group ||
                        a.act == 2)) {
    set_last_error("gemm: ln_stats_in needs ln_colsum, bias, a bf16 output and act none/GELU (no residual / split / remap)");
    return kErrInvalidArg;
  }
  if (a.ln_stats_out || a.out2_bf16) {
    if (!a.ln_stats_out || !a.out2_bf16 || !a.out_f32 || !a.residual || a.N != 384 || a.act != 0 || a.split_out ||
        a.remap_group || a.ldo2 % 8 || (reinterpret_cast<uintptr_t>(a.out2_bf16) & 15)) {
      set_last_error("gemm: ln_stats_out / out2_bf16 go together and need N = 384, fp32 output + residual, no activation");
      return kErrInvalidArg;
    }
    force_block_n = 192;  // the statistics slots are (n-tile, warp half): exactly two n-tiles per row
  }
  const int bn = force_block_n ? force_block_n : pick_block_n(a.M, a.N);
  if (bn == 0 || a.N % bn) {
    set_last_error("gemm: N=%lld not divisible by tile %d", (long long)a.N, bn);
    return kErrShape;
  }
  int mode = kModeGeneric;
  if (a.remap_group == 0 && !a.split_out && !(a.act == 1 && a.split3)) {
    if (a.out_f32 && a.act == 0) mode = kModeF32;
    else if (a.out_bf16 && !a.residual && a.act == 1) mode = kModeBf16Gelu;
    else if (a.out_bf16 && !a.residual) mode = kModeBf16;
  }
  // CTA pairs (cta_group::2) whenever there are at least two m-tiles per SM-pair's worth of work; tiny problems
  // (the C1-sized temporal head) stay on single CTAs.
  const int64_t m_tiles = (a.M + BLOCK_M - 1) / BLOCK_M;
  const int cg = (m_tiles * (a.N / bn) >= 2 * num_sms()) ? 2 : 1;
  // Epilogue warps: the GELU epilogue (fc1) is instruction-bound, so it runs the lean 16-warp path (72.9 -> 67 us at batch
  // 256); the plain bf16 epilogue (qkv) is faster on 8 warps (48.8 vs 51.1 us).  SAIS_GEMM_EW=8|16 forces one for both (the
  // parity suite pins the two to each other bit for bit).
  static const int env_ew = getenv("SAIS_GEMM_EW") ? atoi(getenv("SAIS_GEMM_EW")) : 0;
  const int ewn = (env_ew == 16 || env_ew == 8) ? env_ew : (mode == kModeBf16Gelu ? 16 : 8);
  // A-stationary variant (SAIS_GEMM_ASTAT=1|0 forces / disables): K = 384 consumer GEMMs at CTA-pair sizes
  static const int env_astat = getenv("SAIS_GEMM_ASTAT") ? atoi(getenv("SAIS_GEMM_ASTAT")) : -1;
  const bool astat_ok = cg == 2 && a.K == 384 && !a.split3 && (mode == kModeBf16 || mode == kModeBf16Gelu) && bn >= 192 &&
                        (ewn == 8 || ewn == 16) && !(mode == kModeBf16Gelu && ewn == 8) && !(mode == kModeBf16 && ewn == 16);
  const bool astat = astat_ok && (env_astat < 0 ? kAStatDefault : env_astat != 0);
  if (astat) {
    if (bn == 256) {
      if (mode == kModeBf16Gelu) return launch_gemm<256, kModeBf16Gelu, 2, 16, true>(a, stream);
      return launch_gemm<256, kModeBf16, 2, 8, true>(a, stream);
    }
    if (mode == kModeBf16Gelu) return launch_gemm<192, kModeBf16Gelu, 2, 16, true>(a, stream);
    return launch_gemm<192, kModeBf16, 2, 8, true>(a, stream);
  }
#define SAIS_GEMM_DISPATCH_CG(BN, CG)                                                              \
  switch (mode) {                                                                                  \
    case kModeBf16:                                                                                \
      return ewn == 16 ? launch_gemm<BN, kModeBf16, CG, 16>(a, stream)                             \
                       : launch_gemm<BN, kModeBf16, CG, 8>(a, stream);                             \
    case kModeBf16Gelu:                                                                            \
      return ewn == 16 ? launch_gemm<BN, kModeBf16Gelu, CG, 16>(a, stream)                         \
                       : launch_gemm<BN, kModeBf16Gelu, CG, 8>(a, stream);                         \
    case kModeF32: return launch_gemm<BN, kModeF32, CG, 8>(a, stream);                             \
    default: return launch_gemm<BN, kModeGeneric, CG, 8>(a, stream);                               \
  }
#define SAIS_GEMM_DISPATCH(BN) \
  if (cg == 2) { SAIS_GEMM_DISPATCH_CG(BN, 2) } else { SAIS_GEMM_DISPATCH_CG(BN, 1) }
  switch (bn) {
    case 256: SAIS_GEMM_DISPATCH(256)
    case 192: SAIS_GEMM_DISPATCH(192)
    case 128: SAIS_GEMM_DISPATCH(128)
  }
#undef SAIS_GEMM_DISPATCH
#undef SAIS_GEMM_DISPATCH_CG
  set_last_error("gemm: bad tile %d", bn);
  return kErrShape;
}

}  // namespace sais

// Full-row GEMM with the residual add AND the following LayerNorm fused into its epilogue (sm_100a):
//
//     x  <-  x + A · Wᵀ + bias            (fp32 residual stream, in place)
//     xn <-  LayerNorm(x; gamma, beta)     (bf16 operand of the next GEMM)          N = 384 = one full row per tile
//
// Reference: `x = x + self.attn(self.norm1(x))` / `x = x + self.mlp(self.norm2(x))` followed by the next
// `self.norm2(x)` / `self.norm1(x)` of SAIS/scripts/dino-main/vision_transformer.py:103-108 — proj + norm2 and
// fc2 + (next block's) norm1.  Unfused, each LayerNorm is a separate pass that re-reads the 77 MB fp32 stream from HBM
// and writes 39 MB of bf16 (23 us per pass at batch 256, 12.5% of the step); here the row is normalised while it is
// still in tensor memory.
//
// A CTA PAIR (2-CTA cluster, tcgen05 cta_group::2) owns 256 rows x all 384 columns: the accumulator is two N = 192
// halves in TMEM columns [0,384), fed from a 4-stage TMA ring (A k-block 16 KB + both W halves 24 KB per stage and
// CTA — sharing one A k-block between the two halves cuts the L2->SM operand traffic per flop by 30% against the
// 256x192 tiles of the generic kernel).
//   warps 0..15 : epilogue, four warps per TMEM lane quarter, 96 columns each (thread = row):
//                 pass 1 : acc + bias + residual (TMA-prefetched 32x16 fp32 tiles) -> x (TMA store), the sum back into
//                          TMEM in place, row sum
//                 pass 1b: exact two-pass variance straight from TMEM (torch's biased variance)
//                 pass 2 : (v - mean) * rstd * gamma + beta -> bf16 -> TMA store
//   warp 16     : TMA producer (+ L2 prefetch of the tile's residual rows one mainloop ahead)
//   warp 17     : TMEM allocator + MMA issuer (leader CTA)
// The accumulator is single-buffered (384 of 512 columns), so the epilogue of a tile is exposed; it is bounded by the
// L2 traffic of the residual read and the two stores, which is the HBM roofline of this op anyway.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace sais {

namespace {

constexpr int MT = 128;   // rows per CTA
constexpr int DM = 384;   // N (full row)
constexpr int NH = 192;   // MMA N (half row)
constexpr int kEpiWarps = 16;
constexpr int kThreads = 32 * (kEpiWarps + 2);
constexpr int A_STAGE = MT * 128;          // 16,384
constexpr int W_HALF = (NH / 2) * 128;     // 12,288: 96 rows (this CTA's share of one N half) x 128 B
constexpr int STAGE_BYTES = A_STAGE + 2 * W_HALF;  // 40,960
constexpr int NSTAGE = 4;
constexpr int kStageBuf = 2048;            // 32 rows x 64 B (16 fp32 or 32 bf16 columns), SWIZZLE_64B
constexpr int kStagingBytes = kEpiWarps * 2 * kStageBuf;  // 65,536
constexpr int kBarBytes = 512;
constexpr int kStatBytes = 4 * MT * 4;     // partial row statistics of the 4 column parts
constexpr int kSmemUsed = NSTAGE * STAGE_BYTES + kStagingBytes + kBarBytes + kStatBytes;  // 231,936
constexpr int kSmemBytes = 227 * 1024;
static_assert(kSmemUsed <= kSmemBytes, "shared memory budget");
constexpr int kTmemCols = 512;
constexpr int CPW = DM / 4;                // 96 columns per epilogue warp
constexpr int NC1 = CPW / 16;              // 6 pass-1 chunks of 16 columns
constexpr int NC2 = CPW / 32;              // 3 pass-2 chunks of 32 columns

struct RowLnParams {
  const float* bias;
  const float* gamma;
  const float* beta;
  float eps;
  int num_tiles;  // 256-row tiles
  int k_blocks;   // K / 64
  int do_ln;      // 0: GEMM + residual only
  long long* dbg; // dev knob (SAIS_ROWLN_TIMELINE=<file>): CTA 0 records clock64() per role / tile / event
};

__device__ __forceinline__ void sts128r(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds128r(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void tma_load_2d_u(uint32_t smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d_u(const CUtensorMap* m, uint32_t smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* m, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
               :
               : "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void quarter_bar_sync(int q) {
  asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
}
// byte offset of 16-byte chunk j of 64-byte row r in a SWIZZLE_64B staging tile
__device__ __forceinline__ uint32_t sw64_off(int r, int j) { return uint32_t(r * 64 + ((j ^ ((r >> 1) & 3)) << 4)); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_rowln_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                  const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_xn,
                  const __grid_constant__ CUtensorMap tmap_xpf, const RowLnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (threadIdx.x == 0 && (smem - smem_raw) + kSmemUsed > kSmemBytes) __trap();  // dynamic smem base is 1 KB aligned in practice
  uint8_t* staging = smem + NSTAGE * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + kStagingBytes);
  uint64_t* full_bar = bars;              // [NSTAGE] leader's copy counts
  uint64_t* empty_bar = bars + NSTAGE;    // [NSTAGE]
  uint64_t* acc_full = bars + 8;
  uint64_t* acc_empty = bars + 9;         // leader's, 2 * kEpiWarps arrivals
  uint64_t* res_bar = bars + 10;          // [kEpiWarps][2]
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 10 + 2 * kEpiWarps);
  float* stats = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + kBarBytes);  // [4][128]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;
  auto stamp = [&](int role, int idx, int ev) {
    if (p.dbg != nullptr && blockIdx.x == 0 && idx < 8 && ev < 8) p.dbg[(role * 8 + idx) * 8 + ev] = clock64();
  };

  constexpr int kProducerWarp = kEpiWarps, kMmaWarp = kEpiWarps + 1;
  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_xn);
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 2 * kEpiWarps);
    for (int s = 0; s < 2 * kEpiWarps; ++s) mbar_init(&res_bar[s], 1);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc_cg2(tmem_base_smem, kTmemCols);
    tmem_relinquish_cg2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  if (warp == kProducerWarp) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < p.num_tiles; tile += npairs) {
        const int m0 = tile * (2 * MT) + int(crank) * MT;
        // the epilogue of this tile reads these residual rows one mainloop from now: pull them into L2
#pragma unroll
        for (int c = 0; c < DM; c += 96) tma_prefetch_l2_2d(&tmap_xpf, c, m0);
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (crank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
          const uint32_t lbar = leader_smem_u32(&full_bar[stage]);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          tma_load_2d_cg2(sa, &tmap_a, lbar, kb * 64, m0);
          tma_load_2d_cg2(sa + A_STAGE, &tmap_w, lbar, kb * 64, int(crank) * (NH / 2));
          tma_load_2d_cg2(sa + A_STAGE + W_HALF, &tmap_w, lbar, kb * 64, NH + int(crank) * (NH / 2));
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (leader CTA) =====================
    if (lane == 0 && crank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * MT, NH);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t ti = 0;
      for (int tile = pair; tile < p.num_tiles; tile += npairs, ++ti) {
        stamp(0, ti, 0);
        mbar_wait(acc_empty, (ti & 1) ^ 1);  // the previous tile's epilogue has finished with the accumulator
        tc_fence_after();
        stamp(0, ti, 1);
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t da = umma_desc_sw128_kmajor(sa);
          // all four K steps of one half back to back: consecutive MMAs into the same accumulator run at full rate,
          // every switch of accumulator costs ~45 cycles (tools/mma_alt_bench.cu)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint64_t db = umma_desc_sw128_kmajor(sa + A_STAGE + h * W_HALF);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_cg2(tmem_base + h * NH, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit_cg2_mcast(&empty_bar[stage], uint16_t(0b11));
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
        umma_commit_cg2_mcast(acc_full, uint16_t(0b11));
        stamp(0, ti, 2);
      }
    }
  } else {
    // ===================== epilogue warps 0..15 =====================
    const int ew = warp;
    const int q = ew & 3;      // TMEM lane quarter
    const int part = ew >> 2;  // which 96 columns
    const int col0 = part * CPW;
    const int row = q * 32 + lane;
    const uint32_t t_lane = tmem_base + (uint32_t(q * 32) << 16);
    const uint32_t my_stage = smem_u32(staging) + ew * (2 * kStageBuf);
    uint64_t* my_res_bar = res_bar + 2 * ew;
    uint32_t it = 0;  // pass-1 chunks processed (selects staging buffer and residual barrier phase)
    uint32_t ti = 0;

    if (lane == 0 && pair < p.num_tiles) {  // prime the residual pipeline
      mbar_arrive_expect_tx(&my_res_bar[0], kStageBuf);
      tma_load_2d_u(my_stage, &tmap_x, &my_res_bar[0], col0, pair * (2 * MT) + int(crank) * MT + q * 32);
    }

    for (int tile = pair; tile < p.num_tiles; tile += npairs, ++ti) {
      const int m0 = tile * (2 * MT) + int(crank) * MT + q * 32;  // first row of this warp
      const bool st_on = (ew == 0 && lane == 0);
      if (st_on) stamp(1, ti, 0);
      mbar_wait(acc_full, ti & 1);
      tc_fence_after();
      if (st_on) stamp(1, ti, 1);

      // ---------------- pass 1: v = acc + bias + residual -> x ; v back into TMEM ; row sum
      float sum = 0.f;
      uint32_t v[16];
      tmem_ld_32x16(t_lane + col0, v);
#pragma unroll 1
      for (int c = 0; c < NC1; ++c) {
        const int col = col0 + c * 16;
        float f[16];
        tmem_ld_wait_dep(v);
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
#pragma unroll
        for (int j = 0; j < 16; j += 8)
          asm volatile("" : "+f"(f[j]), "+f"(f[j + 1]), "+f"(f[j + 2]), "+f"(f[j + 3]), "+f"(f[j + 4]), "+f"(f[j + 5]),
                            "+f"(f[j + 6]), "+f"(f[j + 7]));
        if (c + 1 < NC1) tmem_ld_32x16(t_lane + col + 16, v);
        {
          const float4* bp = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 b4 = __ldg(bp + j);
            f[4 * j] += b4.x; f[4 * j + 1] += b4.y; f[4 * j + 2] += b4.z; f[4 * j + 3] += b4.w;
          }
        }
        const uint32_t buf = my_stage + (it & 1) * kStageBuf;
        mbar_wait(&my_res_bar[it & 1], (it >> 1) & 1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 r4 = lds128r(buf + sw64_off(lane, j));
          f[4 * j] += r4.x; f[4 * j + 1] += r4.y; f[4 * j + 2] += r4.z; f[4 * j + 3] += r4.w;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) sum += f[j];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          sts128r(buf + sw64_off(lane, j), __float_as_uint(f[4 * j]), __float_as_uint(f[4 * j + 1]),
                  __float_as_uint(f[4 * j + 2]), __float_as_uint(f[4 * j + 3]));
        if (p.do_ln) {
          uint32_t w[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) w[j] = __float_as_uint(f[j]);
          tmem_st_32x16(t_lane + col, w);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          // all earlier stores have finished reading smem -> the other buffer is free: prefetch the next residual tile
          tma_store_wait_read<0>();
          int nt = tile, nc = c + 1;
          if (nc == NC1) {
            nt = tile + npairs;
            nc = 0;
          }
          if (nt < p.num_tiles) {
            uint64_t* rb = &my_res_bar[(it + 1) & 1];
            mbar_arrive_expect_tx(rb, kStageBuf);
            tma_load_2d_u(my_stage + ((it + 1) & 1) * kStageBuf, &tmap_x, rb, col0 + nc * 16,
                          nt * (2 * MT) + int(crank) * MT + q * 32);
          }
          tma_store_2d_u(&tmap_x, buf, col, m0);
          tma_store_commit();
        }
        ++it;
        if (st_on && c < 6) stamp(2, ti, c);
      }
      if (st_on) stamp(1, ti, 2);

      if (!p.do_ln) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(leader_smem_u32(acc_empty));
        continue;
      }
      tmem_st_wait();

      // ---------------- row statistics: exact two-pass mean / biased variance over the four column parts
      stats[part * MT + row] = sum;
      quarter_bar_sync(q);
      const float mean = ((stats[row] + stats[MT + row]) + (stats[2 * MT + row] + stats[3 * MT + row])) * (1.0f / DM);
      float ss = 0.f;
      {
        uint32_t a[32], b[32];
        tmem_ld_32x32(t_lane + col0, a);
        tmem_ld_32x32(t_lane + col0 + 32, b);
        tmem_ld_wait_dep(a);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float d = __uint_as_float(a[j]) - mean;
          ss = fmaf(d, d, ss);
        }
        tmem_ld_wait_dep(b);
        asm volatile("" : "+f"(ss));
        tmem_ld_32x32(t_lane + col0 + 64, a);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float d = __uint_as_float(b[j]) - mean;
          ss = fmaf(d, d, ss);
        }
        tmem_ld_wait_dep(a);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float d = __uint_as_float(a[j]) - mean;
          ss = fmaf(d, d, ss);
        }
      }
      quarter_bar_sync(q);  // everyone has read the sums
      stats[part * MT + row] = ss;
      quarter_bar_sync(q);
      const float var = ((stats[row] + stats[MT + row]) + (stats[2 * MT + row] + stats[3 * MT + row])) * (1.0f / DM);
      const float rstd = rsqrtf(var + p.eps);
      const float nmr = -mean * rstd;
      if (st_on) stamp(1, ti, 3);

      // ---------------- pass 2: normalise -> bf16 -> xn (single staging buffer: the other one holds the prefetched
      // residual of the next tile)
      const uint32_t buf2 = my_stage + ((it + 1) & 1) * kStageBuf;  // buffer of the last pass-1 chunk
      {
        uint32_t a[32];
        tmem_ld_32x32(t_lane + col0, a);
#pragma unroll 1
        for (int c = 0; c < NC2; ++c) {
          const int col = col0 + c * 32;
          float f[32];
          tmem_ld_wait_dep(a);
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaf(__uint_as_float(a[j]), rstd, nmr);
#pragma unroll
          for (int j = 0; j < 32; j += 8)
            asm volatile("" : "+f"(f[j]), "+f"(f[j + 1]), "+f"(f[j + 2]), "+f"(f[j + 3]), "+f"(f[j + 4]), "+f"(f[j + 5]),
                              "+f"(f[j + 6]), "+f"(f[j + 7]));
          if (c + 1 < NC2) {
            tmem_ld_32x32(t_lane + col + 32, a);
          } else {  // the accumulator columns of this warp are dead: hand them back to the MMA issuer
            tc_fence_before();
            if (lane == 0) mbar_arrive_cluster(leader_smem_u32(acc_empty));
          }
          const float4* gp = reinterpret_cast<const float4*>(p.gamma + col);
          const float4* bp = reinterpret_cast<const float4*>(p.beta + col);
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 g4 = __ldg(gp + j), b4 = __ldg(bp + j);
            pk[2 * j] = pack_bf16x2(fmaf(f[4 * j], g4.x, b4.x), fmaf(f[4 * j + 1], g4.y, b4.y));
            pk[2 * j + 1] = pack_bf16x2(fmaf(f[4 * j + 2], g4.z, b4.z), fmaf(f[4 * j + 3], g4.w, b4.w));
          }
          if (lane == 0) tma_store_wait_read<0>();  // the previous store from this buffer has read it
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts128r(buf2 + sw64_off(lane, j), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d_u(&tmap_xn, buf2, col, m0);
            tma_store_commit();
          }
        }
      }
      if (st_on) stamp(1, ti, 4);
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

int gemm_residual_layernorm(const sais_bf16* a, int64_t lda, const sais_bf16* w, int64_t ldw, const float* bias,
                            float* x, const float* gamma, const float* beta, float eps, sais_bf16* xn, int64_t M,
                            int64_t K, cudaStream_t stream) {
  if (M == 0) return kOk;
  if (!a || !w || !x || !bias || M < 0 || K <= 0 || K % 64 != 0 || (xn && (!gamma || !beta))) {
    set_last_error("gemm_residual_layernorm: bad arguments (need K %% 64 == 0, bias, and gamma/beta with xn)");
    return kErrInvalidArg;
  }
  if (lda % 8 || ldw % 8 ||
      ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(x) |
        reinterpret_cast<uintptr_t>(xn) | reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(gamma) |
        reinterpret_cast<uintptr_t>(beta)) & 15)) {
    set_last_error("gemm_residual_layernorm: operands must be 16-byte aligned");
    return kErrInvalidArg;
  }
  CUtensorMap ta, tw, tx, txn, tpf;
  int rc = make_tmap_2d(&ta, a, kTmapBf16, uint64_t(M), uint64_t(K), uint64_t(lda), MT, 64, 128);
  if (rc) return rc;
  if ((rc = make_tmap_2d(&tw, w, kTmapBf16, DM, uint64_t(K), uint64_t(ldw), NH / 2, 64, 128))) return rc;
  if ((rc = make_tmap_2d(&tx, x, kTmapF32, uint64_t(M), DM, DM, 32, 16, 64))) return rc;
  if ((rc = make_tmap_2d(&tpf, x, kTmapF32, uint64_t(M), DM, DM, MT, 96, 0))) return rc;
  txn = tx;
  if (xn && (rc = make_tmap_2d(&txn, xn, kTmapBf16, uint64_t(M), DM, DM, 32, 32, 64))) return rc;

  static bool attr_set = false;
  if (!attr_set) {
    rc = check_cuda(cudaFuncSetAttribute(gemm_rowln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes),
                    "cudaFuncSetAttribute(gemm_rowln)");
    if (rc) return rc;
    attr_set = true;
  }
  RowLnParams p;
  p.bias = bias;
  p.gamma = gamma;
  p.beta = beta;
  p.eps = eps;
  p.num_tiles = int((M + 2 * MT - 1) / (2 * MT));
  p.k_blocks = int(K / 64);
  p.do_ln = xn != nullptr;
  static const char* timeline = getenv("SAIS_ROWLN_TIMELINE");
  p.dbg = nullptr;
  constexpr int kDbgN = 3 * 8 * 8;
  if (timeline) {
    if (cudaMalloc(&p.dbg, kDbgN * sizeof(long long)) != cudaSuccess) p.dbg = nullptr;
    if (p.dbg) cudaMemsetAsync(p.dbg, 0, kDbgN * sizeof(long long), stream);
  }
  const int pairs_max = num_sms() / 2;
  const int pairs = p.num_tiles < pairs_max ? p.num_tiles : pairs_max;
  {
    LaunchScope ls(kClsGemm, stream, 2.0 * double(M) * DM * double(K));
    gemm_rowln_kernel<<<2 * pairs, kThreads, kSmemBytes, stream>>>(ta, tw, tx, txn, tpf, p);
    rc = check_cuda(cudaGetLastError(), "gemm_rowln_kernel launch");
  }
  if (p.dbg) {
    static long long h[kDbgN];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, p.dbg, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(p.dbg);
    long long t0 = 0;
    for (long long v : h) if (v && (!t0 || v < t0)) t0 = v;
    if (FILE* f = fopen(timeline, "w")) {
      fprintf(f, "# M=%lld K=%lld tiles=%d pairs=%d ln=%d\n", (long long)M, (long long)K, p.num_tiles, pairs, p.do_ln);
      const char* names[3] = {"mma(tile: start, acc_empty ok, all issued)", "epi w0(tile: wait, acc_full, pass1 end, stats end, pass2 end)",
                              "epi w0 pass-1 chunk ends"};
      for (int r = 0; r < 3; ++r) {
        fprintf(f, "%s\n", names[r]);
        for (int i = 0; i < 8; ++i) {
          fprintf(f, "  %2d:", i);
          for (int e = 0; e < 8; ++e) fprintf(f, " %8lld", h[(r * 8 + i) * 8 + e] ? h[(r * 8 + i) * 8 + e] - t0 : -1);
          fprintf(f, "\n");
        }
      }
      fclose(f);
    }
  }
  return rc;
}

}  // namespace sais

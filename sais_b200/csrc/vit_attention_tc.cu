// tcgen05 flash-style self-attention for the ViT blocks: T = 197 tokens, 6 heads x 64
// (SAIS/scripts/dino-main/vision_transformer.py:80-92), one (frame, head) item at a time per persistent CTA.
//
//   warp 8 (one thread): TMA loads of Q, K, V (3-D tensor map [frame, token, 1152]; tokens beyond 196 are
//                        zero-filled by TMA) into a double-buffered smem slot, then
//                        S = Q K^T   : tcgen05.mma  M=128 (x2 row tiles), N=208, K=64,  fp32 in TMEM
//                        O = P V     : tcgen05.mma  M=128 (x2),           N=64,  K=208, V consumed MN-major
//   warps 0..7         : one thread per query row: exact softmax straight out of TMEM (two passes: max, then
//                        exp2 / sum); P is rounded to bf16 and written back INTO TMEM over the dead logits
//                        (tcgen05.st) where the second MMA reads it as its A operand, so probabilities touch
//                        neither shared memory nor HBM; finally O / sum -> bf16 -> global.
// TMEM map per row tile mt (256-column window): S fp32 [0,208) -> P bf16x2 [0,104) + O fp32 [128,192).
// kEmitProbs (get_last_selfattention with precision='bf16', vision_transformer.py:216-223): the softmax warps also write
// the normalised probabilities exp(s - max) / sum as fp32 [B,6,197,197]; the row sum is then taken in an extra pass over
// the logits before the exp pass, so that what is stored is already normalised (a visualiser path: one block, no timing
// pressure).  The exp pass of the two row tiles of an item alternates per lane quarter (`turn` barriers): one tile's
// exponentials hide the other's MMA round trips.  (Experiments that did not pay — P through shared memory, no turns,
// part of the exponentials as a polynomial on the FMA pipe — were removed from the product; DESIGN.md 3.4 has the numbers.)
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace sais {

namespace {

constexpr int T = 197;
constexpr int HEADS = 6;
constexpr int HD = 64;
constexpr int KEYS = 208;                  // keys padded to a multiple of 16 (UMMA N / K granularity)
constexpr int MAT_BYTES = KEYS * 128;      // one of Q / K / V in smem: 208 rows x 128 B (SWIZZLE_128B)
constexpr int SLOT_BYTES = 3 * MAT_BYTES;  // Q, K, V of one item
constexpr int kThreads = 11 * 32;  // 8 softmax warps + loader + one MMA-issuing warp per row-tile window
constexpr int kTmemCols = 512;             // row tile 0 at column 0, row tile 1 at column 256
constexpr int kOCol = 128;                 // O accumulator inside the row tile's window

constexpr int kOutStage = 8 * 4096;        // per softmax warp: 32 rows x 128 B staging tile for the TMA store of O
constexpr int kSmemBytes = 2 * SLOT_BYTES + 16384 /*overrun pad for the 2nd Q row tile*/ + kOutStage + 1024 + 256;

__device__ __forceinline__ void sts128u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <bool kEmitProbs>
__global__ void __launch_bounds__(kThreads, 1)
vit_attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_out,
                        float* __restrict__ probs, int n_items, int flags, long long* __restrict__ dbg) {
  const int reverse = flags & 1;  // (kernels.h g_tile_reverse)
  // dev knob (SAIS_ATTN_TIMELINE=<file>): CTA 0 records clock64() at every phase boundary, [role][item][event]
  auto stamp = [&](int role, int idx, int ev) {
    if (dbg != nullptr && blockIdx.x == 0 && idx < 16) dbg[(role * 16 + idx) * 8 + ev] = clock64();
  };
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* p_smem = smem + 2 * SLOT_BYTES;
  uint8_t* o_stage = p_smem + 16384;  // (16 KB pad: the second Q row tile's UMMA descriptor may run past row 207)
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_smem + 16384 + kOutStage);
  uint64_t* ld_full = bars;       // [2] item slot loaded
  uint64_t* ld_empty = bars + 2;  // [2] item slot free again
  uint64_t* s_full = bars + 4;    // [2] S[mt] in TMEM
  uint64_t* p_full = bars + 6;    // [2] P[mt] ready
  uint64_t* o_full = bars + 8;    // [2] O[mt] in TMEM (also: P no longer read by the tensor core)
  uint64_t* t_free = bars + 10;   // [2] TMEM window of row tile mt drained
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 12);
  uint64_t* turn = bars + 14;     // [2][4] exp-pass turn of (row tile, lane quarter): the two warps of a scheduler alternate

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ld_full[i], 1);
      mbar_init(&ld_empty[i], 2);  // both windows' MMA streams must have retired
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
      mbar_init(&t_free[i], 4);
      for (int j = 0; j < 4; ++j) mbar_init(&turn[i * 4 + j], 1);
    }
    fence_mbar_init();
  }
  if (warp == 8) {
    tmem_alloc(tmem_base_smem, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  pdl_wait();  // (PDL) set-up done under the previous kernel's tail; global memory only from here on

  if (warp >= 8) {
    // ===================== single-thread roles: loader (warp 8), MMA issuer of window 0 / 1 (warps 9 / 10) ======
    // Each role blocks on its own mbarriers, so the two row-tile windows advance independently: window w may
    // already run S = QK^T of the next item while the other window is still in its softmax.
    // All 32 lanes of a role warp walk its loop (warp-uniform control flow and values); only the TMA / tcgen05
    // instructions sit under elect_one().  Run by one divergent lane, every tcgen05.mma cost an ELECT / R2UR.BROADCAST
    // "waterfall" plus 64-bit vector address arithmetic (~200 cycles per instruction, more than the 13 N = 64 MMAs of
    // O = P V take on the tensor pipe).
    {
      const int n_my = (n_items - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
      if (warp == 8) {
        for (int idx = 0; idx < n_my; ++idx) {
          const int slot = idx & 1;
          if (idx >= 2) mbar_wait(&ld_empty[slot], ((idx - 2) >> 1) & 1);  // previous occupant fully consumed
          const int item0 = int(blockIdx.x) + idx * int(gridDim.x);
          const int item = reverse ? n_items - 1 - item0 : item0;  // (kernels.h g_tile_reverse)
          const int b = item / HEADS, h = item % HEADS;
          uint8_t* dst = smem + slot * SLOT_BYTES;
          if (elect_one()) {
            mbar_arrive_expect_tx(&ld_full[slot], SLOT_BYTES);
            tma_load_3d(dst, &tmap_qkv, &ld_full[slot], h * HD, 0, b);
            tma_load_3d(dst + MAT_BYTES, &tmap_qkv, &ld_full[slot], 384 + h * HD, 0, b);
            tma_load_3d(dst + 2 * MAT_BYTES, &tmap_qkv, &ld_full[slot], 768 + h * HD, 0, b);
          }
          __syncwarp();
        }
      } else {
        constexpr uint32_t idesc_s = umma_idesc_bf16(128, KEYS, 0, 0);  // Q (K-major) x K (K-major)
        constexpr uint32_t idesc_o = umma_idesc_bf16(128, HD, 0, 1);    // P (K-major) x V (MN-major)
        const int w = warp - 9;
        const uint32_t win = tmem_base + w * 256;

        for (int idx = 0; idx < n_my; ++idx) {
          const int slot = idx & 1;
          const uint32_t q_s = smem_u32(smem + slot * SLOT_BYTES);
          const uint32_t k_s = q_s + MAT_BYTES, v_s = q_s + 2 * MAT_BYTES;
          mbar_wait(&ld_full[slot], (idx >> 1) & 1);
          if (lane == 0) stamp(2 + w, idx, 0);
          mbar_wait(&t_free[w], (idx & 1) ^ 1);  // window drained by the softmax warps
          tc_fence_after();
          if (lane == 0) stamp(2 + w, idx, 1);
          {
            const uint64_t da = umma_desc_sw128_kmajor(q_s + w * 128 * 128);
            const uint64_t db = umma_desc_sw128_kmajor(k_s);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < HD / 16; ++k) umma_f16(win, da + 2 * k, db + 2 * k, idesc_s, k != 0);
              umma_commit(&s_full[w]);
            }
            __syncwarp();
          }
          if (lane == 0) stamp(2 + w, idx, 2);
          mbar_wait(&p_full[w], idx & 1);
          tc_fence_after();
          if (lane == 0) stamp(2 + w, idx, 3);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < KEYS / 16; ++ks) {
              const uint64_t db = umma_desc_sw128_mnmajor(v_s + ks * 2048, 0);
              umma_f16_ts(win + kOCol, win + ks * 8, db, idesc_o, ks != 0);
            }
            umma_commit(&o_full[w]);
            umma_commit(&ld_empty[slot]);  // this window's reads of the slot have retired (barrier counts both windows)
          }
          __syncwarp();
          if (lane == 0) stamp(2 + w, idx, 4);
        }
      }
    }
  } else {
    // ===================== softmax / output warps =====================
    const int mt = warp >> 2, q = warp & 3;
    const int row = mt * 128 + q * 32 + lane;  // query row inside the frame
    const bool warp_has_rows = (mt * 128 + q * 32) < T;
    const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + mt * 256;
    const int sw = lane & 7;
    const float sl2 = 0.125f * 1.4426950408889634f;  // head_dim^-0.5 * log2(e)

    int it = 0;
    for (int item0 = blockIdx.x; item0 < n_items; item0 += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      const int item = reverse ? n_items - 1 - item0 : item0;
      const int b = item / HEADS, h = item % HEADS;
      const bool st_on = (q == 0 && lane == 0);
      if (st_on) stamp(mt, it, 0);
      mbar_wait(&s_full[mt], ph);
      tc_fence_after();
      if (st_on) stamp(mt, it, 1);
      float row_sum = 1.0f;
      if (warp_has_rows) {
        // pass 1: row maximum over the 197 real keys (next chunk streams in while the current one is reduced)
        float m = -INFINITY;
        {
          uint32_t v[32], w[32];
          tmem_ld_32x32(t_row, v);
#pragma unroll
          for (int c = 0; c < 6; c += 2) {
            tmem_ld_wait_dep(v);
            tmem_ld_32x32(t_row + (c + 1) * 32, w);  // streams in while v is reduced
            float mv = m;
#pragma unroll
            for (int j = 0; j < 32; ++j) mv = fmaxf(mv, __uint_as_float(v[j]));
            asm volatile("" : "+f"(mv));  // v fully consumed before it is handed to the next asynchronous load
            tmem_ld_wait_dep(w);
            if (c + 2 < 6) tmem_ld_32x32(t_row + (c + 2) * 32, v); else tmem_ld_32x16(t_row + 192, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) mv = fmaxf(mv, __uint_as_float(w[j]));
            asm volatile("" : "+f"(mv));
            m = mv;
          }
          tmem_ld_wait_dep(v);
#pragma unroll
          for (int j = 0; j < T - 192; ++j) m = fmaxf(m, __uint_as_float(v[j]));
        }
        const float mb = m * sl2;
        if (st_on) stamp(mt, it, 2);
        // The exp pass is MUFU-bound (208 ex2 per row at 4 /clk per scheduler) and the two warps of a scheduler (row tile
        // 0 and 1 of the same lane quarter) would otherwise run it at the same time at half speed each and leave the
        // MUFU idle while both wait for their MMAs: take turns, so one tile's exponentials hide the other's MMA round trips.
        mbar_wait(&turn[mt * 4 + q], (mt == 0) ? (ph ^ 1) : ph);
        // (probabilities requested) the row sum first, with the very exponentials pass 2 evaluates, so that pass 2 can
        // store exp / sum directly
        float inv_emit = 0.f;
        float* prow = nullptr;
        if (kEmitProbs) {
          float se = 0.f;
          uint32_t v[32];
#pragma unroll 1
          for (int c = 0; c < 6; ++c) {
            tmem_ld_32x32(t_row + c * 32, v);
            tmem_ld_wait_dep(v);
#pragma unroll
            for (int j = 0; j < 16; ++j)
              se += ex2_approx(fmaf(__uint_as_float(v[2 * j]), sl2, -mb)) + ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), sl2, -mb));
          }
          tmem_ld_32x16(t_row + 192, v);
          tmem_ld_wait_dep(v);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float p0 = (2 * j < T - 192) ? ex2_approx(fmaf(__uint_as_float(v[2 * j]), sl2, -mb)) : 0.0f;
            const float p1 = (2 * j + 1 < T - 192) ? ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), sl2, -mb)) : 0.0f;
            se += p0 + p1;
          }
          inv_emit = 1.0f / se;
          if (row < T) prow = probs + ((int64_t(b) * HEADS + h) * T + row) * T;
        }
        // pass 2: p = exp2((s - max) * scale), row sum, bf16 P
        float sum = 0.f;
        auto softmax_chunk = [&](const uint32_t(&v)[32], int c) {
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * j]), sl2, -mb));
            const float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), sl2, -mb));
            sum += p0 + p1;
            pk[j] = pack_bf16x2(p0, p1);
            if (kEmitProbs && prow != nullptr) {  // (rows are 788 bytes apart: 4-byte stores; not a hot path)
              prow[c * 32 + 2 * j] = p0 * inv_emit;
              prow[c * 32 + 2 * j + 1] = p1 * inv_emit;
            }
          }
          tmem_st_32x16(t_row + c * 16, pk);  // over logits this thread has already consumed
        };
        {
          uint32_t v[32], w[32];
          tmem_ld_32x32(t_row, v);
#pragma unroll 1
          for (int c = 0; c < 6; c += 2) {
            tmem_ld_wait_dep(v);
            tmem_ld_32x32(t_row + (c + 1) * 32, w);  // next chunk streams in under this chunk's exponentials
            softmax_chunk(v, c);
            asm volatile("" : "+f"(sum));
            tmem_ld_wait_dep(w);
            if (c + 2 < 6) tmem_ld_32x32(t_row + (c + 2) * 32, v); else tmem_ld_32x16(t_row + 192, v);
            softmax_chunk(w, c + 1);
            asm volatile("" : "+f"(sum));
          }
          tmem_ld_wait_dep(v);
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float p0 = (2 * j < T - 192) ? ex2_approx(fmaf(__uint_as_float(v[2 * j]), sl2, -mb)) : 0.0f;
            const float p1 = (2 * j + 1 < T - 192) ? ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), sl2, -mb)) : 0.0f;
            sum += p0 + p1;
            pk[j] = pack_bf16x2(p0, p1);
            if (kEmitProbs && prow != nullptr) {
              if (2 * j < T - 192) prow[192 + 2 * j] = p0 * inv_emit;
              if (2 * j + 1 < T - 192) prow[192 + 2 * j + 1] = p1 * inv_emit;
            }
          }
          tmem_st_32x8(t_row + 96, pk);
        }
        row_sum = sum;
        if (lane == 0) mbar_arrive(&turn[(mt ^ 1) * 4 + q]);
        tmem_st_wait();
      } else {
        // rows 224..255 do not exist: pass the turn straight on, keep the barrier protocol, skip the math (their P rows
        // stay stale / unused)
        mbar_wait(&turn[mt * 4 + q], (mt == 0) ? (ph ^ 1) : ph);
        if (lane == 0) mbar_arrive(&turn[(mt ^ 1) * 4 + q]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[mt]);
      if (st_on) stamp(mt, it, 3);

      // ---- O[mt] / sum -> bf16 -> global ----
      mbar_wait(&o_full[mt], ph);
      tc_fence_after();
      if (st_on) stamp(mt, it, 4);
      if (warp_has_rows) {
        const float inv = 1.0f / row_sum;
        uint32_t v[32], w[32];
        tmem_ld_32x32(t_row + kOCol, v);
        tmem_ld_32x32(t_row + kOCol + 32, w);
        tmem_ld_wait_dep(v);
        tmem_ld_wait_dep(w);
        // O is in registers: the whole TMEM window of this row tile is dead, so hand it back BEFORE the staging / store
        // work — the next item's S = QK^T then runs under this epilogue instead of after it
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_free[mt]);
        {
          // O tile of this warp (32 rows x 64) -> swizzled staging -> one TMA store; rows beyond token 196 are
          // clipped by the [frame, token, 384] tensor map
          const uint32_t st = smem_u32(o_stage) + warp * 4096 + lane * 128;
          if (lane == 0) tma_store_wait_read<0>();  // previous item's store has finished reading the tile
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts128u(st + ((j ^ sw) << 4),
                    pack_bf16x2(__uint_as_float(v[8 * j]) * inv, __uint_as_float(v[8 * j + 1]) * inv),
                    pack_bf16x2(__uint_as_float(v[8 * j + 2]) * inv, __uint_as_float(v[8 * j + 3]) * inv),
                    pack_bf16x2(__uint_as_float(v[8 * j + 4]) * inv, __uint_as_float(v[8 * j + 5]) * inv),
                    pack_bf16x2(__uint_as_float(v[8 * j + 6]) * inv, __uint_as_float(v[8 * j + 7]) * inv));
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts128u(st + (((j + 4) ^ sw) << 4),
                    pack_bf16x2(__uint_as_float(w[8 * j]) * inv, __uint_as_float(w[8 * j + 1]) * inv),
                    pack_bf16x2(__uint_as_float(w[8 * j + 2]) * inv, __uint_as_float(w[8 * j + 3]) * inv),
                    pack_bf16x2(__uint_as_float(w[8 * j + 4]) * inv, __uint_as_float(w[8 * j + 5]) * inv),
                    pack_bf16x2(__uint_as_float(w[8 * j + 6]) * inv, __uint_as_float(w[8 * j + 7]) * inv));
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                         :
                         : "l"(reinterpret_cast<uint64_t>(&tmap_out)), "r"(smem_u32(o_stage) + warp * 4096),
                           "r"(h * HD), "r"(mt * 128 + q * 32), "r"(b)
                         : "memory");
            tma_store_commit();
          }
        }
      }
      if (!warp_has_rows) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_free[mt]);
      }
      if (st_on) stamp(mt, it, 5);
    }
    if (lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <bool kEmitProbs>
int launch_attn(const CUtensorMap& tm, const CUtensorMap& tm_out, float* probs, int items, cudaStream_t stream,
                long long* dbg) {
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(vit_attention_tc_kernel<kEmitProbs>), kSmemBytes,
                                   "vit_attention_tc"))
    return rc;
  const int grid = balanced_ctas(items, num_sms());  // 1,536 items at batch 256: 140 CTAs, eleven rounds
  return check_cuda(launch_pdl(vit_attention_tc_kernel<kEmitProbs>, dim3(grid), dim3(kThreads), size_t(kSmemBytes), stream, 1,
                               tm, tm_out, probs, items, g_tile_reverse ? 1 : 0, dbg),
                    "vit_attention_tc launch");
}

}  // namespace

int vit_attention_tc(const sais_bf16* qkv, int B, sais_bf16* out, float* probs, cudaStream_t stream) {
  if (B == 0) return kOk;
  CUtensorMap tm;
  int rc = make_tmap_bf16_3d(&tm, qkv, /*d0=*/1152, /*d1=*/T, /*d2=*/uint64_t(B), /*ld1=*/1152,
                             /*ld2=*/uint64_t(T) * 1152, /*box0=*/HD, /*box1=*/KEYS);
  if (rc) return rc;
  CUtensorMap tm_out;
  rc = make_tmap_bf16_3d(&tm_out, out, /*d0=*/384, /*d1=*/T, /*d2=*/uint64_t(B), /*ld1=*/384, /*ld2=*/uint64_t(T) * 384,
                         /*box0=*/HD, /*box1=*/32);
  if (rc) return rc;
  static const char* timeline = getenv("SAIS_ATTN_TIMELINE");
  long long* dbg = nullptr;
  if (timeline) {
    if (cudaMalloc(&dbg, 4 * 16 * 8 * sizeof(long long)) != cudaSuccess) dbg = nullptr;
    if (dbg) cudaMemsetAsync(dbg, 0, 4 * 16 * 8 * sizeof(long long), stream);
  }
  {
    LaunchScope ls(kClsVitAttn, stream, 4.0 * double(B) * 6 * 197 * 197 * 64);
    rc = probs ? launch_attn<true>(tm, tm_out, probs, B * HEADS, stream, dbg)
               : launch_attn<false>(tm, tm_out, nullptr, B * HEADS, stream, dbg);
  }
  if (dbg) {
    long long h[4 * 16 * 8];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(dbg);
    if (FILE* f = fopen(timeline, "w")) {
      for (int r = 0; r < 4; ++r)
        for (int i = 0; i < 16; ++i) {
          fprintf(f, "%d %d", r, i);
          for (int e = 0; e < 8; ++e) fprintf(f, " %lld", h[(r * 16 + i) * 8 + e]);
          fprintf(f, "\n");
        }
      fclose(f);
    }
  }
  return rc;
}

}  // namespace sais

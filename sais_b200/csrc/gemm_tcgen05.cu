// Persistent warp-specialised bf16 GEMM for sm_100a:  out = act(A · Wᵀ + bias) [+ residual].
//
//   warp 0      : TMA producer   (cp.async.bulk.tensor, 128-byte swizzle, kStages-deep smem ring)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (128 x BLOCK_N x 16 per instruction)
//   warps 2..5  : epilogue       (tcgen05.ld -> bias / GELU(erf) / ReLU / residual / pos-embed -> global)
//
// Accumulators live in TMEM and are double buffered (2 x BLOCK_N fp32 columns), so the epilogue of
// tile i overlaps the MMAs of tile i+1.  Tiles are walked n-fastest so that CTAs running concurrently
// share the same A rows through L2 while the (small) weight matrix stays L2 resident.
//
// Reference call sites this kernel replaces: nn.Linear / conv-as-GEMM in
// SAIS/scripts/dino-main/vision_transformer.py:60-63,82,90,126-130 and the in/out/FF projections of
// nn.TransformerEncoderLayer reached via SAIS/scripts/prepare_model.py:213.
#include "common.cuh"
#include "kernels.h"

namespace sais {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kGemmThreads = 192;
constexpr int kEpiThreads = 128;

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BLOCK_N == 256) ? 4 : (BLOCK_N == 192 ? 5 : 6);
  static constexpr int kTmemCols = (2 * BLOCK_N <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
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
  int split3;     // 1: A and W hold [hi | lo] bf16 halves (2K columns); accumulate hi*hi + lo*hi + hi*lo
  int split_out;  // 1: out_bf16 has 2N columns, value v is stored as hi = bf16(v) at n and lo = bf16(v - hi) at N + n
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

template <int BLOCK_N>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled UMMA/TMA tiles need 1024-byte aligned bases
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;                       // [kStages]
  uint64_t* empty_bar = bars + Cfg::kStages;       // [kStages]
  uint64_t* tfull_bar = bars + 2 * Cfg::kStages;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;            // [2]
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int n_tiles = p.N / BLOCK_N;
  const int num_tiles = m_tiles * n_tiles;
  const int kb_per_pass = p.K / BLOCK_K;
  const int k_blocks = p.split3 ? 3 * kb_per_pass : kb_per_pass;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], kEpiThreads);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_base_smem, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * BLOCK_M;
        const int n0 = (tile % n_tiles) * BLOCK_N;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          // split3 passes: (A_hi, W_hi), (A_lo, W_hi), (A_hi, W_lo); halves sit side by side along K
          int ka = kb, kw = kb;
          if (kb >= 2 * kb_per_pass) {
            ka = kb - 2 * kb_per_pass;
            kw = kb - kb_per_pass;
          } else if (kb >= kb_per_pass) {
            kw = kb - kb_per_pass;
          }
          tma_load_2d(sa, &tmap_a, &full_bar[stage], ka * BLOCK_K, m0);
          tma_load_2d(sb, &tmap_b, &full_bar[stage], kw * BLOCK_K, n0);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int astage = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[astage], aphase ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + astage * BLOCK_N;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kABytes;
          const uint64_t da = umma_desc_sw128_kmajor(sa);
          const uint64_t db = umma_desc_sw128_kmajor(sb);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advancing K inside the 128-byte swizzle row: +32 bytes = +2 in 16-byte address units
            umma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[astage]);  // accumulator complete -> epilogue
        if (++astage == 2) {
          astage = 0;
          aphase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int astage = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / n_tiles) * BLOCK_M;
      const int n0 = (tile % n_tiles) * BLOCK_N;
      mbar_wait(&tfull_bar[astage], aphase);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      int64_t orow = row;
      const float* radd = nullptr;
      if (p.remap_group > 0) {
        const int g = row / p.remap_group;
        const int pidx = row - g * p.remap_group;
        orow = int64_t(g) * (p.remap_group + 1) + 1 + pidx;
        radd = p.row_add + int64_t(pidx) * p.N;
      }
      const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + astage * BLOCK_N;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(t_row + c, v);
        tmem_ld_wait();
        if (row_ok) {
          const int n = n0 + c;
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n + j));
              f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
            }
          }
          if (p.act == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
          } else if (p.act == 2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
          }
          if (radd != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 a4 = __ldg(reinterpret_cast<const float4*>(radd + n + j));
              f[j] += a4.x; f[j + 1] += a4.y; f[j + 2] += a4.z; f[j + 3] += a4.w;
            }
          }
          if (p.residual != nullptr) {
            const float* rp = p.residual + orow * p.ldr + n;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 r4 = *reinterpret_cast<const float4*>(rp + j);
              f[j] += r4.x; f[j + 1] += r4.y; f[j + 2] += r4.z; f[j + 3] += r4.w;
            }
          }
          if (p.out_f32 != nullptr) {
            float* op = p.out_f32 + orow * p.ldo32 + n;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(op + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
          }
          if (p.out_bf16 != nullptr) {
            __nv_bfloat16* op = p.out_bf16 + orow * p.ldo16 + n;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 o;
              o.x = pack_bf16x2(f[j], f[j + 1]);
              o.y = pack_bf16x2(f[j + 2], f[j + 3]);
              o.z = pack_bf16x2(f[j + 4], f[j + 5]);
              o.w = pack_bf16x2(f[j + 6], f[j + 7]);
              *reinterpret_cast<uint4*>(op + j) = o;
              if (p.split_out) {  // residual halves for the split-precision consumer
                uint4 l;
                l.x = pack_bf16x2(f[j] - bf16_lo(o.x), f[j + 1] - bf16_hi(o.x));
                l.y = pack_bf16x2(f[j + 2] - bf16_lo(o.y), f[j + 3] - bf16_hi(o.y));
                l.z = pack_bf16x2(f[j + 4] - bf16_lo(o.z), f[j + 5] - bf16_hi(o.z));
                l.w = pack_bf16x2(f[j + 6] - bf16_lo(o.w), f[j + 7] - bf16_hi(o.w));
                *reinterpret_cast<uint4*>(op + p.N + j) = l;
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[astage]);
      if (++astage == 2) {
        astage = 0;
        aphase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int BLOCK_N>
int launch_gemm(const SaisGemmArgs& a, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N>;
  CUtensorMap ta, tb;
  const uint64_t kcols = uint64_t(a.K) * (a.split3 ? 2 : 1);
  int rc = make_tmap_bf16_2d(&ta, a.a, uint64_t(a.M), kcols, uint64_t(a.lda), BLOCK_M, BLOCK_K);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tb, a.w, uint64_t(a.N), kcols, uint64_t(a.ldw), BLOCK_N, BLOCK_K);
  if (rc) return rc;

  static bool attr_set = false;
  if (!attr_set) {
    rc = check_cuda(cudaFuncSetAttribute(gemm_tcgen05_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes),
                    "cudaFuncSetAttribute(gemm)");
    if (rc) return rc;
    attr_set = true;
  }
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
  const int m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int tiles = m_tiles * (p.N / BLOCK_N);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  LaunchScope ls(kClsGemm, stream, 2.0 * double(a.M) * double(a.N) * double(a.K) * (a.split3 ? 3 : 1));
  gemm_tcgen05_kernel<BLOCK_N><<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(ta, tb, p);
  return check_cuda(cudaGetLastError(), "gemm_tcgen05_kernel launch");
}

}  // namespace

int pick_block_n(int64_t M, int64_t N) {
  // Prefer the widest tile that divides N while still giving every SM work; wide tiles halve the
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
  if (!a.a || !a.w || (!a.out_f32 && !a.out_bf16)) {
    set_last_error("gemm: null operand/output pointer");
    return kErrInvalidArg;
  }
  if (a.M <= 0 || a.N <= 0 || a.K <= 0 || a.K % BLOCK_K != 0 || a.N % 128 != 0 && a.N % 192 != 0) {
    set_last_error("gemm: unsupported shape M=%lld N=%lld K=%lld (need K%%64==0, N%%128==0 or N%%192==0)",
                   (long long)a.M, (long long)a.N, (long long)a.K);
    return kErrShape;
  }
  if (a.lda % 8 || a.ldw % 8 || (a.out_bf16 && a.ldo16 % 8) || (a.out_f32 && a.ldo32 % 4) ||
      (a.residual && a.ldr % 4)) {
    set_last_error("gemm: row pitches must keep 16-byte alignment");
    return kErrInvalidArg;
  }
  if ((reinterpret_cast<uintptr_t>(a.a) | reinterpret_cast<uintptr_t>(a.w)) & 15) {
    set_last_error("gemm: A/W must be 16-byte aligned");
    return kErrInvalidArg;
  }
  if (a.remap_group > 0 && a.row_add == nullptr) {
    set_last_error("gemm: remap_group needs row_add");
    return kErrInvalidArg;
  }
  if (a.act < 0 || a.act > 2) {
    set_last_error("gemm: bad activation %d", a.act);
    return kErrInvalidArg;
  }
  const int bn = force_block_n ? force_block_n : pick_block_n(a.M, a.N);
  if (bn == 0 || a.N % bn) {
    set_last_error("gemm: N=%lld not divisible by tile %d", (long long)a.N, bn);
    return kErrShape;
  }
  switch (bn) {
    case 256: return launch_gemm<256>(a, stream);
    case 192: return launch_gemm<192>(a, stream);
    case 128: return launch_gemm<128>(a, stream);
  }
  set_last_error("gemm: bad tile %d", bn);
  return kErrShape;
}

}  // namespace sais

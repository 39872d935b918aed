// Exact (fp32) multi-head attention over packed short sequences.
//
// Primary use: the SAIS temporal encoder's nn.MultiheadAttention core (4 heads x 96, reached through
// prepare_model.py:213 with the README-patched layer that also returns the head-averaged weights).
// Sequences are short (S = nframes + 1, typically 10..65) and packed back to back, so this is a
// latency/L2-bound SIMT kernel rather than a tensor-core one: one CTA per sequence keeps that sequence's K and
// V (all heads) in shared memory; the CTA's warps form an (H heads) x (RG row groups) grid, so RG query rows are in
// flight with all their heads at once (fp32 logits, -inf on padded keys, exact softmax).  The per-head probabilities
// of a row batch stay in shared memory and are averaged in a fixed head order (deterministic, no atomics), so
// the head-mean map [S,S] is written once, coalesced.  Second use: the fp32-equivalent ("precise") mode of the ViT
// (6 heads x 64, S = 197, vision_transformer.py:80-92), which optionally emits per-head probabilities.
//
// q, k, v arrive as fp32 (the split-precision in-proj GEMM writes fp32); the output is written as bf16
// [hi | lo] halves for the split-precision out-proj GEMM.
#include "common.cuh"
#include "kernels.h"

namespace sais {

namespace {


__host__ __device__ inline int round_up32(int s) { return (s + 31) & ~31; }

template <int H, int HD>
struct AttnCfg {
  static constexpr int E = H * HD;
  static constexpr int LD = 3 * E;        // qkv row pitch (fp32 elements)
  static constexpr int KV_PITCH = E + 4;  // smem row pitch: conflict-free float4 row reads
  static constexpr int kMaxSmemS = 64;
  static constexpr int RG = (H == 4) ? 8 : 3;  // row groups: 32 warps for the temporal head (short sequences: fewer rounds), 18 for the ViT (6 heads)
  static constexpr int kWarps = H * RG;
  static constexpr int kThreads = 32 * kWarps;
  static size_t smem_bytes(int max_S, bool smem_kv) {
    size_t b = size_t(kWarps) * (HD * 4 + size_t(round_up32(max_S)) * 4);
    if (smem_kv) b += 2 * size_t(max_S) * KV_PITCH * 4;
    return b;
  }
};

template <int H, int HD, bool kSmemKV>
__global__ void __launch_bounds__(32 * H * ((H == 4) ? 8 : 3))
seq_attention_f32_kernel(const float* __restrict__ qkv, const int32_t* __restrict__ seq_offsets,
                         const uint8_t* __restrict__ key_pad, const int64_t* __restrict__ attn_offsets, int max_S,
                         float scale, __nv_bfloat16* __restrict__ out_split, float* __restrict__ attn_mean,
                         float* __restrict__ probs_per_head) {
  pdl_wait();  // (PDL, common.cuh) no global access above this line
  using Cfg = AttnCfg<H, HD>;
  constexpr int E = Cfg::E, LD = Cfg::LD;
  constexpr int kWarps = Cfg::kWarps, kThreads = Cfg::kThreads, RG = Cfg::RG;
  constexpr int DPL = HD / 32;  // output dims per lane
  extern __shared__ __align__(16) uint8_t smem[];
  const int i = blockIdx.x;
  const int t0 = seq_offsets[i];
  const int S = seq_offsets[i + 1] - t0;
  if (S <= 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Sp = round_up32(max_S);

  float* qs = reinterpret_cast<float*>(smem) + warp * HD;                    // [kWarps][HD]
  float* sc_all = reinterpret_cast<float*>(smem) + kWarps * HD;              // [kWarps][Sp]
  float* sc = sc_all + warp * Sp;

  const float* Kb;
  const float* Vb;
  int pitch;
  if constexpr (kSmemKV) {
    float* Ks = reinterpret_cast<float*>(smem + size_t(kWarps) * (HD * 4 + size_t(Sp) * 4));
    float* Vs = Ks + size_t(max_S) * Cfg::KV_PITCH;
    constexpr int CH = 2 * E / 4;  // float4 chunks per token (K then V)
    for (int e = threadIdx.x; e < S * CH; e += kThreads) {
      const int j = e / CH, c = e % CH;
      const float4 v = __ldg(reinterpret_cast<const float4*>(qkv + int64_t(t0 + j) * LD + E) + c);
      float* dst = (c < E / 4) ? (Ks + j * Cfg::KV_PITCH + c * 4) : (Vs + j * Cfg::KV_PITCH + (c - E / 4) * 4);
      *reinterpret_cast<float4*>(dst) = v;
    }
    __syncthreads();
    Kb = Ks; Vb = Vs; pitch = Cfg::KV_PITCH;
  } else {
    Kb = qkv + int64_t(t0) * LD + E;
    Vb = qkv + int64_t(t0) * LD + 2 * E;
    pitch = LD;
  }

  const int64_t aoff = (attn_mean != nullptr && attn_offsets != nullptr) ? attn_offsets[i] : -1;
  const uint8_t* pad = key_pad ? key_pad + t0 : nullptr;

  const int h = warp % H;    // this warp's head
  const int rg = warp / H;   // and row group
  for (int r0 = 0; r0 < S; r0 += RG) {
    const int r = r0 + rg;
    if (r < S) {
      const float* qp = qkv + int64_t(t0 + r) * LD + h * HD;
#pragma unroll
      for (int c = 0; c < DPL; ++c) qs[lane + 32 * c] = qp[lane + 32 * c];
      __syncwarp();
      float lmax = -INFINITY;
      for (int j = lane; j < S; j += 32) {
        const float4* kp = reinterpret_cast<const float4*>(Kb + int64_t(j) * pitch + h * HD);
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < HD / 4; ++c) {
          const float4 kk = kp[c];
          const float4 qq = *reinterpret_cast<const float4*>(qs + c * 4);
          dot = fmaf(kk.x, qq.x, dot); dot = fmaf(kk.y, qq.y, dot);
          dot = fmaf(kk.z, qq.z, dot); dot = fmaf(kk.w, qq.w, dot);
        }
        float sv = dot * scale;
        if (pad && pad[j]) sv = -INFINITY;
        sc[j] = sv;
        lmax = fmaxf(lmax, sv);
      }
      const float m = warp_max(lmax);
      float lsum = 0.f;
      for (int j = lane; j < S; j += 32) {
        const float pe = expf(sc[j] - m);
        sc[j] = pe;
        lsum += pe;
      }
      const float inv = 1.0f / warp_sum(lsum);
      float* ph = probs_per_head ? probs_per_head + ((int64_t(i) * H + h) * S + r) * S : nullptr;
      for (int j = lane; j < S; j += 32) {
        const float pn = sc[j] * inv;
        sc[j] = pn;
        if (ph) ph[j] = pn;
      }
      __syncwarp();
      // O = P V for this head: lane owns dims lane, lane+32, ...
      float o[DPL];
#pragma unroll
      for (int c = 0; c < DPL; ++c) o[c] = 0.f;
      const float* vp = Vb + h * HD + lane;
      for (int j = 0; j < S; ++j) {
        const float pj = sc[j];
        const float* v = vp + int64_t(j) * pitch;
#pragma unroll
        for (int c = 0; c < DPL; ++c) o[c] = fmaf(pj, v[32 * c], o[c]);
      }
      __nv_bfloat16* op = out_split + int64_t(t0 + r) * (2 * E) + h * HD + lane;
#pragma unroll
      for (int c = 0; c < DPL; ++c) {
        const __nv_bfloat16 hi = __float2bfloat16(o[c]);
        op[32 * c] = hi;
        op[E + 32 * c] = __float2bfloat16(o[c] - __bfloat162float(hi));
      }
    }
    if (aoff >= 0) {  // head mean of this row batch, heads summed in a fixed order (block-uniform branch)
      __syncthreads();
      for (int e = threadIdx.x; e < RG * S; e += kThreads) {
        const int g = e / S, j = e % S;
        if (r0 + g < S) {
          float a = 0.f;
#pragma unroll
          for (int hh = 0; hh < H; ++hh) a += sc_all[(g * H + hh) * Sp + j];
          attn_mean[aoff + int64_t(r0 + g) * S + j] = a * (1.0f / H);
        }
      }
      __syncthreads();
    } else {
      __syncwarp();
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Tensor-core variant for the temporal head (4 heads x 96, S <= 80): warp-level TF32 MMAs on [hi | lo] operand splits
// ("3xTF32": hi*hi + lo*hi + hi*lo with hi = tf32(x), lo = x - hi; the dropped lo*lo term and the representation error are
// both 2^-22 relative, i.e. fp32-grade like the scalar kernel above — a bf16 split would be 2^-17, visible at the 2e-5
// tolerance this kernel is tested to).
//
// Why warp-level mma.sync and not tcgen05 here: one (sequence, head) problem is S x S x 96 with S = 11..65 rows, a small
// fraction of the 128-row UMMA tile.  Packing consecutive sequences into one 128-row tile makes the logits block-diagonal
// (3/4 of the MMA work wasted at S = 31, on top of the 3x of the split) and needs a staging pass that converts the fp32
// q|k|v rows to operand tiles in shared memory first; the 16-row warp tile pads S = 31 to 32 and takes its operand
// fragments straight from the fp32 rows in L2.  The work is a few GFLOP per launch either way; what the scalar kernel
// above lost was shared-memory bandwidth (one LDS.128 per 4 FMA).
//
// CTA = (16-row block, sequence); warp = head; g = lane / 4, t = lane % 4 (the m16n8k8 fragment coordinates).
// The summation index of an MMA, and the column index of its B / D operands, can be permuted freely as long as both
// operands (resp. the consumer of D) agree — used three times so that EVERY global access is a contiguous 16-byte load or
// store and the code is branch-free (all loads of a phase issue back to back; out-of-range keys are predicated to zero):
//   S = Q Kt  : k slot (step ks, k = t) stands for dim 24t + 2ks and slot (ks, k = t + 4) for dim 24t + 2ks + 1: a lane owns
//               dims [24t, 24t + 24) of its Q rows (r0 + g, r0 + g + 8) and of key nt * 8 + g — six float4 loads per row,
//               the four lanes of a quad read one whole 384-byte head row
//   softmax   : exact, in the accumulator layout (row statistics across the 4 lanes of a quad); padded / out-of-range keys
//               get -inf and come out as exactly 0
//   head mean : the four warps park their normalised rows in shared memory and sum them in a fixed head order
//   O = P V   : the probability accumulators ARE the A fragments of the second MMA: k slot t stands for key 2t, slot t + 4
//               for key 2t + 1 of the 8-key tile — exactly the two columns a lane's accumulators hold; B column g of dim
//               tile dt stands for dim 12g + dt, so a lane reads 12 contiguous floats of V rows 2t and 2t + 1, and its
//               output columns (2t, 2t + 1) of the 12 dim tiles are the contiguous dims [24t, 24t + 24) of rows g, g + 8
// Output: bf16 [hi | lo] halves for the split-precision out-proj GEMM, as above.
__device__ __forceinline__ void mma_tf32_1688(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  lo = __float_as_uint(x - __uint_as_float(hi));  // exact; the MMA reads its upper 19 bits
}

template <int MAXNT>
__global__ void __launch_bounds__(128, MAXNT <= 4 ? 4 : 2)
seq_attention_mma_kernel(const float* __restrict__ qkv, const int32_t* __restrict__ seq_offsets,
                         const uint8_t* __restrict__ key_pad, const int64_t* __restrict__ attn_offsets, float scale,
                         __nv_bfloat16* __restrict__ out_split, float* __restrict__ attn_mean) {
  constexpr int H = 4, HD = 96, E = H * HD, LD = 3 * E;
  constexpr int SP = MAXNT * 8 + 4;  // smem row pitch of the parked probabilities
  __shared__ float probs_s[H][16][SP];
  pdl_wait();  // (PDL, common.cuh) no global access above this line
  const int i = blockIdx.y;
  const int t0 = seq_offsets[i];
  const int S = seq_offsets[i + 1] - t0;
  const int r0 = blockIdx.x * 16;
  if (r0 >= S) return;
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const float* base = qkv + int64_t(t0) * LD + h * HD;

  // ---- S = Q Kt
  float sacc[MAXNT][4];
#pragma unroll
  for (int nt = 0; nt < MAXNT; ++nt) sacc[nt][0] = sacc[nt][1] = sacc[nt][2] = sacc[nt][3] = 0.f;
  const bool qa = (r0 + g) < S, qb = (r0 + g + 8) < S;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 qra[6], qrb[6];  // dims [24t, 24t + 24) of rows r0 + g, r0 + g + 8
  {
    const float4* qpa = reinterpret_cast<const float4*>(base + int64_t(r0 + g) * LD + 24 * t);
    const float4* qpb = reinterpret_cast<const float4*>(base + int64_t(r0 + g + 8) * LD + 24 * t);
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      qra[c] = qa ? __ldg(qpa + c) : zero4;
      qrb[c] = qb ? __ldg(qpb + c) : zero4;
    }
  }
#pragma unroll
  for (int nt = 0; nt < MAXNT; ++nt) {
    const int key = nt * 8 + g;
    const bool kv = key < S;
    const float4* kp = reinterpret_cast<const float4*>(base + E + int64_t(key) * LD + 24 * t);
    float4 kr[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) kr[c] = kv ? __ldg(kp + c) : zero4;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {  // k-step ks = 2c + hh covers dims 24t + 4c + 2hh, + 1
        uint32_t ah[4], al[4], bh0, bl0, bh1, bl1;
        split_tf32(hh ? qra[c].z : qra[c].x, ah[0], al[0]);
        split_tf32(hh ? qrb[c].z : qrb[c].x, ah[1], al[1]);
        split_tf32(hh ? qra[c].w : qra[c].y, ah[2], al[2]);
        split_tf32(hh ? qrb[c].w : qrb[c].y, ah[3], al[3]);
        split_tf32(hh ? kr[c].z : kr[c].x, bh0, bl0);
        split_tf32(hh ? kr[c].w : kr[c].y, bh1, bl1);
        mma_tf32_1688(sacc[nt], al, bh0, bh1);
        mma_tf32_1688(sacc[nt], ah, bl0, bl1);
        mma_tf32_1688(sacc[nt], ah, bh0, bh1);
      }
    }
  }

  // ---- exact softmax over the keys of this sequence (rows g and g + 8 of the block; columns nt * 8 + 2t, + 1)
  const uint8_t* pad = key_pad ? key_pad + t0 : nullptr;
  float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < MAXNT; ++nt) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int key = nt * 8 + 2 * t + c;
      const bool ok = key < S && !(pad && pad[key]);
      sacc[nt][c] = ok ? sacc[nt][c] * scale : -INFINITY;
      sacc[nt][2 + c] = ok ? sacc[nt][2 + c] * scale : -INFINITY;
      ma = fmaxf(ma, sacc[nt][c]);
      mb = fmaxf(mb, sacc[nt][2 + c]);
    }
  }
  ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 1));
  ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 2));
  mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 1));
  mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 2));
  float sa = 0.f, sb = 0.f;
#pragma unroll
  for (int nt = 0; nt < MAXNT; ++nt) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const float pa = expf(sacc[nt][c] - ma), pb = expf(sacc[nt][2 + c] - mb);
      sacc[nt][c] = pa;
      sacc[nt][2 + c] = pb;
      sa += pa;
      sb += pb;
    }
  }
  sa += __shfl_xor_sync(0xffffffffu, sa, 1);
  sa += __shfl_xor_sync(0xffffffffu, sa, 2);
  sb += __shfl_xor_sync(0xffffffffu, sb, 1);
  sb += __shfl_xor_sync(0xffffffffu, sb, 2);
  const float ia = 1.0f / sa, ib = 1.0f / sb;
#pragma unroll
  for (int nt = 0; nt < MAXNT; ++nt) {
    sacc[nt][0] *= ia; sacc[nt][1] *= ia;
    sacc[nt][2] *= ib; sacc[nt][3] *= ib;
  }

  // ---- head-averaged map of this row block (heads summed in a fixed order: deterministic)
  const int64_t aoff = (attn_mean != nullptr && attn_offsets != nullptr) ? attn_offsets[i] : -1;  // block-uniform
  if (aoff >= 0) {
#pragma unroll
    for (int nt = 0; nt < MAXNT; ++nt) {
      *reinterpret_cast<float2*>(&probs_s[h][g][nt * 8 + 2 * t]) = make_float2(sacc[nt][0], sacc[nt][1]);
      *reinterpret_cast<float2*>(&probs_s[h][g + 8][nt * 8 + 2 * t]) = make_float2(sacc[nt][2], sacc[nt][3]);
    }
    __syncthreads();
    const int rows = min(16, S - r0);
    for (int e = threadIdx.x; e < rows * S; e += 128) {
      const int r = e / S, j = e - r * S;
      const float a = ((probs_s[0][r][j] + probs_s[1][r][j]) + probs_s[2][r][j]) + probs_s[3][r][j];
      attn_mean[aoff + int64_t(r0 + r) * S + j] = a * (1.0f / H);
    }
  }

  // ---- O = P V  (k slot t <-> key 2t, slot t + 4 <-> key 2t + 1 of the tile; B / D column g of tile dt <-> dim 12g + dt)
  float oacc[HD / 8][4];
#pragma unroll
  for (int dt = 0; dt < HD / 8; ++dt) oacc[dt][0] = oacc[dt][1] = oacc[dt][2] = oacc[dt][3] = 0.f;
#pragma unroll
  for (int nt = 0; nt < MAXNT; ++nt) {
    uint32_t ph[4], pl[4];
    split_tf32(sacc[nt][0], ph[0], pl[0]);  // (row g,     key 2t)
    split_tf32(sacc[nt][2], ph[1], pl[1]);  // (row g + 8, key 2t)
    split_tf32(sacc[nt][1], ph[2], pl[2]);  // (row g,     key 2t + 1)
    split_tf32(sacc[nt][3], ph[3], pl[3]);  // (row g + 8, key 2t + 1)
    const int k0 = nt * 8 + 2 * t;
    const float4* v0 = reinterpret_cast<const float4*>(base + 2 * E + int64_t(k0) * LD + 12 * g);
    const float4* v1 = reinterpret_cast<const float4*>(base + 2 * E + int64_t(k0 + 1) * LD + 12 * g);
    const bool e0 = k0 < S, e1 = k0 + 1 < S;
    float4 va[3], vb[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      va[c] = e0 ? __ldg(v0 + c) : zero4;
      vb[c] = e1 ? __ldg(v1 + c) : zero4;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float xa[4] = {va[c].x, va[c].y, va[c].z, va[c].w};
      const float xb[4] = {vb[c].x, vb[c].y, vb[c].z, vb[c].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t bh0, bl0, bh1, bl1;
        split_tf32(xa[j], bh0, bl0);
        split_tf32(xb[j], bh1, bl1);
        mma_tf32_1688(oacc[4 * c + j], pl, bh0, bh1);
        mma_tf32_1688(oacc[4 * c + j], ph, bl0, bl1);
        mma_tf32_1688(oacc[4 * c + j], ph, bh0, bh1);
      }
    }
  }
  // D columns (2t, 2t + 1) of dim tile dt are dims 24t + dt and 24t + 12 + dt: rows g / g + 8, dims [24t, 24t + 24) contiguous
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    if (rr ? qb : qa) {
      __nv_bfloat16* op = out_split + int64_t(t0 + r0 + g + 8 * rr) * (2 * E) + h * HD + 24 * t;
      uint32_t hi[12], lo[12];
#pragma unroll
      for (int j = 0; j < 12; ++j) {  // dims 24t + 2j, + 1
        const int d0 = 2 * j, d1 = 2 * j + 1;
        const float x0 = oacc[d0 % 12][2 * rr + d0 / 12], x1 = oacc[d1 % 12][2 * rr + d1 / 12];
        hi[j] = pack_bf16x2(x0, x1);
        lo[j] = pack_bf16x2(x0 - bf16_lo(hi[j]), x1 - bf16_hi(hi[j]));
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        *reinterpret_cast<uint4*>(op + 8 * c) = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
        *reinterpret_cast<uint4*>(op + E + 8 * c) = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
      }
    }
  }
}

template <int MAXNT>
int launch_seq_attention_mma(const float* qkv, const int32_t* seq_offsets, const uint8_t* key_pad,
                             const int64_t* attn_offsets, int nseq, int max_S, float scale, sais_bf16* out_split,
                             float* attn_mean, cudaStream_t stream) {
  LaunchScope ls(kClsTemporalAttn, stream, 0.0);
  return check_cuda(launch_pdl(seq_attention_mma_kernel<MAXNT>, dim3((max_S + 15) / 16, nseq), dim3(128), size_t(0), stream,
                               1, qkv, seq_offsets, key_pad, attn_offsets, scale,
                               reinterpret_cast<__nv_bfloat16*>(out_split), attn_mean),
                    "seq_attention_mma launch");
}

template <int H, int HD>
int launch_seq_attention(const float* qkv, const int32_t* seq_offsets, const uint8_t* key_pad,
                         const int64_t* attn_offsets, int nseq, int max_S, float scale, sais_bf16* out_split,
                         float* attn_mean, float* probs_per_head, int cls, cudaStream_t stream) {
  using Cfg = AttnCfg<H, HD>;
  if (nseq == 0) return kOk;
  if (!qkv || !seq_offsets || !out_split || nseq < 0 || max_S <= 0) {
    set_last_error("seq_attention: bad arguments");
    return kErrInvalidArg;
  }
  const bool smem_kv = max_S <= Cfg::kMaxSmemS;
  const size_t sm = Cfg::smem_bytes(max_S, smem_kv);
  if (sm > 227 * 1024) {
    set_last_error("seq_attention: max_S=%d needs %zu bytes of shared memory", max_S, sm);
    return kErrShape;
  }
  const void* fn = smem_kv ? reinterpret_cast<const void*>(seq_attention_f32_kernel<H, HD, true>)
                           : reinterpret_cast<const void*>(seq_attention_f32_kernel<H, HD, false>);
  if (int rc = ensure_dynamic_smem(fn, sm > 48 * 1024 ? 227 * 1024 : 48 * 1024, "seq_attention")) return rc;
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out_split);
  LaunchScope ls(cls, stream, 0.0);
  auto kern = smem_kv ? seq_attention_f32_kernel<H, HD, true> : seq_attention_f32_kernel<H, HD, false>;
  return check_cuda(launch_pdl(kern, dim3(nseq), dim3(Cfg::kThreads), size_t(sm), stream, 1, qkv, seq_offsets, key_pad,
                               attn_offsets, max_S, scale, o, attn_mean, probs_per_head),
                    "seq_attention launch");
}

}  // namespace

int temporal_attention(const float* qkv, const int32_t* seq_offsets, const uint8_t* key_pad,
                       const int64_t* attn_offsets, int nseq, int max_S, sais_bf16* out_split, float* attn_out,
                       cudaStream_t stream) {
  constexpr float scale = 0.10206207261596575f;  // 96^-0.5
  if (max_S <= 80 && nseq <= 65535) {  // the temporal head's range (clips of 10..65 tokens): tensor-core kernel
    if (nseq == 0) return kOk;
    if (!qkv || !seq_offsets || !out_split || nseq < 0 || max_S <= 0) {
      set_last_error("seq_attention: bad arguments");
      return kErrInvalidArg;
    }
    if (max_S <= 32)
      return launch_seq_attention_mma<4>(qkv, seq_offsets, key_pad, attn_offsets, nseq, max_S, scale, out_split, attn_out, stream);
    if (max_S <= 48)
      return launch_seq_attention_mma<6>(qkv, seq_offsets, key_pad, attn_offsets, nseq, max_S, scale, out_split, attn_out, stream);
    return launch_seq_attention_mma<10>(qkv, seq_offsets, key_pad, attn_offsets, nseq, max_S, scale, out_split, attn_out, stream);
  }
  return launch_seq_attention<4, 96>(qkv, seq_offsets, key_pad, attn_offsets, nseq, max_S, scale, out_split, attn_out,
                                     nullptr, kClsTemporalAttn, stream);
}

int vit_attention_precise(const float* qkv, const int32_t* seq_offsets, int B, sais_bf16* out_split, float* probs,
                          cudaStream_t stream) {
  return launch_seq_attention<6, 64>(qkv, seq_offsets, nullptr, nullptr, B, 197, 0.125f, out_split, nullptr, probs,
                                     kClsVitAttn, stream);
}

}  // namespace sais

// Fused ViT self-attention for T = 197 tokens, 6 heads x 64 (vision_transformer.py:80-92).
// One CTA per (frame, head): Q, K, V of the head are staged once in shared memory (XOR-swizzled,
// cp.async), each warp owns 16-query-row tiles and runs QKᵀ -> online softmax -> PV over two
// 112-key chunks entirely in registers (warp-level bf16 MMA, fp32 accumulate); the probabilities never
// touch HBM unless the caller asks for them (get_last_selfattention, :216-223), in which case a second
// pass recomputes the logits and writes exp(s - max)/sum as fp32 [B,6,197,197].
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace sais {

namespace {

constexpr int T = 197;
constexpr int HD = 64;
constexpr int HEADS = 6;
constexpr int QKV_LD = 1152;
constexpr int OUT_LD = 384;
constexpr int Q_ROWS = 208;   // 13 tiles of 16 query rows
constexpr int K_ROWS = 224;   // 2 chunks of 112 keys
constexpr int CHUNK = 112;
constexpr int NT = CHUNK / 8;  // 14 n-tiles per chunk
constexpr int kThreads = 256;
constexpr int kSmem = (Q_ROWS + 2 * K_ROWS) * HD * 2;

// byte offset of (row, 16-byte chunk) inside a [rows][64] bf16 tile with the 8-row XOR swizzle
__device__ __forceinline__ uint32_t swz(int row, int chunk) { return uint32_t(row * 128 + ((chunk ^ (row & 7)) << 4)); }

__device__ __forceinline__ void qk_chunk(float (&s)[NT][4], const uint32_t (&qf)[4][4], uint32_t ks_base, int key0,
                                         int lane) {
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
  const int mi = lane >> 3;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int np = 0; np < NT / 2; ++np) {
      const int key = key0 + np * 16 + (mi >> 1) * 8 + (lane & 7);
      uint32_t b0, b1, b2, b3;
      ldmatrix_x4(b0, b1, b2, b3, ks_base + swz(key, 2 * ks + (mi & 1)));
      mma_bf16_16816(s[2 * np], qf[ks], b0, b1);
      mma_bf16_16816(s[2 * np + 1], qf[ks], b2, b3);
    }
  }
}

template <bool kEmitProbs>
__global__ void __launch_bounds__(kThreads, 2)
vit_attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                     float* __restrict__ probs) {
  pdl_trigger();
  pdl_wait();  // (PDL, common.cuh) no global access above this line
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t qs = smem_u32(smem);
  const uint32_t ks = qs + Q_ROWS * 128;
  const uint32_t vs = ks + K_ROWS * 128;
  const int b = blockIdx.x / HEADS, h = blockIdx.x % HEADS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const __nv_bfloat16* base = qkv + int64_t(b) * T * QKV_LD + h * HD;

  // ---- stage Q, K, V (zero fill beyond row 196) ----
  for (int e = threadIdx.x; e < (Q_ROWS + 2 * K_ROWS) * 8; e += kThreads) {
    const int chunk = e & 7;
    int row = e >> 3;
    uint32_t dst;
    int col;
    if (row < Q_ROWS) {
      dst = qs; col = 0;
    } else if (row < Q_ROWS + K_ROWS) {
      row -= Q_ROWS; dst = ks; col = 384;
    } else {
      row -= Q_ROWS + K_ROWS; dst = vs; col = 768;
    }
    const bool ok = row < T;
    const __nv_bfloat16* src = base + int64_t(ok ? row : 0) * QKV_LD + col + chunk * 8;
    cp_async_16(dst + swz(row, chunk), src, ok);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  const float sl2 = 0.125f * 1.4426950408889634f;  // head_dim^-0.5 * log2(e)
  const int r0 = lane >> 2;                        // row within the 16-row tile (and r0 + 8)
  const int cq = (lane & 3) * 2;                   // column pair within an 8-wide n-tile

  for (int mt = warp; mt < Q_ROWS / 16; mt += kThreads / 32) {
    // Q fragments for the 4 k-steps of d = 64
    uint32_t qf[4][4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
      ldmatrix_x4(qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3], qs + swz(mt * 16 + (lane & 15), 2 * kk + (lane >> 4)));

    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

#pragma unroll 1
    for (int ch = 0; ch < 2; ++ch) {
      float s[NT][4];
      qk_chunk(s, qf, ks, ch * CHUNK, lane);
      if (ch == 1) {  // keys >= 197 are padding
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int key = CHUNK + nt * 8 + cq;
          if (key >= T) s[nt][0] = s[nt][2] = -INFINITY;
          if (key + 1 >= T) s[nt][1] = s[nt][3] = -INFINITY;
        }
      }
      float cm0 = -INFINITY, cm1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        cm0 = fmaxf(cm0, fmaxf(s[nt][0], s[nt][1]));
        cm1 = fmaxf(cm1, fmaxf(s[nt][2], s[nt][3]));
      }
      cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 1));
      cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 2));
      cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 1));
      cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 2));
      const float nm0 = fmaxf(m0, cm0), nm1 = fmaxf(m1, cm1);
      const float a0 = exp2f((m0 - nm0) * sl2), a1 = exp2f((m1 - nm1) * sl2);
      m0 = nm0; m1 = nm1;
      l0 *= a0; l1 *= a1;
#pragma unroll
      for (int i = 0; i < 8; ++i) { o[i][0] *= a0; o[i][1] *= a0; o[i][2] *= a1; o[i][3] *= a1; }
      const float mb0 = m0 * sl2, mb1 = m1 * sl2;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        s[nt][0] = exp2f(fmaf(s[nt][0], sl2, -mb0));
        s[nt][1] = exp2f(fmaf(s[nt][1], sl2, -mb0));
        s[nt][2] = exp2f(fmaf(s[nt][2], sl2, -mb1));
        s[nt][3] = exp2f(fmaf(s[nt][3], sl2, -mb1));
        l0 += s[nt][0] + s[nt][1];
        l1 += s[nt][2] + s[nt][3];
      }
      // O += P V
      const int mi = lane >> 3;
#pragma unroll
      for (int kk = 0; kk < NT / 2; ++kk) {
        uint32_t pa[4];
        pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
        pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
        pa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        pa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
        const int key = ch * CHUNK + kk * 16 + (mi & 1) * 8 + (lane & 7);
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
          uint32_t b0, b1, b2, b3;
          ldmatrix_x4_trans(b0, b1, b2, b3, vs + swz(key, 2 * dp + (mi >> 1)));
          mma_bf16_16816(o[2 * dp], pa, b0, b1);
          mma_bf16_16816(o[2 * dp + 1], pa, b2, b3);
        }
      }
    }
    // full row sums across the 4 lanes that share a row
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float il0 = 1.0f / l0, il1 = 1.0f / l1;

    // ---- write O through this warp's own (now dead) Q rows for 16-byte coalesced stores ----
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t v0 = pack_bf16x2(o[i][0] * il0, o[i][1] * il0);
      const uint32_t v1 = pack_bf16x2(o[i][2] * il1, o[i][3] * il1);
      const int row_a = mt * 16 + r0, row_b = row_a + 8;
      *reinterpret_cast<uint32_t*>(smem + swz(row_a, i) + cq * 2) = v0;
      *reinterpret_cast<uint32_t*>(smem + swz(row_b, i) + cq * 2) = v1;
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = i * 32 + lane;
      const int row = mt * 16 + (e >> 3), chunk = e & 7;
      if (row < T) {
        const uint4 v = *reinterpret_cast<const uint4*>(smem + swz(row, chunk));
        *reinterpret_cast<uint4*>(out + (int64_t(b) * T + row) * OUT_LD + h * HD + chunk * 8) = v;
      }
    }

    if constexpr (kEmitProbs) {
      // second pass: recompute logits, write normalised probabilities (fp32)
      const float mb0 = m0 * sl2, mb1 = m1 * sl2;
      const int row_a = mt * 16 + r0, row_b = row_a + 8;
      float* pa_out = probs + ((int64_t(b) * HEADS + h) * T + row_a) * T;
      float* pb_out = probs + ((int64_t(b) * HEADS + h) * T + row_b) * T;
#pragma unroll 1
      for (int ch = 0; ch < 2; ++ch) {
        float s[NT][4];
        qk_chunk(s, qf, ks, ch * CHUNK, lane);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int key = ch * CHUNK + nt * 8 + cq;
          if (row_a < T) {
            if (key < T) pa_out[key] = exp2f(fmaf(s[nt][0], sl2, -mb0)) * il0;
            if (key + 1 < T) pa_out[key + 1] = exp2f(fmaf(s[nt][1], sl2, -mb0)) * il0;
          }
          if (row_b < T) {
            if (key < T) pb_out[key] = exp2f(fmaf(s[nt][2], sl2, -mb1)) * il1;
            if (key + 1 < T) pb_out[key + 1] = exp2f(fmaf(s[nt][3], sl2, -mb1)) * il1;
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// CLS-query attention for the LAST block: VisionTransformer.forward returns only x[:, 0] (vision_transformer.py:
// 213-214), so after the last block's K and V are known, every other query row of that block — and every other row of
// its proj / MLP — is dead work.  One CTA per (frame, head): logits of the CLS query against the 197 keys, exact
// softmax, then the probability-weighted sum of V.  fp32 math on bf16 q/k/v,
// bf16 output [B,384] — the same rounding points as the full kernel.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) vit_cls_attention_kernel(const __nv_bfloat16* __restrict__ qkv, int B,
                                                                __nv_bfloat16* __restrict__ out_cls) {
  pdl_trigger();
  pdl_wait();  // (PDL, common.cuh) no global access above this line
  // One CTA (4 warps) per (frame, head).  lane = (key group kg = lane / 8, 16-byte chunk dc = lane % 8): one warp
  // instruction covers four whole 128-byte K (or V) rows, fully coalesced; warp w takes keys 16 i + 4 w + kg, i < 13, and
  // issues all 13 K-row and 13 V-row loads up front (26 x 16 bytes in flight per thread) — the kernel is a pure
  // stream over K and V (77 MB at batch 256), so memory-level parallelism is what sets its time.
  constexpr int NI = 13;  // ceil(197 / 16)
  __shared__ float red_s[8];
  __shared__ float o_s[4][HD];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x;
  const int b = item / HEADS, h = item % HEADS;
  const int kg = lane >> 3, dc = lane & 7;
  const __nv_bfloat16* base = qkv + int64_t(b) * T * QKV_LD + h * HD;
  uint4 kr[NI], vr[NI];
  const uint4 qq = __ldg(reinterpret_cast<const uint4*>(base) + dc);  // CLS token = row 0 of the frame
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int j = 16 * i + 4 * warp + kg;
    kr[i] = (j < T) ? __ldg(reinterpret_cast<const uint4*>(base + int64_t(j) * QKV_LD + 384) + dc) : make_uint4(0, 0, 0, 0);
  }
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int j = 16 * i + 4 * warp + kg;
    vr[i] = (j < T) ? __ldg(reinterpret_cast<const uint4*>(base + int64_t(j) * QKV_LD + 768) + dc) : make_uint4(0, 0, 0, 0);
  }
  float q[8];
  q[0] = bf16_lo(qq.x); q[1] = bf16_hi(qq.x); q[2] = bf16_lo(qq.y); q[3] = bf16_hi(qq.y);
  q[4] = bf16_lo(qq.z); q[5] = bf16_hi(qq.z); q[6] = bf16_lo(qq.w); q[7] = bf16_hi(qq.w);
  float m = -INFINITY;
  float lg[NI];
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int j = 16 * i + 4 * warp + kg;
    const uint4 kk = kr[i];
    float dot = 0.f;
    dot = fmaf(bf16_lo(kk.x), q[0], dot); dot = fmaf(bf16_hi(kk.x), q[1], dot);
    dot = fmaf(bf16_lo(kk.y), q[2], dot); dot = fmaf(bf16_hi(kk.y), q[3], dot);
    dot = fmaf(bf16_lo(kk.z), q[4], dot); dot = fmaf(bf16_hi(kk.z), q[5], dot);
    dot = fmaf(bf16_lo(kk.w), q[6], dot); dot = fmaf(bf16_hi(kk.w), q[7], dot);
    dot += __shfl_xor_sync(0xffffffffu, dot, 1);
    dot += __shfl_xor_sync(0xffffffffu, dot, 2);
    dot += __shfl_xor_sync(0xffffffffu, dot, 4);
    dot *= 0.125f;  // head_dim^-0.5
    lg[i] = dot;
    if (j < T) m = fmaxf(m, dot);
  }
  m = warp_max(m);
  if (lane == 0) red_s[warp] = m;
  __syncthreads();
  m = fmaxf(fmaxf(red_s[0], red_s[1]), fmaxf(red_s[2], red_s[3]));
  // every lane of a key group holds that key's logit: exponentials and the row sum come straight from registers
  float sum = 0.f;
  float pr[NI];
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int j = 16 * i + 4 * warp + kg;
    pr[i] = (j < T) ? __expf(lg[i] - m) : 0.f;
    if (dc == 0) sum += pr[i];
  }
  sum = warp_sum(sum);
  if (lane == 0) red_s[4 + warp] = sum;
  float o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = 0.f;
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const uint4 vv = vr[i];
    const float pj = pr[i];
    o[0] = fmaf(pj, bf16_lo(vv.x), o[0]); o[1] = fmaf(pj, bf16_hi(vv.x), o[1]);
    o[2] = fmaf(pj, bf16_lo(vv.y), o[2]); o[3] = fmaf(pj, bf16_hi(vv.y), o[3]);
    o[4] = fmaf(pj, bf16_lo(vv.z), o[4]); o[5] = fmaf(pj, bf16_hi(vv.z), o[5]);
    o[6] = fmaf(pj, bf16_lo(vv.w), o[6]); o[7] = fmaf(pj, bf16_hi(vv.w), o[7]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {  // sum the four key groups of the warp
    o[i] += __shfl_xor_sync(0xffffffffu, o[i], 8);
    o[i] += __shfl_xor_sync(0xffffffffu, o[i], 16);
  }
  if (kg == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) o_s[warp][dc * 8 + i] = o[i];
  }
  __syncthreads();
  if (warp == 0 && lane < 8) {  // fixed-order sum over the four warps: deterministic
    const float inv = 1.0f / ((red_s[4] + red_s[5]) + (red_s[6] + red_s[7]));
    float r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
      r[i] = ((o_s[0][lane * 8 + i] + o_s[1][lane * 8 + i]) + (o_s[2][lane * 8 + i] + o_s[3][lane * 8 + i])) * inv;
    uint4 w;
    w.x = pack_bf16x2(r[0], r[1]);
    w.y = pack_bf16x2(r[2], r[3]);
    w.z = pack_bf16x2(r[4], r[5]);
    w.w = pack_bf16x2(r[6], r[7]);
    *(reinterpret_cast<uint4*>(out_cls + int64_t(b) * OUT_LD + h * HD) + lane) = w;
  }
}

}  // namespace

int vit_cls_attention(const sais_bf16* qkv, int B, sais_bf16* out_cls, cudaStream_t stream) {
  if (B == 0) return kOk;
  if (!qkv || !out_cls || B < 0) {
    set_last_error("vit_cls_attention: bad arguments");
    return kErrInvalidArg;
  }
  LaunchScope ls(kClsVitAttn, stream, 4.0 * double(B) * 6 * 197 * 64);
  return check_cuda(launch_pdl(vit_cls_attention_kernel, dim3(B * HEADS), dim3(128), size_t(0), stream, 1, reinterpret_cast<const __nv_bfloat16*>(qkv), B,
                                                                    reinterpret_cast<__nv_bfloat16*>(out_cls)),
                    "vit_cls_attention launch");
}

int vit_attention(const sais_bf16* qkv, int B, sais_bf16* out, float* probs, cudaStream_t stream) {
  if (B == 0) return kOk;
  if (!qkv || !out || B < 0) {
    set_last_error("vit_attention: bad arguments");
    return kErrInvalidArg;
  }
  // default path: tcgen05 kernel (vit_attention_tc.cu); this register-level kernel serves the probabilities-
  // emitting variant (and SAIS_ATTN_LEGACY=1 for A/B comparisons)
  static const bool legacy = getenv("SAIS_ATTN_LEGACY") != nullptr && atoi(getenv("SAIS_ATTN_LEGACY")) != 0;
  if (probs == nullptr && !legacy) return vit_attention_tc(qkv, B, out, stream);
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(vit_attention_kernel<false>), kSmem, "vit_attention"))
    return rc;
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(vit_attention_kernel<true>), kSmem, "vit_attention probs"))
    return rc;
  const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(qkv);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  LaunchScope ls(kClsVitAttn, stream, 4.0 * double(B) * 6 * 197 * 197 * 64);
  auto kern = probs ? vit_attention_kernel<true> : vit_attention_kernel<false>;
  return check_cuda(launch_pdl(kern, dim3(B * HEADS), dim3(kThreads), size_t(kSmem), stream, 1, q, o, probs),
                    "vit_attention launch");
}

}  // namespace sais

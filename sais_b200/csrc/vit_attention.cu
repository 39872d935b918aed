// ViT self-attention entry points (T = 197 tokens, 6 heads x 64; vision_transformer.py:80-92).
//   vit_attention     : every query row — dispatches to the tcgen05 kernel (vit_attention_tc.cu), which also serves
//                       the probabilities-emitting variant of get_last_selfattention (:216-223);
//   vit_cls_attention : the CLS query only (last block when neither all tokens nor the probabilities are wanted):
//                       a small SIMT kernel, K and V rows streamed once, coalesced — memory-bound by construction.
// (The round-1 register-level mma.sync kernel that used to emit the probabilities is gone: no legacy tensor path is left.)
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

// ---------------------------------------------------------------------------------------------------------------
// CLS-query attention for the LAST block: VisionTransformer.forward returns only x[:, 0] (vision_transformer.py:
// 213-214), so after the last block's K and V are known, every other query row of that block — and every other row of
// its proj / MLP — is dead work.  One CTA per (frame, head): logits of the CLS query against the 197 keys, exact
// softmax, then the probability-weighted sum of V.  fp32 math on bf16 q/k/v,
// bf16 output [B,384] — the same rounding points as the full kernel.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) vit_cls_attention_kernel(const __nv_bfloat16* __restrict__ qkv, int B,
                                                                __nv_bfloat16* __restrict__ out_cls) {
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
  return vit_attention_tc(qkv, B, out, probs, stream);
}

}  // namespace sais

// Memory-bound kernels of the SAIS hot path: LayerNorm, frame normalisation + patch layout,
// CLS/positional-embedding assembly, clip head and prototype scoring.  All are HBM/L2-bound:
// 16-byte vector accesses, warp-shuffle reductions, no shared-memory staging where there is no reuse.
#include "common.cuh"
#include "kernels.h"
#include "rowcast.cuh"

namespace sais {

namespace {

constexpr int D = 384;

// ----------------------------------------------------------------------------------------------
// LayerNorm(384): one warp per row, the row lives in registers (12 floats per lane), two-pass
// statistics in fp32 (matches torch's mean / biased variance).
// Reference: nn.LayerNorm in vision_transformer.py:99,103,156 (eps 1e-6) and
// nn.TransformerEncoderLayer.norm1/norm2 (eps 1e-5).
// ----------------------------------------------------------------------------------------------
constexpr int kLnRows = 4;  // rows per warp per iteration: 12 independent 16-byte loads in flight per lane

// Fan-out of the fp32 output rows to the other GPUs of a symmetric-memory group (SaisFanout, include/sais_b200.h): the
// exchange step of the path fused into the kernel that produces the embeddings.  mc != nullptr: one multimem.st per 16
// bytes to the NVSwitch multicast mapping (replicated to every GPU by the switch); else n plain stores to peer mappings.
struct LnFan {
  float* mc;
  float* peer[SAIS_MAX_PEERS];
  int n;
};
__device__ __forceinline__ void multimem_st_f32x4(float* addr, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

__global__ void __launch_bounds__(256) layernorm384_kernel(const float* __restrict__ x, int64_t in_pitch,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float eps, int64_t rows,
                                                           float* __restrict__ out_f32,
                                                           __nv_bfloat16* __restrict__ out_bf16, int split,
                                                           float* __restrict__ out_plus,
                                                           const float* __restrict__ plus_vec, const LnFan fan) {
  pdl_wait();  // (PDL, common.cuh) no global access above this line
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t warps_total = int64_t(gridDim.x) * (blockDim.x >> 5);
  float4 g[3], b[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    g[i] = __ldg(reinterpret_cast<const float4*>(gamma + i * 128 + lane * 4));
    b[i] = __ldg(reinterpret_cast<const float4*>(beta + i * 128 + lane * 4));
  }
  for (int64_t row0 = warp_global * kLnRows; row0 < rows; row0 += warps_total * kLnRows) {
    float4 v[kLnRows][3];
#pragma unroll
    for (int r = 0; r < kLnRows; ++r) {
      const int64_t row = (row0 + r < rows) ? row0 + r : rows - 1;  // clamp: tail rows are recomputed, not stored
      const float* xr = x + row * in_pitch;
#pragma unroll
      for (int i = 0; i < 3; ++i) v[r][i] = *reinterpret_cast<const float4*>(xr + i * 128 + lane * 4);
    }
    float mean[kLnRows], rstd[kLnRows];
#pragma unroll
    for (int r = 0; r < kLnRows; ++r) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 3; ++i) s += (v[r][i].x + v[r][i].y) + (v[r][i].z + v[r][i].w);
      mean[r] = s;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int r = 0; r < kLnRows; ++r) mean[r] += __shfl_xor_sync(0xffffffffu, mean[r], o);
    }
#pragma unroll
    for (int r = 0; r < kLnRows; ++r) {
      mean[r] *= (1.0f / D);
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float a = v[r][i].x - mean[r], bb = v[r][i].y - mean[r], c = v[r][i].z - mean[r],
                    d = v[r][i].w - mean[r];
        ss += (a * a + bb * bb) + (c * c + d * d);
      }
      rstd[r] = ss;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int r = 0; r < kLnRows; ++r) rstd[r] += __shfl_xor_sync(0xffffffffu, rstd[r], o);
    }
#pragma unroll
    for (int r = 0; r < kLnRows; ++r) {
      const int64_t row = row0 + r;
      if (row >= rows) break;
      const float rs = rsqrtf(rstd[r] * (1.0f / D) + eps);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int c = i * 128 + lane * 4;
        float4 y;
        y.x = (v[r][i].x - mean[r]) * rs * g[i].x + b[i].x;
        y.y = (v[r][i].y - mean[r]) * rs * g[i].y + b[i].y;
        y.z = (v[r][i].z - mean[r]) * rs * g[i].z + b[i].z;
        y.w = (v[r][i].w - mean[r]) * rs * g[i].w + b[i].w;
        if (out_f32) *reinterpret_cast<float4*>(out_f32 + row * D + c) = y;
        if (fan.mc != nullptr) {
          multimem_st_f32x4(fan.mc + row * D + c, y);
        } else {
          for (int pi = 0; pi < fan.n; ++pi) *reinterpret_cast<float4*>(fan.peer[pi] + row * D + c) = y;
        }
        if (out_plus) {  // y + vector: the pre-loaded accumulator of the next residual GEMM (accumulate mode)
          const float4 pv = __ldg(reinterpret_cast<const float4*>(plus_vec + c));
          *reinterpret_cast<float4*>(out_plus + row * D + c) = make_float4(y.x + pv.x, y.y + pv.y, y.z + pv.z, y.w + pv.w);
        }
        if (out_bf16) {
          uint2 o;
          o.x = pack_bf16x2(y.x, y.y);
          o.y = pack_bf16x2(y.z, y.w);
          if (!split) {
            *reinterpret_cast<uint2*>(out_bf16 + row * D + c) = o;
          } else {  // [hi | lo] halves for the split-precision GEMM
            uint2 l;
            l.x = pack_bf16x2(y.x - bf16_lo(o.x), y.y - bf16_hi(o.x));
            l.y = pack_bf16x2(y.z - bf16_lo(o.y), y.w - bf16_hi(o.y));
            *reinterpret_cast<uint2*>(out_bf16 + row * 2 * D + c) = o;
            *reinterpret_cast<uint2*>(out_bf16 + row * 2 * D + D + c) = l;
          }
        }
      }
    }
  }
}

// Row statistics + bf16 cast of the fp32 residual stream: the entry point of the LayerNorm-folded fast path (the
// GEMM that consumes the rows applies mean / rstd in its epilogue, see gemm_tcgen05.cu).  One warp per row.
// stats[row] = {sum, sum of squares, 0, 0, 0, 0, 0, 0}  (slot 0 of the four partial slots the GEMM epilogues fill).
__global__ void __launch_bounds__(256) rowstats_cast384_kernel(const float* __restrict__ x, int64_t rows,
                                                               __nv_bfloat16* __restrict__ xb,
                                                               float* __restrict__ stats, int reverse) {
  pdl_wait();  // (PDL, common.cuh) no global access above this line
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t warps_total = int64_t(gridDim.x) * (blockDim.x >> 5);
  // two rows per warp iteration: all six 16-byte loads of a lane are in flight before the first reduction (the one-row loop
  // left the memory pipeline idle during the shuffles: 20 us for 118 MB)
  for (int64_t r0 = warp_global * 2; r0 < rows; r0 += warps_total * 2) {
    const bool two = (r0 + 1 < rows);
    const int64_t ra = reverse ? rows - 1 - r0 : r0;  // (kernels.h g_tile_reverse: start on the rows written last)
    rowcast_rows<2, false>(x, xb, stats, ra, reverse ? -1 : 1, two ? 2 : 1, lane);
  }
}

// ----------------------------------------------------------------------------------------------
// Frame normalisation fused with the patch layout the patch-embed GEMM consumes.
// One thread = one (frame, image row y, patch column px): reads 16 pixels x 3 channels (48 bytes,
// contiguous) and writes three 32-byte runs, one per channel, at k = c*256 + (y%16)*16.
// Reference: ToTensor + Normalize, extract_representations.py:158-162; PatchEmbed conv weight
// layout [384,3,16,16], vision_transformer.py:126.
// ----------------------------------------------------------------------------------------------
struct NormConsts {
  float scale[3];  // 1 / (255 * std)
  float shift[3];  // -mean / std
};

__global__ void __launch_bounds__(256) normalize_patchify_u8_kernel(const uint8_t* __restrict__ frames, int B,
                                                                    NormConsts nc,
                                                                    __nv_bfloat16* __restrict__ patches, int split) {
  pdl_wait();  // (PDL, common.cuh) no global access above this line
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = int64_t(B) * 224 * 14;
  if (t >= total) return;
  const int px = int(t % 14);
  const int y = int((t / 14) % 224);
  const int64_t b = t / (14 * 224);
  const uint8_t* src = frames + ((b * 224 + y) * 224 + px * 16) * 3;
  uint32_t w[12];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(src) + i);
    w[4 * i] = u.x; w[4 * i + 1] = u.y; w[4 * i + 2] = u.z; w[4 * i + 3] = u.w;
  }
  const int py = y >> 4, ky = y & 15;
  const int pitch = split ? 1536 : 768;
  __nv_bfloat16* dst = patches + (b * 196 + py * 14 + px) * pitch + ky * 16;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    uint32_t o[8], l[8];
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      const int i0 = j * 3 + c, i1 = (j + 1) * 3 + c;
      const float f0 = float((w[i0 >> 2] >> ((i0 & 3) * 8)) & 0xff);
      const float f1 = float((w[i1 >> 2] >> ((i1 & 3) * 8)) & 0xff);
      const float v0 = fmaf(f0, nc.scale[c], nc.shift[c]), v1 = fmaf(f1, nc.scale[c], nc.shift[c]);
      o[j >> 1] = pack_bf16x2(v0, v1);
      l[j >> 1] = pack_bf16x2(v0 - bf16_lo(o[j >> 1]), v1 - bf16_hi(o[j >> 1]));
    }
    uint4* d4 = reinterpret_cast<uint4*>(dst + c * 256);
    d4[0] = make_uint4(o[0], o[1], o[2], o[3]);
    d4[1] = make_uint4(o[4], o[5], o[6], o[7]);
    if (split) {
      uint4* l4 = reinterpret_cast<uint4*>(dst + 768 + c * 256);
      l4[0] = make_uint4(l[0], l[1], l[2], l[3]);
      l4[1] = make_uint4(l[4], l[5], l[6], l[7]);
    }
  }
}

// fp32 NCHW (already normalised, what the reference model is called with) -> bf16 patch matrix.
__global__ void __launch_bounds__(256) patchify_f32_kernel(const float* __restrict__ frames, int B,
                                                           __nv_bfloat16* __restrict__ patches, int split) {
  pdl_wait();  // (PDL, common.cuh) no global access above this line
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = int64_t(B) * 3 * 224 * 14;
  if (t >= total) return;
  const int px = int(t % 14);
  const int y = int((t / 14) % 224);
  const int c = int((t / (14 * 224)) % 3);
  const int64_t b = t / (14 * 224 * 3);
  const float4* src = reinterpret_cast<const float4*>(frames + ((b * 3 + c) * 224 + y) * 224 + px * 16);
  const int py = y >> 4, ky = y & 15;
  uint32_t o[8], l[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 f = __ldg(src + i);
    o[2 * i] = pack_bf16x2(f.x, f.y);
    o[2 * i + 1] = pack_bf16x2(f.z, f.w);
    l[2 * i] = pack_bf16x2(f.x - bf16_lo(o[2 * i]), f.y - bf16_hi(o[2 * i]));
    l[2 * i + 1] = pack_bf16x2(f.z - bf16_lo(o[2 * i + 1]), f.w - bf16_hi(o[2 * i + 1]));
  }
  const int pitch = split ? 1536 : 768;
  __nv_bfloat16* dst = patches + (b * 196 + py * 14 + px) * pitch + c * 256 + ky * 16;
  uint4* d4 = reinterpret_cast<uint4*>(dst);
  d4[0] = make_uint4(o[0], o[1], o[2], o[3]);
  d4[1] = make_uint4(o[4], o[5], o[6], o[7]);
  if (split) {
    uint4* l4 = reinterpret_cast<uint4*>(dst + 768);
    l4[0] = make_uint4(l[0], l[1], l[2], l[3]);
    l4[1] = make_uint4(l[4], l[5], l[6], l[7]);
  }
}

// seq_offsets[i] = i * stride (packed-sequence offsets of equally long sequences, built on device)
__global__ void fill_offsets_kernel(int32_t* __restrict__ offs, int n, int stride) {
  pdl_wait();  // (PDL, common.cuh) no global access above this line
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) offs[t] = t * stride;
}

// x[b, 0, :] = cls_token + pos_embed[0]   (vision_transformer.py:201-205)
__global__ void write_cls_rows_kernel(const float* __restrict__ cls_pos0, int B, float* __restrict__ x) {
  pdl_wait();  // (PDL, common.cuh) no global access above this line
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * (D / 4)) return;
  const int b = t / (D / 4), c = (t % (D / 4)) * 4;
  *reinterpret_cast<float4*>(x + int64_t(b) * 197 * D + c) = __ldg(reinterpret_cast<const float4*>(cls_pos0 + c));
}

// ----------------------------------------------------------------------------------------------
// Temporal token assembly (prepare_model.py:189-194): token 0 = frame_cls, token s = frame[s-1] + pos[s-1].
// One block per packed sequence.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) temporal_prep_kernel(const float* __restrict__ x_frames,
                                                            const int32_t* __restrict__ seq_offsets,
                                                            const float* __restrict__ frame_cls,
                                                            const float* __restrict__ frame_pos,
                                                            float* __restrict__ tok_f32,
                                                            __nv_bfloat16* __restrict__ tok_bf16,
                                                            float* __restrict__ tok_plus,
                                                            const float* __restrict__ plus_vec) {
  pdl_wait();  // (PDL, common.cuh) no global access above this line
  const int i = blockIdx.x;
  const int t0 = seq_offsets[i];
  const int S = seq_offsets[i + 1] - t0;
  const float* xf = x_frames + int64_t(t0 - i) * D;
  for (int e = threadIdx.x; e < S * (D / 4); e += blockDim.x) {
    const int s = e / (D / 4), c = (e % (D / 4)) * 4;
    float4 v;
    if (s == 0) {
      v = __ldg(reinterpret_cast<const float4*>(frame_cls + c));
    } else {
      const float4 a = __ldg(reinterpret_cast<const float4*>(xf + int64_t(s - 1) * D + c));
      const float4 p = __ldg(reinterpret_cast<const float4*>(frame_pos + int64_t(s - 1) * D + c));
      v = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
    }
    const int64_t o = int64_t(t0 + s) * D + c;
    *reinterpret_cast<float4*>(tok_f32 + o) = v;
    if (tok_plus) {
      const float4 pv = __ldg(reinterpret_cast<const float4*>(plus_vec + c));
      *reinterpret_cast<float4*>(tok_plus + o) = make_float4(v.x + pv.x, v.y + pv.y, v.z + pv.z, v.w + pv.w);
    }
    uint2 b, l;
    b.x = pack_bf16x2(v.x, v.y);
    b.y = pack_bf16x2(v.z, v.w);
    l.x = pack_bf16x2(v.x - bf16_lo(b.x), v.y - bf16_hi(b.x));
    l.y = pack_bf16x2(v.z - bf16_lo(b.y), v.w - bf16_hi(b.y));
    const int64_t o2 = int64_t(t0 + s) * 2 * D + c;  // [hi | lo] halves, row pitch 768
    *reinterpret_cast<uint2*>(tok_bf16 + o2) = b;
    *reinterpret_cast<uint2*>(tok_bf16 + o2 + D) = l;
  }
}

// out_cls[i] = relu(tokens[seq_offsets[i]])   (prepare_model.py:215,220)
__global__ void gather_cls_relu_kernel(const float* __restrict__ tok, const int32_t* __restrict__ seq_offsets,
                                       int nseq, float* __restrict__ out_cls) {
  pdl_wait();  // (PDL, common.cuh) no global access above this line
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nseq * (D / 4)) return;
  const int i = t / (D / 4), c = (t % (D / 4)) * 4;
  float4 v = *reinterpret_cast<const float4*>(tok + int64_t(seq_offsets[i]) * D + c);
  v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
  *reinterpret_cast<float4*>(out_cls + int64_t(i) * D + c) = v;
}

// ----------------------------------------------------------------------------------------------
// Clip head (prepare_model.py:378-382, 405, 409): v = relu(mean_s a[b,s] + mean_s f[b,s]); out = W v + bias.
// fp32 throughout.  4 clips per block so each W row is read once per 4 clips; warp per output row.
// ----------------------------------------------------------------------------------------------
constexpr int kClipsPerBlock = 4;
constexpr int kOut = 256;

__global__ void __launch_bounds__(256) clip_head_kernel(const float* __restrict__ cls_a,
                                                        const float* __restrict__ cls_b, int B, int nsnip,
                                                        const float* __restrict__ W, const float* __restrict__ bias,
                                                        float* __restrict__ out) {
  pdl_wait();  // (PDL, common.cuh) no global access above this line
  __shared__ float v[kClipsPerBlock][D];
  const int b0 = blockIdx.x * kClipsPerBlock;
  const float inv = 1.0f / float(nsnip);
  for (int e = threadIdx.x; e < kClipsPerBlock * D; e += blockDim.x) {
    const int cb = e / D, k = e % D;
    const int b = b0 + cb;
    float acc = 0.f;
    if (b < B) {
      float sa = 0.f, sb = 0.f;
      for (int s = 0; s < nsnip; ++s) {
        sa += cls_a[(int64_t(b) * nsnip + s) * D + k];
        if (cls_b) sb += cls_b[(int64_t(b) * nsnip + s) * D + k];
      }
      // torch.mean(...) per modality, then the sum of the two modalities, then ReLU
      acc = sa * inv + (cls_b ? sb * inv : 0.f);
      acc = fmaxf(acc, 0.f);
    }
    v[cb][k] = acc;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // blockIdx.y owns 32 of the 256 output rows (4 per warp): with a handful of clips the kernel is a chain of dependent
  // weight-row loads, so the rows are spread over 8x more CTAs instead of walked 32 deep by each warp
  for (int o = blockIdx.y * (kOut / 8) + warp; o < (blockIdx.y + 1) * (kOut / 8); o += 8) {
    float w[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) w[i] = __ldg(W + int64_t(o) * D + i * 32 + lane);
    const float bo = bias ? __ldg(bias + o) : 0.f;
#pragma unroll
    for (int cb = 0; cb < kClipsPerBlock; ++cb) {
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < 12; ++i) acc = fmaf(w[i], v[cb][i * 32 + lane], acc);
      acc = warp_sum(acc);
      if (lane == 0 && b0 + cb < B) out[int64_t(b0 + cb) * kOut + o] = acc + bo;
    }
  }
}

// ----------------------------------------------------------------------------------------------
// Prototype scoring (prepare_miscellaneous.py:102-125; process_inference_results.py:76-91):
// s_n = s/||s||, p_n = p/||p||, sim = s_n·p_n, probs = exp(sim)/sum exp(sim), pred = argmax.
// One warp per clip; prototypes (P <= 64) are re-normalised per warp from L1/L2.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) prototype_score_kernel(const float* __restrict__ reps,
                                                              const float* __restrict__ protos, int B, int P, int Dd,
                                                              float* __restrict__ probs, float* __restrict__ sims,
                                                              int32_t* __restrict__ pred) {
  pdl_wait();  // (PDL, common.cuh) no global access above this line
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const float* s = reps + int64_t(b) * Dd;
  float ss = 0.f;
  for (int k = lane; k < Dd; k += 32) ss = fmaf(s[k], s[k], ss);
  const float sn = sqrtf(warp_sum(ss));
  float my_e = 0.f, my_sim = 0.f;  // lane p%32 keeps prototype p (and p+32 in the second slot)
  float my_e2 = 0.f, my_sim2 = 0.f;
  float esum = 0.f;
  for (int p = 0; p < P; ++p) {
    const float* pr = protos + int64_t(p) * Dd;
    float pp = 0.f;
    for (int k = lane; k < Dd; k += 32) pp = fmaf(pr[k], pr[k], pp);
    const float pn = sqrtf(warp_sum(pp));
    float dot = 0.f;
    for (int k = lane; k < Dd; k += 32) dot = fmaf(s[k] / sn, pr[k] / pn, dot);
    dot = warp_sum(dot);
    const float e = expf(dot);
    esum += e;
    if ((p & 31) == lane) {
      if (p < 32) { my_e = e; my_sim = dot; } else { my_e2 = e; my_sim2 = dot; }
    }
  }
  // argmax over probabilities (first maximal index, like torch.argmax)
  float best = -1.f;
  int best_i = 0x7fffffff;
  if (lane < P) { best = my_e / esum; best_i = lane; }
  if (lane + 32 < P) {
    const float pv = my_e2 / esum;
    if (pv > best) { best = pv; best_i = lane + 32; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
  }
  if (lane < P) {
    probs[int64_t(b) * P + lane] = my_e / esum;
    if (sims) sims[int64_t(b) * P + lane] = my_sim;
  }
  if (lane + 32 < P) {
    probs[int64_t(b) * P + lane + 32] = my_e2 / esum;
    if (sims) sims[int64_t(b) * P + lane + 32] = my_sim2;
  }
  if (lane == 0 && pred) pred[b] = best_i;
}

}  // namespace

int layernorm(const float* x, int64_t in_pitch, const float* gamma, const float* beta, float eps, int64_t rows,
              float* out_f32, sais_bf16* out_bf16, cudaStream_t stream, int split, float* out_plus,
              const float* plus_vec, const SaisFanout* fan, int64_t fan_offset) {
  if (rows == 0) return kOk;
  if (!x || !gamma || !beta || (!out_f32 && !out_bf16) || rows < 0 || in_pitch % 4 || (out_plus && !plus_vec)) {
    set_last_error("layernorm: bad arguments");
    return kErrInvalidArg;
  }
  LnFan lf;
  memset(&lf, 0, sizeof(lf));
  if (fan != nullptr && (fan->multicast != nullptr || fan->n_peers > 0)) {
    if (!out_f32 || fan->n_peers < 0 || fan->n_peers > SAIS_MAX_PEERS) {
      set_last_error("layernorm: fan-out needs the fp32 output and 0..%d peers", SAIS_MAX_PEERS);
      return kErrInvalidArg;
    }
    uintptr_t bits = reinterpret_cast<uintptr_t>(fan->multicast);
    if (fan->multicast != nullptr) {
      lf.mc = static_cast<float*>(fan->multicast) + fan_offset;
    } else {
      lf.n = fan->n_peers;
      for (int i = 0; i < lf.n; ++i) {
        if (!fan->peers[i]) {
          set_last_error("layernorm: fan-out peer %d is NULL", i);
          return kErrInvalidArg;
        }
        bits |= reinterpret_cast<uintptr_t>(fan->peers[i]);
        lf.peer[i] = static_cast<float*>(fan->peers[i]) + fan_offset;
      }
    }
    if ((bits | uintptr_t(fan_offset * 4)) & 15) {
      set_last_error("layernorm: fan-out addresses must be 16-byte aligned");
      return kErrInvalidArg;
    }
  }
  int64_t blocks = (rows + 8 * kLnRows - 1) / (8 * kLnRows);
  const int64_t cap = int64_t(num_sms()) * 8;  // persistent beyond one full wave (8 blocks x 8 warps per SM)
  if (blocks > cap) blocks = cap;
  LaunchScope ls(kClsLayerNorm, stream, double(rows) * D * (4 + (out_f32 ? 4 : 0) + (out_bf16 ? 2 : 0)));
  return check_cuda(launch_pdl(layernorm384_kernel, dim3(unsigned(blocks)), dim3(256), size_t(0), stream, 1, x, in_pitch, gamma, beta, eps, rows, out_f32,
                                                            reinterpret_cast<__nv_bfloat16*>(out_bf16), split, out_plus,
                                                            plus_vec, lf),
                    "layernorm launch");
}

int rowstats_cast(const float* x, int64_t rows, sais_bf16* xb, float* stats, cudaStream_t stream) {
  if (rows == 0) return kOk;
  if (!x || !xb || !stats || rows < 0) {
    set_last_error("rowstats_cast: bad arguments");
    return kErrInvalidArg;
  }
  int64_t blocks = (rows + 15) / 16;  // 8 warps per block, two rows per warp iteration
  const int64_t cap = int64_t(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  LaunchScope ls(kClsLayerNorm, stream, double(rows) * D * 6);
  return check_cuda(launch_pdl(rowstats_cast384_kernel, dim3(unsigned(blocks)), dim3(256), size_t(0), stream, 1, x, rows, reinterpret_cast<__nv_bfloat16*>(xb), stats, g_tile_reverse),
                    "rowstats_cast launch");
}

int normalize_patchify_u8(const uint8_t* frames, int B, const float* mean3, const float* std3, sais_bf16* patches,
                          cudaStream_t stream, int split) {
  if (B == 0) return kOk;
  if (!frames || !patches || !mean3 || !std3 || B < 0) {
    set_last_error("normalize_patchify_u8: bad arguments");
    return kErrInvalidArg;
  }
  NormConsts nc;
  for (int c = 0; c < 3; ++c) {
    nc.scale[c] = 1.0f / (255.0f * std3[c]);
    nc.shift[c] = -mean3[c] / std3[c];
  }
  const int64_t total = int64_t(B) * 224 * 14;
  LaunchScope ls(kClsPatchify, stream, double(B) * 224 * 224 * 3 * 3);
  return check_cuda(launch_pdl(normalize_patchify_u8_kernel, dim3(unsigned((total + 255) / 256)), dim3(256), size_t(0), stream, 1, 
      frames, B, nc, reinterpret_cast<__nv_bfloat16*>(patches), split),
                    "normalize_patchify_u8 launch");
}

int patchify_f32(const float* frames, int B, sais_bf16* patches, cudaStream_t stream, int split) {
  if (B == 0) return kOk;
  if (!frames || !patches || B < 0) {
    set_last_error("patchify_f32: bad arguments");
    return kErrInvalidArg;
  }
  const int64_t total = int64_t(B) * 3 * 224 * 14;
  LaunchScope ls(kClsPatchify, stream, double(B) * 224 * 224 * 3 * 6);
  return check_cuda(launch_pdl(patchify_f32_kernel, dim3(unsigned((total + 255) / 256)), dim3(256), size_t(0), stream, 1, 
      frames, B, reinterpret_cast<__nv_bfloat16*>(patches), split),
                    "patchify_f32 launch");
}

int fill_offsets(int32_t* offs, int n, int stride, cudaStream_t stream) {
  if (n <= 0) return kOk;
  LaunchScope ls(kClsMisc, stream, double(n) * 4);
  return check_cuda(launch_pdl(fill_offsets_kernel, dim3((n + 255) / 256), dim3(256), size_t(0), stream, 1, offs, n, stride),
                    "fill_offsets launch");
}

int write_cls_rows(const float* cls_pos0, int B, float* x, cudaStream_t stream) {
  if (B == 0) return kOk;
  const int total = B * (D / 4);
  LaunchScope ls(kClsMisc, stream, double(B) * D * 8);
  return check_cuda(launch_pdl(write_cls_rows_kernel, dim3((total + 255) / 256), dim3(256), size_t(0), stream, 1, cls_pos0, B, x),
                    "write_cls_rows launch");
}

int temporal_prep(const float* x_frames, const int32_t* seq_offsets, int nseq, int total_tokens,
                  const float* frame_cls, const float* frame_pos, int n_pos, float* tok_f32, sais_bf16* tok_bf16,
                  cudaStream_t stream, float* tok_plus, const float* plus_vec) {
  (void)n_pos;
  if (nseq == 0) return kOk;
  if (!seq_offsets || !frame_cls || !frame_pos || !tok_f32 || !tok_bf16 || nseq < 0) {
    set_last_error("temporal_prep: bad arguments");
    return kErrInvalidArg;
  }
  LaunchScope ls(kClsMisc, stream, double(total_tokens) * D * 14);
  return check_cuda(launch_pdl(temporal_prep_kernel, dim3(nseq), dim3(128), size_t(0), stream, 1, x_frames, seq_offsets, frame_cls, frame_pos, tok_f32,
                                                 reinterpret_cast<__nv_bfloat16*>(tok_bf16), tok_plus, plus_vec),
                    "temporal_prep launch");
}

int gather_cls_relu(const float* tok_f32, const int32_t* seq_offsets, int nseq, float* out_cls,
                    cudaStream_t stream) {
  if (nseq == 0) return kOk;
  const int total = nseq * (D / 4);
  LaunchScope ls(kClsMisc, stream, double(nseq) * D * 8);
  return check_cuda(launch_pdl(gather_cls_relu_kernel, dim3((total + 255) / 256), dim3(256), size_t(0), stream, 1, tok_f32, seq_offsets, nseq, out_cls),
                    "gather_cls_relu launch");
}

int clip_head(const float* cls_a, const float* cls_b, int B, int nsnip, const float* lin_w, const float* lin_b,
              float* out, cudaStream_t stream) {
  if (B == 0) return kOk;
  if (!cls_a || !lin_w || !out || B < 0 || nsnip <= 0) {
    set_last_error("clip_head: bad arguments");
    return kErrInvalidArg;
  }
  LaunchScope ls(kClsMisc, stream, double(B) * (nsnip * D * 8 + 1024));
  return check_cuda(launch_pdl(clip_head_kernel, dim3((B + kClipsPerBlock - 1) / kClipsPerBlock, 8), dim3(256), size_t(0), stream, 1, cls_a, cls_b, B, nsnip, lin_w,
                                                                                 lin_b, out),
                    "clip_head launch");
}

// ----------------------------------------------------------------------------------------------
// MIL pathway (fullModel.forward task='MIL', prepare_model.py:359-363, 451-488).
//   add_pos_rows : tokens[b, s, :] = x[b, s, :] + pos[s, :]  — getClipReps' clip positional embeddings (:455-457)
//   mil_head     : ReLU of the clip encoder's output, then per class c the gated attention of calcAttention (:131-139):
//                  a = tanh(A r), g = sigmoid(B r), score_s = w_c . (a * g) + b_c, att = softmax over the snippets,
//                  video_rep = sum_s att_s r_s (:141-144), logit_c = f_c . video_rep + f_c0 (:146-149).
// One CTA per batch element; everything fp32 (a few thousand MACs per snippet: latency, not throughput).
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) add_pos_rows_kernel(const float* __restrict__ x, const float* __restrict__ pos,
                                                           int64_t rows, int period, float* __restrict__ out) {
  pdl_wait();
  const int64_t total = rows * (D / 4);
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = e / (D / 4);
    const int k4 = int(e - r * (D / 4));
    const float4 a = reinterpret_cast<const float4*>(x)[e];
    const float4 b = __ldg(reinterpret_cast<const float4*>(pos + (r % period) * D) + k4);
    reinterpret_cast<float4*>(out)[e] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
}

constexpr int kMilGate = 256;
constexpr int kMilMaxClasses = 3;  // attentionModules / finalModules hold three heads (prepare_model.py:84-101)

__global__ void __launch_bounds__(256) mil_head_kernel(const float* __restrict__ enc_out, int nsnip, int ncls,
                                                       const float* __restrict__ wa, const float* __restrict__ ba,
                                                       const float* __restrict__ wb, const float* __restrict__ bb,
                                                       const float* __restrict__ wc, const float* __restrict__ bc,
                                                       const float* __restrict__ wf, const float* __restrict__ bf,
                                                       float* __restrict__ reps_out, float* __restrict__ logits,
                                                       float* __restrict__ attn_out, int B) {
  pdl_wait();
  extern __shared__ __align__(16) float mil_s[];
  float* reps = mil_s;                       // [nsnip][D]   relu(enc_out)
  float* gated = reps + size_t(nsnip) * D;   // [nsnip][256] tanh(A r) * sigmoid(B r)
  float* score = gated + size_t(nsnip) * kMilGate;  // [kMilMaxClasses][nsnip]
  float* vrep = score + kMilMaxClasses * nsnip;     // [D]
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int e = threadIdx.x; e < nsnip * D; e += blockDim.x) {
    const float v = fmaxf(enc_out[int64_t(b) * nsnip * D + e], 0.f);
    reps[e] = v;
    reps_out[int64_t(b) * nsnip * D + e] = v;
  }
  __syncthreads();
  // gated attention features: warp w takes gate outputs w, w + 8, ... ; weight rows live in registers across the snippets
  for (int o = warp; o < kMilGate; o += 8) {
    float ra[12], rb[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      ra[i] = __ldg(wa + int64_t(o) * D + i * 32 + lane);
      rb[i] = __ldg(wb + int64_t(o) * D + i * 32 + lane);
    }
    const float ba_o = __ldg(ba + o), bb_o = __ldg(bb + o);
    for (int s = 0; s < nsnip; ++s) {
      float da = 0.f, db = 0.f;
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        const float r = reps[s * D + i * 32 + lane];
        da = fmaf(ra[i], r, da);
        db = fmaf(rb[i], r, db);
      }
      da = warp_sum(da);
      db = warp_sum(db);
      if (lane == 0) gated[s * kMilGate + o] = tanhf(da + ba_o) * (1.0f / (1.0f + expf(-(db + bb_o))));
    }
  }
  __syncthreads();
  // per-class snippet scores
  for (int e = warp; e < ncls * nsnip; e += 8) {
    const int c = e / nsnip, s = e - c * nsnip;
    float d = 0.f;
#pragma unroll
    for (int i = 0; i < kMilGate / 32; ++i) d = fmaf(__ldg(wc + c * kMilGate + i * 32 + lane), gated[s * kMilGate + i * 32 + lane], d);
    d = warp_sum(d);
    if (lane == 0) score[c * nsnip + s] = d + __ldg(bc + c);
  }
  __syncthreads();
  for (int c = 0; c < ncls; ++c) {
    // softmax over the snippets (every thread computes the same max / sum: nsnip is small)
    float m = -INFINITY;
    for (int s = 0; s < nsnip; ++s) m = fmaxf(m, score[c * nsnip + s]);
    float sum = 0.f;
    for (int s = 0; s < nsnip; ++s) sum += expf(score[c * nsnip + s] - m);
    const float inv = 1.0f / sum;
    for (int s = threadIdx.x; s < nsnip; s += blockDim.x)
      attn_out[(int64_t(c) * B + b) * nsnip + s] = expf(score[c * nsnip + s] - m) * inv;
    for (int k = threadIdx.x; k < D; k += blockDim.x) {
      float acc = 0.f;
      for (int s = 0; s < nsnip; ++s) acc = fmaf(expf(score[c * nsnip + s] - m) * inv, reps[s * D + k], acc);
      vrep[k] = acc;
    }
    __syncthreads();
    if (warp == 0) {
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < 12; ++i) d = fmaf(__ldg(wf + c * D + i * 32 + lane), vrep[i * 32 + lane], d);
      d = warp_sum(d);
      if (lane == 0) logits[int64_t(b) * ncls + c] = d + __ldg(bf + c);
    }
    __syncthreads();
  }
}

int add_pos_rows(const float* x, const float* pos, int64_t rows, int period, float* out, cudaStream_t stream) {
  if (rows == 0) return kOk;
  if (!x || !pos || !out || rows < 0 || period <= 0) {
    set_last_error("add_pos_rows: bad arguments");
    return kErrInvalidArg;
  }
  const int64_t total = rows * (D / 4);
  int64_t blocks = (total + 255) / 256;
  if (blocks > int64_t(num_sms()) * 8) blocks = int64_t(num_sms()) * 8;
  LaunchScope ls(kClsMisc, stream, double(rows) * D * 8);
  return check_cuda(launch_pdl(add_pos_rows_kernel, dim3(unsigned(blocks)), dim3(256), size_t(0), stream, 1, x, pos, rows, period, out),
                    "add_pos_rows launch");
}

int mil_head(const float* enc_out, int B, int nsnip, int ncls, const float* wa, const float* ba, const float* wb,
             const float* bb, const float* wc, const float* bc, const float* wf, const float* bf, float* reps_out,
             float* logits, float* attn_out, cudaStream_t stream) {
  if (B == 0) return kOk;
  if (!enc_out || !wa || !ba || !wb || !bb || !wc || !bc || !wf || !bf || !reps_out || !logits || !attn_out || B < 0 ||
      nsnip <= 0 || nsnip > 64 || ncls <= 0 || ncls > kMilMaxClasses) {
    set_last_error("mil_head: bad arguments (need 1 <= nsnippets <= 64, 1 <= classes <= 3)");
    return kErrInvalidArg;
  }
  const size_t smem = (size_t(nsnip) * (D + kMilGate) + kMilMaxClasses * size_t(nsnip) + D) * sizeof(float);
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(mil_head_kernel), int(smem), "mil_head")) return rc;
  LaunchScope ls(kClsMisc, stream, double(B) * nsnip * D * 8);
  return check_cuda(launch_pdl(mil_head_kernel, dim3(B), dim3(256), smem, stream, 1, enc_out, nsnip, ncls, wa, ba, wb, bb, wc, bc,
                               wf, bf, reps_out, logits, attn_out, B),
                    "mil_head launch");
}

int prototype_score(const float* reps, const float* protos, int B, int P, int Dd, float* probs, float* sims,
                    int32_t* pred, cudaStream_t stream) {
  if (B == 0) return kOk;
  if (!reps || !protos || !probs || B < 0 || P <= 0 || P > 64 || Dd <= 0) {
    set_last_error("prototype_score: bad arguments (need 1 <= P <= 64)");
    return kErrInvalidArg;
  }
  LaunchScope ls(kClsMisc, stream, double(B) * (Dd * 4 + P * 8));
  return check_cuda(launch_pdl(prototype_score_kernel, dim3((B + 3) / 4), dim3(128), size_t(0), stream, 1, reps, protos, B, P, Dd, probs, sims, pred),
                    "prototype_score launch");
}

}  // namespace sais

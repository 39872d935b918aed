// Multi-head attention core of the SAIS temporal encoder (nn.MultiheadAttention, 4 heads x 96,
// reached through prepare_model.py:213 with the README-patched layer that also returns the
// head-averaged attention weights).  Sequences are short (S = nframes + 1, typically 10..65) and
// packed back to back, so this is a latency/L2-bound SIMT kernel rather than a tensor-core one:
// one CTA per sequence keeps that sequence's K and V (all 4 heads) in shared memory, each warp owns
// query rows, runs the 4 heads in turn (fp32 logits, -inf on padded keys, exact softmax) and sums the
// per-head probabilities in shared memory so the head-mean map [S,S] is written once, coalesced.
#include "common.cuh"
#include "kernels.h"

namespace sais {

namespace {

constexpr int E = 384;
constexpr int H = 4;
constexpr int HD = 96;
constexpr int LD = 3 * E;      // qkv row pitch
constexpr int KV_PITCH = 392;  // smem row pitch (bf16 elements): 784 B, conflict-free 16-byte row reads
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxSmemS = 128;

__host__ __device__ inline int round_up32(int s) { return (s + 31) & ~31; }

inline size_t smem_bytes(int max_S, bool smem_kv) {
  size_t b = size_t(kWarps) * (HD * 4 + 2 * size_t(round_up32(max_S)) * 4);
  if (smem_kv) b += 2 * size_t(max_S) * KV_PITCH * 2;
  return b;
}

template <bool kSmemKV>
__global__ void __launch_bounds__(kThreads)
temporal_attention_kernel(const __nv_bfloat16* __restrict__ qkv, const int32_t* __restrict__ seq_offsets,
                          const uint8_t* __restrict__ key_pad, const int64_t* __restrict__ attn_offsets,
                          int max_S, __nv_bfloat16* __restrict__ out, float* __restrict__ attn_out) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int i = blockIdx.x;
  const int t0 = seq_offsets[i];
  const int S = seq_offsets[i + 1] - t0;
  if (S <= 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Sp = round_up32(max_S);

  float* qs = reinterpret_cast<float*>(smem) + warp * HD;                          // [kWarps][96]
  float* sc = reinterpret_cast<float*>(smem) + kWarps * HD + warp * 2 * Sp;        // [kWarps][Sp]
  float* acc = sc + Sp;                                                            // [kWarps][Sp]
  __nv_bfloat16* kv_s = reinterpret_cast<__nv_bfloat16*>(smem + size_t(kWarps) * (HD * 4 + 2 * size_t(Sp) * 4));

  const __nv_bfloat16* Kb;
  const __nv_bfloat16* Vb;
  int pitch;
  if constexpr (kSmemKV) {
    __nv_bfloat16* Ks = kv_s;
    __nv_bfloat16* Vs = kv_s + size_t(max_S) * KV_PITCH;
    for (int e = threadIdx.x; e < S * 96; e += kThreads) {  // 96 16-byte chunks per token (K then V)
      const int j = e / 96, c = e % 96;
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(qkv + int64_t(t0 + j) * LD + E) + c);
      __nv_bfloat16* dst = (c < 48) ? (Ks + j * KV_PITCH + c * 8) : (Vs + j * KV_PITCH + (c - 48) * 8);
      *reinterpret_cast<uint4*>(dst) = v;
    }
    __syncthreads();
    Kb = Ks; Vb = Vs; pitch = KV_PITCH;
  } else {
    Kb = qkv + int64_t(t0) * LD + E;
    Vb = qkv + int64_t(t0) * LD + 2 * E;
    pitch = LD;
  }

  const float scale = 0.10206207261596575f;  // 96^-0.5
  const int64_t aoff = (attn_out != nullptr && attn_offsets != nullptr) ? attn_offsets[i] : -1;
  const uint8_t* pad = key_pad ? key_pad + t0 : nullptr;

  for (int r = warp; r < S; r += kWarps) {
    for (int j = lane; j < S; j += 32) acc[j] = 0.f;
    for (int h = 0; h < H; ++h) {
      // q row of this head -> fp32 in smem (broadcast reads below)
      const __nv_bfloat16* qp = qkv + int64_t(t0 + r) * LD + h * HD;
#pragma unroll
      for (int c = 0; c < 3; ++c) qs[lane + 32 * c] = __bfloat162float(qp[lane + 32 * c]);
      __syncwarp();
      // logits
      float lmax = -INFINITY;
      for (int j = lane; j < S; j += 32) {
        const uint4* kp = reinterpret_cast<const uint4*>(Kb + int64_t(j) * pitch + h * HD);
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < 12; ++c) {
          const uint4 kk = kp[c];
          const float4 qa = *reinterpret_cast<const float4*>(qs + c * 8);
          const float4 qb = *reinterpret_cast<const float4*>(qs + c * 8 + 4);
          dot = fmaf(bf16_lo(kk.x), qa.x, dot); dot = fmaf(bf16_hi(kk.x), qa.y, dot);
          dot = fmaf(bf16_lo(kk.y), qa.z, dot); dot = fmaf(bf16_hi(kk.y), qa.w, dot);
          dot = fmaf(bf16_lo(kk.z), qb.x, dot); dot = fmaf(bf16_hi(kk.z), qb.y, dot);
          dot = fmaf(bf16_lo(kk.w), qb.z, dot); dot = fmaf(bf16_hi(kk.w), qb.w, dot);
        }
        float s = dot * scale;
        if (pad && pad[j]) s = -INFINITY;
        sc[j] = s;
        lmax = fmaxf(lmax, s);
      }
      const float m = warp_max(lmax);
      float lsum = 0.f;
      for (int j = lane; j < S; j += 32) {
        const float p = expf(sc[j] - m);
        sc[j] = p;
        lsum += p;
      }
      const float inv = 1.0f / warp_sum(lsum);
      for (int j = lane; j < S; j += 32) {
        const float p = sc[j] * inv;
        sc[j] = p;
        acc[j] += p;
      }
      __syncwarp();
      // O = P V for this head: lane owns dims lane, lane+32, lane+64
      float o0 = 0.f, o1 = 0.f, o2 = 0.f;
      const __nv_bfloat16* vp = Vb + h * HD + lane;
      for (int j = 0; j < S; ++j) {
        const float p = sc[j];
        const __nv_bfloat16* v = vp + int64_t(j) * pitch;
        o0 = fmaf(p, __bfloat162float(v[0]), o0);
        o1 = fmaf(p, __bfloat162float(v[32]), o1);
        o2 = fmaf(p, __bfloat162float(v[64]), o2);
      }
      __nv_bfloat16* op = out + int64_t(t0 + r) * E + h * HD + lane;
      op[0] = __float2bfloat16(o0);
      op[32] = __float2bfloat16(o1);
      op[64] = __float2bfloat16(o2);
      __syncwarp();
    }
    if (aoff >= 0) {
      float* ap = attn_out + aoff + int64_t(r) * S;
      for (int j = lane; j < S; j += 32) ap[j] = acc[j] * (1.0f / H);
    }
    __syncwarp();
  }
}

}  // namespace

int temporal_attention(const sais_bf16* qkv, const int32_t* seq_offsets, const uint8_t* key_pad,
                       const int64_t* attn_offsets, int nseq, int max_S, sais_bf16* out, float* attn_out,
                       cudaStream_t stream) {
  if (nseq == 0) return kOk;
  if (!qkv || !seq_offsets || !out || nseq < 0 || max_S <= 0) {
    set_last_error("temporal_attention: bad arguments");
    return kErrInvalidArg;
  }
  const bool smem_kv = max_S <= kMaxSmemS;
  const size_t sm = smem_bytes(max_S, smem_kv);
  if (sm > 227 * 1024) {
    set_last_error("temporal_attention: max_S=%d needs %zu bytes of shared memory", max_S, sm);
    return kErrShape;
  }
  static size_t attr_smem[2] = {0, 0};
  const void* fn = smem_kv ? reinterpret_cast<const void*>(temporal_attention_kernel<true>)
                           : reinterpret_cast<const void*>(temporal_attention_kernel<false>);
  if (sm > attr_smem[smem_kv]) {
    const size_t want = sm > 48 * 1024 ? 227 * 1024 : 48 * 1024;
    int rc = check_cuda(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(want)),
                        "cudaFuncSetAttribute(temporal_attention)");
    if (rc) return rc;
    attr_smem[smem_kv] = want;
  }
  const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(qkv);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  LaunchScope ls(kClsTemporalAttn, stream, 0.0);
  if (smem_kv)
    temporal_attention_kernel<true><<<nseq, kThreads, sm, stream>>>(q, seq_offsets, key_pad, attn_offsets, max_S, o,
                                                                    attn_out);
  else
    temporal_attention_kernel<false><<<nseq, kThreads, sm, stream>>>(q, seq_offsets, key_pad, attn_offsets, max_S,
                                                                     o, attn_out);
  return check_cuda(cudaGetLastError(), "temporal_attention launch");
}

}  // namespace sais

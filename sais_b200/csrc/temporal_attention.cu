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
  return launch_seq_attention<4, 96>(qkv, seq_offsets, key_pad, attn_offsets, nseq, max_S,
                                     0.10206207261596575f /* 96^-0.5 */, out_split, attn_out, nullptr,
                                     kClsTemporalAttn, stream);
}

int vit_attention_precise(const float* qkv, const int32_t* seq_offsets, int B, sais_bf16* out_split, float* probs,
                          cudaStream_t stream) {
  return launch_seq_attention<6, 64>(qkv, seq_offsets, nullptr, nullptr, B, 197, 0.125f, out_split, nullptr, probs,
                                     kClsVitAttn, stream);
}

}  // namespace sais

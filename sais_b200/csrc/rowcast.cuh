// bf16 copy + LayerNorm row statistics of fp32 rows of 384 — the operand of a LayerNorm-folded consumer GEMM
// (gemm_tcgen05.cu, "LayerNorm folding"):   xb[row] = bf16(x[row]),   stats[row] = {sum, sum of squares, 0 x 6}.
// One warp per row; lane l owns columns i * 128 + 4 l + {0..3}, i = 0..2.  Shared by the stand-alone kernel
// (elementwise.cu: rowstats_cast384_kernel) and the cast warps of the fused MLP (mlp_fused.cu), so the two produce
// bit-identical copies and statistics: same per-lane accumulation order, same xor-shuffle tree.
#pragma once

#include "common.cuh"

namespace sais {

constexpr int kRowCastD = 384;

template <bool kBypassL1>
__device__ __forceinline__ float4 rowcast_load(const float* p) {
  if (kBypassL1) return __ldcg(reinterpret_cast<const float4*>(p));
  return *reinterpret_cast<const float4*>(p);
}

// NR rows per call — row0, row0 + stride, ... (the first `nvalid` of them exist, 1 <= nvalid <= NR, warp-uniform): all
// 3 * NR 16-byte loads of a lane are in flight before the first reduction.
template <int NR, bool kBypassL1>
__device__ __forceinline__ void rowcast_rows(const float* __restrict__ x, __nv_bfloat16* __restrict__ xb,
                                             float* __restrict__ stats, int64_t row0, int stride, int nvalid, int lane) {
  float4 v[NR][3];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const int64_t row = row0 + (r < nvalid ? r * stride : 0);  // (a missing slot re-reads row0: no divergence, no OOB)
#pragma unroll
    for (int i = 0; i < 3; ++i) v[r][i] = rowcast_load<kBypassL1>(x + row * kRowCastD + i * 128 + lane * 4);
  }
  float s[NR], q[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    s[r] = 0.f;
    q[r] = 0.f;
    if (r < nvalid) {  // warp-uniform
      const int64_t row = row0 + r * stride;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float4 w = v[r][i];
        s[r] += (w.x + w.y) + (w.z + w.w);
        q[r] = fmaf(w.x, w.x, fmaf(w.y, w.y, fmaf(w.z, w.z, fmaf(w.w, w.w, q[r]))));
        uint2 o;
        o.x = pack_bf16x2(w.x, w.y);
        o.y = pack_bf16x2(w.z, w.w);
        *reinterpret_cast<uint2*>(xb + row * kRowCastD + i * 128 + lane * 4) = o;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    s[r] = warp_sum(s[r]);
    q[r] = warp_sum(q[r]);
  }
  // lanes 2r, 2r + 1 write the two 16-byte halves of row r's 32-byte statistics record
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    if (r < nvalid && (lane >> 1) == r) {
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if ((lane & 1) == 0) {
        o.x = s[r];
        o.y = q[r];
      }
      *reinterpret_cast<float4*>(stats + (row0 + r * stride) * 8 + (lane & 1) * 4) = o;
    }
  }
}

}  // namespace sais

// Internal launcher declarations shared by the .cu files (the public surface is include/sais_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sais_b200.h"

namespace sais {

// kernel classes for the launch counter / optional CUDA-event profiler (sais_profile_*)
enum LaunchClass : int {
  kClsGemm = 0, kClsVitAttn, kClsLayerNorm, kClsPatchify, kClsTemporalAttn, kClsMisc,
  kClsGemmSplit,  // split-precision (3-pass) GEMM launches: the temporal head and the fp32-equivalent ViT mode
  kClsMlpFused,   // mlp_fused_kernel (the dominant kernel of a step; work = 4 * rows * 384 * 1536 flop)
  kClsGemmQkv,    // bf16 qkv GEMM of the ViT blocks (M >= 4096, N = 1152, K = 384)
  kClsGemmProj,   // fp32-output proj GEMM + residual of the ViT blocks (M >= 4096, N = 384, K = 384)
  kNumClasses
};

// RAII around one kernel launch: counts it and, when profiling is on, brackets it with CUDA events on `stream`
// and books `work` (flops for GEMM/attention classes, algorithmic bytes for the memory-bound ones).
struct LaunchScope {
  LaunchScope(int cls, cudaStream_t stream, double work = 0.0);
  ~LaunchScope();
  int cls_;
  cudaStream_t stream_;
  int slot_;
};

// gemm_tcgen05.cu
int gemm_bias_act(const SaisGemmArgs& a, cudaStream_t stream, int force_block_n = 0);
// Scheduling hint for the row-tiled kernels (GEMM m-tiles, attention items): 1 = walk the rows from the last tile to the
// first.  sais_vit_forward alternates it from kernel to kernel ("snake" order): every kernel then starts on the rows its
// producer wrote LAST, which are the ones still resident in the 126 MB L2 — walking in the producer's own order evicts
// each line just before it is needed once a tensor is larger than the cache (qkv 116 MB, MLP hidden 155 MB at batch
// 256).  Results do not depend on it.
extern thread_local int g_tile_reverse;
int pick_block_n(int64_t M, int64_t N);

// mlp_fused.cu: x += fc2(GELU(fc1(xn) + b1)) + b2 over bf16 xn [rows,384], fp32 x [rows,384] (in place)
// xb_out / stats_out (both or neither; may alias xn / ln_stats): bf16 copy + row statistics of the UPDATED stream, written by
// the kernel's cast warps — what rowstats_cast would produce from x afterwards, bit for bit
int vit_mlp_fused(const sais_bf16* xn, const sais_bf16* w1, const float* b1, const sais_bf16* w2, const float* b2,
                  float* x, int64_t rows, cudaStream_t stream, const float* ln_stats = nullptr,
                  const float* ln_colsum = nullptr, float ln_eps = 0.0f, sais_bf16* xb_out = nullptr,
                  float* stats_out = nullptr);

// elementwise.cu
// out_plus (optional) = LayerNorm output + plus_vec[384]: the pre-loaded accumulator of an accumulate-mode GEMM
// fan (optional, with out_f32): the rows are also stored to the other GPUs' mappings of out_f32 (+ fan_offset elements):
// SaisFanout in include/sais_b200.h
int layernorm(const float* x, int64_t in_pitch, const float* gamma, const float* beta, float eps, int64_t rows,
              float* out_f32, sais_bf16* out_bf16, cudaStream_t stream, int split = 0, float* out_plus = nullptr,
              const float* plus_vec = nullptr, const SaisFanout* fan = nullptr, int64_t fan_offset = 0);
int rowstats_cast(const float* x, int64_t rows, sais_bf16* xb, float* stats, cudaStream_t stream);
int normalize_patchify_u8(const uint8_t* frames, int B, const float* mean3, const float* std3, sais_bf16* patches,
                          cudaStream_t stream, int split = 0);
int patchify_f32(const float* frames, int B, sais_bf16* patches, cudaStream_t stream, int split = 0);
int fill_offsets(int32_t* offs, int n, int stride, cudaStream_t stream);
int write_cls_rows(const float* cls_pos0, int B, float* x, cudaStream_t stream);
int temporal_prep(const float* x_frames, const int32_t* seq_offsets, int nseq, int total_tokens,
                  const float* frame_cls, const float* frame_pos, int n_pos, float* tok_f32, sais_bf16* tok_bf16,
                  cudaStream_t stream, float* tok_plus = nullptr, const float* plus_vec = nullptr);
int gather_cls_relu(const float* tok_f32, const int32_t* seq_offsets, int nseq, float* out_cls, cudaStream_t stream);
int clip_head(const float* cls_a, const float* cls_b, int B, int nsnip, const float* lin_w, const float* lin_b,
              float* out, cudaStream_t stream);
int prototype_score(const float* reps, const float* protos, int B, int P, int D, float* probs, float* sims,
                    int32_t* pred, cudaStream_t stream);
int add_pos_rows(const float* x, const float* pos, int64_t rows, int period, float* out, cudaStream_t stream);
int mil_head(const float* enc_out, int B, int nsnip, int ncls, const float* wa, const float* ba, const float* wb,
             const float* bb, const float* wc, const float* bc, const float* wf, const float* bf, float* reps_out,
             float* logits, float* attn_out, cudaStream_t stream);

// vit_attention.cu
int vit_attention(const sais_bf16* qkv, int B, sais_bf16* out, float* probs, cudaStream_t stream);

// last block, CLS query only: qkv bf16 [B*197,1152] -> out_cls bf16 [B,384]
int vit_cls_attention(const sais_bf16* qkv, int B, sais_bf16* out_cls, cudaStream_t stream);

// vit_attention_tc.cu (tcgen05 kernel; probs: optional fp32 [B,6,197,197] softmax probabilities)
int vit_attention_tc(const sais_bf16* qkv, int B, sais_bf16* out, float* probs, cudaStream_t stream);

// temporal_attention.cu
// fp32 qkv in, bf16 [hi | lo] (row pitch 2*384) out
int temporal_attention(const float* qkv, const int32_t* seq_offsets, const uint8_t* key_pad,
                       const int64_t* attn_offsets, int nseq, int max_S, sais_bf16* out_split, float* attn_out,
                       cudaStream_t stream);
int vit_attention_precise(const float* qkv, const int32_t* seq_offsets, int B, sais_bf16* out_split, float* probs,
                          cudaStream_t stream);

}  // namespace sais

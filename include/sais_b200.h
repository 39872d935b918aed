/*
 * sais_b200 — C ABI of the B200 (sm_100a) implementation of SAIS's inference hot path.
 *
 * The reference (danikiyasseh/SAIS) is pure Python and has no FFI; the interfaces each entry point
 * replaces are therefore Python call sites, cited per function below (paths relative to the reference
 * repo root).  All pointers are DEVICE pointers unless a parameter says "host".  Every function
 * returns 0 on success or a negative error code (sais_last_error() gives the text), never throws,
 * never allocates device memory (workspaces are caller-provided) and enqueues its work on `stream`
 * (a cudaStream_t passed as void*).  There is no CPU fallback: without an sm_100 device the calls fail.
 */
#ifndef SAIS_B200_H_
#define SAIS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SAIS_OK 0
#define SAIS_ERR_INVALID_ARG (-1)
#define SAIS_ERR_SHAPE (-2)
#define SAIS_ERR_CUDA (-3)
#define SAIS_ERR_WORKSPACE (-4)
#define SAIS_ERR_DRIVER (-5)

#define SAIS_ACT_NONE 0
#define SAIS_ACT_GELU_ERF 1 /* nn.GELU() default (erf form), vision_transformer.py:50 */
#define SAIS_ACT_RELU 2     /* nn.TransformerEncoderLayer default activation */

#define SAIS_VIT_DIM 384
#define SAIS_VIT_DEPTH 12
#define SAIS_VIT_HEADS 6
#define SAIS_VIT_TOKENS 197
#define SAIS_VIT_PATCHES 196
#define SAIS_VIT_PATCH_K 768
#define SAIS_VIT_HIDDEN 1536
#define SAIS_TMP_LAYERS 4
#define SAIS_TMP_HEADS 4
#define SAIS_TMP_FF 2048
#define SAIS_TMP_OUT 256

typedef void* sais_stream_t; /* cudaStream_t */
typedef uint16_t sais_bf16;  /* raw bfloat16 bits */

int sais_version(void);
const char* sais_last_error(void);
/* number of kernels this library has launched since load (for bench.py's gpu_launches) */
int64_t sais_launch_count(void);

/* Which kernels run the MLP of a ViT block on the bf16 path.  0 (default): the fused MLP kernel for every batch size —
 * a frame's embedding is then independent of the batch it arrives in, bit for bit.  1: the fc1 / fc2 GEMM pair below 48
 * frames per chunk (lower latency for single clips: 10 + 10 frames 1.11 -> 0.97 ms; rounding points differ from the fused
 * kernel's, tolerances unchanged).  2: the GEMM pair always.  Returns the previous policy (>= 0) or a negative error code.
 * Process-wide; SAIS_MLP_FOLD=auto / 0 set the initial value. */
int sais_set_mlp_policy(int32_t policy);

/* Spatial partitioning of the GPU between two streams of this library's kernels.  Every persistent kernel sizes its grid
 * from the SM count; with a limit set (an even number below the device's count, 0 = no limit) the kernels LAUNCHED while it is
 * in force occupy at most that many SMs and leave the rest to kernels launched on another stream without it.  Used to run
 * the latency-bound temporal head of batch i (~30 small dependent kernels, 0.3 ms on an otherwise idle GPU) on 4 reserved
 * SMs concurrently with the ViT of batch i + 1 instead of after it.  Process-wide, read at launch time; returns the
 * previous limit (>= 0) or a negative error code. */
int sais_set_sm_limit(int32_t n_sms);

/* Measurement aid: a one-thread kernel enqueued on `stream` that records {globaltimer ns, SM cycle counter} before and
 * after a chain of spin_iters dependent FMAs into out4 (device int64[4]): (out4[3]-out4[1]) / (out4[2]-out4[0]) is the
 * effective SM clock in GHz at that point of the stream — what the kernels around it actually ran at (NVML sampling
 * is too coarse to see power-cap clock dips inside a millisecond-scale step).  Not counted by sais_launch_count. */
int sais_clock_probe(int64_t* out4, int32_t spin_iters, sais_stream_t stream);

/* Optional per-kernel-class CUDA-event profiler (bench.py's roofline leg).  Between begin and end every
 * launch is bracketed by events on its stream; end synchronises the device and returns, per class
 * (0 other bf16 GEMMs, 1 ViT attention, 2 LayerNorm / row statistics, 3 patchify, 4 temporal attention, 5 misc,
 * 6 split-precision GEMM of the temporal head / fp32 mode, 7 fused ViT MLP kernel, 8 ViT qkv GEMM, 9 ViT proj GEMM):
 * summed milliseconds, summed algorithmic work (flops for classes 0, 1, 6, 7, 8, 9; bytes otherwise) and launch counts.
 * All arrays are HOST. */
#define SAIS_NUM_KERNEL_CLASSES 10
void sais_profile_begin(void);
int sais_profile_end(double* ms_per_class, double* work_per_class, int64_t* launches_per_class, int32_t n_classes);

/* ---------------------------------------------------------------------------------------------
 * GEMM  out = act(A · Wᵀ + bias) [+ residual]           (tcgen05 / TMEM / TMA kernel)
 * replaces nn.Linear / nn.Conv2d-as-GEMM call sites: vision_transformer.py:60-63 (fc1, fc2),
 * :82 (qkv), :90 (proj), :126-130 (patch embed); torch MultiheadAttention in/out projections and
 * TransformerEncoderLayer.linear1/linear2 reached through prepare_model.py:213.
 * A: bf16 [M,K] (row pitch lda), W: bf16 [N,K] (nn.Linear layout, row pitch ldw); fp32 accumulate.
 * N must be a multiple of 128, K a multiple of 64.  out_f32 and/or out_bf16 may be given.
 * Patch-embed row remap: if remap_group > 0, GEMM row r = g*remap_group + p is written to output
 * row g*(remap_group+1) + 1 + p and row_add[p, :] (fp32 [remap_group, N]) is added (pos_embed).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const sais_bf16* a;
  const sais_bf16* w;
  const float* bias;     /* [N] or NULL */
  const float* residual; /* fp32 [M,N] pitch ldr or NULL; added after the activation */
  float* out_f32;        /* pitch ldo32, or NULL */
  sais_bf16* out_bf16;   /* pitch ldo16, or NULL */
  const float* row_add;  /* see remap_group */
  int64_t M, N, K;
  int64_t lda, ldw, ldr, ldo32, ldo16;
  int32_t act;
  int32_t remap_group;
  /* Split-precision ("3 x bf16", fp32-equivalent) mode.  split3: A and W each hold [hi | lo] bf16 halves
   * side by side (2K columns; x = hi + lo to ~2^-17) and the kernel accumulates hi*hi + lo*hi + hi*lo in
   * fp32.  split_out: out_bf16 has 2N columns and receives the result as [hi | lo] for the next split GEMM. */
  int32_t split3;
  int32_t split_out;
  /* LayerNorm folding (bf16 fast path).  LayerNorm is affine per row, so  LN(x) Wᵀ + b = rstd (x W'ᵀ - mean c) + d
   * with W' = gamma ∘ W, c_n = sum_k W'_nk, d_n = sum_k beta_k W_nk + b_n.
   * Consumer (qkv, fc1): a = bf16 copy of the UN-normalised rows, w = W', bias = d, ln_colsum = c and
   *   ln_stats_in = fp32 [M][4][2] partial (sum, sum of squares) of each input row; mean / rstd are applied in the
   *   epilogue (before the activation).  bf16 output only.
   * Producer (proj, fc2; N = 384, fp32 out + residual): additionally writes out2_bf16 = bf16(out) (pitch ldo2) and
   *   ln_stats_out = fp32 [M][4][2] partials of its output rows for the consumer that follows. */
  const float* ln_stats_in;
  const float* ln_colsum;
  float* ln_stats_out;
  sais_bf16* out2_bf16;
  int64_t ldo2;
  float ln_eps;
  /* Accumulate mode (0 = off).  k_slices >= 1: out_f32 += A · Wᵀ — the product is ADDED to what out_f32 already holds
   * (the caller pre-loads residual + bias there), computed in k_slices slices of K by independent CTAs whose partial
   * sums meet in L2 through TMA reduce-add stores.  For small-M GEMMs with a long K (the temporal FF2) this spreads the
   * operand stream over many SMs instead of 9.  bias, residual and act must be NULL / 0. */
  int32_t k_slices;
} SaisGemmArgs;
int sais_gemm_bias_act(const SaisGemmArgs* args, sais_stream_t stream);

/* Fused ViT MLP with residual:  x += fc2(GELU_erf(fc1(xn) + fc1_b)) + fc2_b   — Mlp.forward and the residual add of
 * Block.forward (vision_transformer.py:56-65,107) in ONE kernel: the [rows,1536] hidden activations stay in
 * shared / tensor memory.  xn bf16 [rows,384] (LayerNorm output), fc1_w bf16 [1536,384], fc2_w bf16 [384,1536],
 * biases fp32, x fp32 [rows,384] updated in place (the add happens in L2 through a TMA reduce store). */
int sais_vit_mlp(const sais_bf16* xn, const sais_bf16* fc1_w, const float* fc1_b, const sais_bf16* fc2_w,
                 const float* fc2_b, float* x, int64_t rows, sais_stream_t stream);
/* Same with norm2 (Block.norm2, vision_transformer.py:103,111) folded in, as in SaisGemmArgs.ln_stats_in: xb holds the RAW
 * bf16 copy of the residual stream, ln_stats its per-row (sum, sum of squares) partials [rows][4][2], fc1_wg = gamma-scaled
 * fc1 weights, fc1_c their column sums, fc1_d = fc1(beta) + fc1_b:   x += fc2(GELU_erf(rstd (xb fc1_wg^T - mean c) + d)) + fc2_b.
 * xb_out / stats_out (both or neither, may alias xb / ln_stats): the bf16 copy and row statistics of the UPDATED stream,
 * i.e. exactly what sais_rowstats_cast(x) would produce afterwards (bit for bit) — the operand of the next block's folded
 * qkv GEMM (Block.norm1, vision_transformer.py:99,108), written by the kernel's cast warps from L2 while the tensor pipe
 * works on the next row tile, which saves a separate pass over the residual stream. */
int sais_vit_mlp_ln(const sais_bf16* xb, const float* ln_stats, float ln_eps, const sais_bf16* fc1_wg, const float* fc1_c,
                    const float* fc1_d, const sais_bf16* fc2_w, const float* fc2_b, float* x, int64_t rows,
                    sais_bf16* xb_out, float* stats_out, sais_stream_t stream);

/* LayerNorm over the last dim (cols == 384) — nn.LayerNorm at vision_transformer.py:99,103,156
 * (eps 1e-6) and TransformerEncoderLayer.norm1/norm2 (eps 1e-5).  x: fp32, row pitch in_pitch
 * elements; writes fp32 [rows,384] and/or bf16 ([rows,384], or [rows,768] = [hi | lo] if split_out). */
int sais_layernorm(const float* x, int64_t in_pitch, const float* gamma, const float* beta, float eps,
                   int64_t rows, int32_t cols, float* out_f32, sais_bf16* out_bf16, int32_t split_out,
                   sais_stream_t stream);

/* Entry of the LayerNorm-folded path: xb = bf16(x) and stats[row] = {sum, sum of squares, 0,0,0,0,0,0} for fp32 rows
 * of 384 (x: [rows,384]; xb: bf16 [rows,384]; stats: fp32 [rows,8]). */
int sais_rowstats_cast(const float* x, int64_t rows, sais_bf16* xb, float* stats, sais_stream_t stream);

/* JPEG front-end (replaces the host-side PIL decode of dino-main/main_dino.py:295-301,313 /
 * extract_representations.py:178 for RGB frames and for the `flows_%08d.jpg` optical-flow frames of :246-261).
 *   sais_jpeg_info: HOST helper, parses the frame header of one JPEG stream: hw2_host = {height, width}.  No CUDA context.
 *   sais_jpeg_decode_batch: n same-sized JPEG streams (HOST pointers / lengths) -> out_device u8 [n,H,W,3] interleaved RGB,
 *   one batched nvJPEG decode enqueued on `stream` (entropy decode + IDCT are nvJPEG's; results can differ from libjpeg's
 *   by a few levels per pixel, see tests/test_frames.py for the measured bound). */
int sais_jpeg_info(const uint8_t* data, size_t len, int32_t* hw2_host);
int sais_jpeg_decode_batch(const uint8_t* const* data_host, const size_t* lengths_host, int32_t n, int32_t H, int32_t W,
                           uint8_t* out_device, sais_stream_t stream);
/* Which nvJPEG decoder served the last sais_jpeg_decode_batch call on the current device: 1 = the GPU's fixed-function JPEG
 * engines (NVJPEG_BACKEND_HARDWARE), 2 = the batched CUDA decoder (NVJPEG_BACKEND_GPU_HYBRID / HYBRID), 3 = the threaded decoder (default: the frames of a call dealt
 * to T host threads, one nvJPEG state + CUDA stream each), 0 = none yet. */
int sais_jpeg_last_backend(void);

/* Frame front-end: centre crop + Pillow-exact antialiased bilinear resize to 224 x 224.
 * Replaces, for decoded uint8 frames, `transforms.CenterCrop((height_frac*height, width_frac*width))`
 * (dino-main/main_dino.py:298-301; fractions from getCropDims :317-322) and `transforms.Resize((224,224))`
 * (extract_representations.py:158-162) = PIL Image.resize(BILINEAR): Pillow's two-pass 8-bit fixed-point triangle
 * filter (libImaging/Resample.c), horizontal pass first.  Integer arithmetic: results are bit-identical to Pillow.
 *   sais_center_crop_box: HOST helper, box4_host = {top, left, crop_h, crop_w} with torchvision's / PIL's rounding.
 *   sais_resize_table_ints / sais_resize_build_table: HOST helpers; the coefficient table for one source side
 *     (crop_w for the horizontal pass, crop_h for the vertical one): int32 [224][2] (first index, taps) followed by
 *     [ksize][224] coefficients.  The caller copies the tables to the device (they depend on the geometry only).
 *   sais_crop_resize_u8: frames u8 [N,H,W,3] -> out u8 [N,224,224,3]; tmp = u8 [N,crop_h,224,3] workspace. */
int sais_center_crop_box(int32_t height, int32_t width, double height_frac, double width_frac, int32_t* box4_host);
int64_t sais_resize_table_ints(int32_t in_size);
int sais_resize_build_table(int32_t in_size, int32_t* table_host);
int sais_crop_resize_u8(const uint8_t* frames, int32_t N, int32_t H, int32_t W, int32_t top, int32_t left,
                        int32_t crop_h, int32_t crop_w, const int32_t* table_h, const int32_t* table_v, uint8_t* tmp,
                        uint8_t* out, sais_stream_t stream);

/* Frame normalisation + patch layout.  u8 variant replaces ToTensor+Normalize
 * (extract_representations.py:158-162): frames u8 [B,224,224,3] -> patches bf16 [B*196,768],
 * k = c*256 + ky*16 + kx, value = (u8/255 - mean[c]) / std[c].  f32 variant takes the already
 * normalised fp32 [B,3,224,224] tensor the reference model is called with (:370).  With split_out the
 * patch matrix is bf16 [B*196,1536] = [hi | lo] halves for the split-precision patch-embed GEMM. */
int sais_normalize_patchify_u8(const uint8_t* frames, int32_t B, const float* mean3_host, const float* std3_host,
                               sais_bf16* patches, int32_t split_out, sais_stream_t stream);
int sais_patchify_f32(const float* frames_chw, int32_t B, sais_bf16* patches, int32_t split_out,
                      sais_stream_t stream);

/* ViT self-attention for one block: qkv bf16 [B*197,1152] (q|k|v, head-major inside each third,
 * vision_transformer.py:82-89) -> out bf16 [B*197,384].  probs (optional) receives the softmax
 * probabilities fp32 [B,6,197,197] (get_last_selfattention, :216-223). */
int sais_vit_attention(const sais_bf16* qkv, int32_t B, sais_bf16* out, float* probs, sais_stream_t stream);

/* CLS-query attention of one block: out_cls bf16 [B,384] = attention output of token 0 of every frame only.  Used for
 * the LAST block, whose other rows VisionTransformer.forward discards (vision_transformer.py:213-214). */
int sais_vit_cls_attention(const sais_bf16* qkv, int32_t B, sais_bf16* out_cls, sais_stream_t stream);

/* Whole ViT-S/16 backbone. */
typedef struct {
  const float* ln1_w; const float* ln1_b;
  const sais_bf16* qkv_w; const float* qkv_b;   /* [1152,384], [1152] */
  const sais_bf16* proj_w; const float* proj_b; /* [384,384], [384] */
  const float* ln2_w; const float* ln2_b;
  const sais_bf16* fc1_w; const float* fc1_b;   /* [1536,384], [1536] */
  const sais_bf16* fc2_w; const float* fc2_b;   /* [384,1536], [384] */
  /* LayerNorm-folded operands of the bf16 fast path (see SaisGemmArgs): W' = gamma ∘ W as bf16, c_n = sum_k W'_nk,
   * d_n = sum_k beta_k W_nk + b_n.  All NULL -> the forward keeps separate LayerNorm kernels. */
  const sais_bf16* qkv_wg; const float* qkv_c; const float* qkv_d; /* norm1 folded into qkv */
  const sais_bf16* fc1_wg; const float* fc1_c; const float* fc1_d; /* norm2 folded into fc1 */
} SaisVitBlockWeights;
typedef struct {
  const sais_bf16* patch_w; /* [384,768] = patch_embed.proj.weight.view(384,-1) */
  const float* patch_b;     /* [384] */
  const float* cls_pos0;    /* [384] = cls_token + pos_embed[0] */
  const float* pos_patch;   /* [196,384] = pos_embed[1:] + patch_b is NOT folded; plain pos_embed[1:] */
  SaisVitBlockWeights blocks[SAIS_VIT_DEPTH];
  const float* norm_w; const float* norm_b;
} SaisVitWeights;

#define SAIS_INPUT_F32_CHW 0 /* normalised fp32 [B,3,224,224] */
#define SAIS_INPUT_U8_HWC 1  /* raw u8 [B,224,224,3], normalised with ImageNet mean/std */
size_t sais_vit_workspace_bytes(int32_t chunk_frames, int32_t precise);
/* VisionTransformer.forward (vision_transformer.py:209-214): out_cls fp32 [B,384].
 * If out_probs != NULL also writes block 12's attention probabilities fp32 [B,6,197,197]
 * (get_last_selfattention).  Frames are processed in chunks of `chunk_frames` (workspace sized for it).
 * out_tokens (optional) receives the final-LayerNorm'd tokens fp32 [B,197,384] (get_intermediate_layers n=1).
 * precise = 0: bf16 operands, fp32 accumulate/residual (the fast path; embeddings within 1e-2 of fp32).
 * precise = 1: fp32-equivalent mode — every bf16 weight matrix in w_host must then be packed as [hi | lo]
 * (twice the columns), GEMMs run split-precision (3 passes) and attention runs in exact fp32. */
int sais_vit_forward(const SaisVitWeights* w_host, const void* input, int32_t input_kind, int32_t B,
                     int32_t chunk_frames, int32_t precise, void* workspace, size_t workspace_bytes,
                     float* out_cls, float* out_probs, float* out_tokens, sais_stream_t stream);

/* get_intermediate_layers(x, n) (vision_transformer.py:225-233; eval_linear.py calls it with n = 4): the final-LayerNorm'd
 * tokens after each of the last n_last blocks, earliest first, fp32 [n_last,B,197,384]; out_cls as in sais_vit_forward. */
int sais_vit_forward_layers(const SaisVitWeights* w_host, const void* input, int32_t input_kind, int32_t B,
                            int32_t chunk_frames, int32_t precise, void* workspace, size_t workspace_bytes,
                            float* out_cls, int32_t n_last, float* out_tokens_stack, sais_stream_t stream);

/* The same forward with the path's ONE exchange step fused into its last kernel (when a video is sharded by
 * frame range, every rank needs all ranks' [n/R,384] embeddings ahead of the temporal head).  `out_cls` is this rank's
 * slice of a gather buffer that exists at the same offset on every GPU of the group (symmetric memory); `fan` tells the
 * final-LayerNorm kernel where that slice lives in the other GPUs' mappings, and the kernel stores every embedding row
 * there as it produces it — over NVLink, either with ONE multimem.st per 16 bytes to the NVSwitch multicast address
 * (`multicast`: the address of `out_cls` in the group's multicast mapping; the switch replicates the store to every GPU)
 * or with one plain store per peer (`peers[i]`: the address of `out_cls` in peer i's mapping, self excluded).  The local
 * slice is always written with a plain store as well, so this rank's own consumers need no cross-GPU ordering.  There
 * is no collective and no NCCL kernel on the data path; the caller orders consumers on other ranks behind a barrier that
 * follows this call in stream order (sais_b200.pipeline.PeerGatherer).  fan == NULL: exactly sais_vit_forward. */
#define SAIS_MAX_PEERS 15
typedef struct {
  void* multicast;
  void* peers[SAIS_MAX_PEERS];
  int32_t n_peers;
} SaisFanout;
int sais_vit_forward_fanout(const SaisVitWeights* w_host, const void* input, int32_t input_kind, int32_t B,
                            int32_t chunk_frames, int32_t precise, void* workspace, size_t workspace_bytes,
                            float* out_cls, float* out_probs, float* out_tokens, const SaisFanout* fan,
                            sais_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * SAIS temporal head (prepare_model.py:179-221 + README-patched nn.TransformerEncoder).
 * Sequences are PACKED: sequence i owns tokens [seq_offsets[i], seq_offsets[i+1]) with S_i = T_i+1
 * (token 0 = frame_cls, token t+1 = frame t + frame_pos_embeddings[t]); x_frames holds the T_i frame
 * embeddings of all sequences back to back (fp32 [total_tokens - nseq, 384]).
 * ------------------------------------------------------------------------------------------- */
/* The temporal head always runs split-precision: every bf16 matrix below is [N, 2K] = [hi | lo]. */
typedef struct {
  const sais_bf16* in_w; const float* in_b;   /* [1152,2*384], [1152] */
  const sais_bf16* out_w; const float* out_b; /* [384,2*384], [384] */
  const float* n1_w; const float* n1_b;
  const sais_bf16* ff1_w; const float* ff1_b; /* [2048,2*384], [2048] */
  const sais_bf16* ff2_w; const float* ff2_b; /* [384,2*2048], [384] */
  const float* n2_w; const float* n2_b;
} SaisTemporalLayerWeights;
typedef struct {
  const float* frame_cls; /* [384] */
  const float* frame_pos; /* [2000,384] stacked frame_pos_embeddings['0'..'1999'] */
  int32_t n_pos;
  SaisTemporalLayerWeights layers[SAIS_TMP_LAYERS];
} SaisTemporalWeights;

/* +pos-emb, prepend CLS (prepare_model.py:189-194): writes the token matrix as fp32 [total_tokens,384] and as
 * bf16 [total_tokens,768] = [hi | lo] halves (operand of the split-precision in-proj GEMM). */
int sais_temporal_prep(const float* x_frames, const int32_t* seq_offsets, int32_t nseq, int32_t total_tokens,
                       const float* frame_cls, const float* frame_pos, int32_t n_pos, float* tok_f32,
                       sais_bf16* tok_split, sais_stream_t stream);

/* Multi-head attention core of one temporal layer: qkv fp32 [total_tokens,1152] -> out bf16
 * [total_tokens,768] = [hi | lo]; 4 heads x 96, scale 96^-0.5, key_pad (u8 [total_tokens], 1 = padded key) adds -inf.
 * If attn_out != NULL, sequence i with attn_offsets[i] >= 0 gets its head-averaged probabilities
 * fp32 [S_i,S_i] written at attn_out + attn_offsets[i] (need_weights=True semantics). */
int sais_temporal_attention(const float* qkv, const int32_t* seq_offsets, const uint8_t* key_pad,
                            const int64_t* attn_offsets, int32_t nseq, int32_t max_S, sais_bf16* out_split,
                            float* attn_out, sais_stream_t stream);

size_t sais_temporal_workspace_bytes(int32_t total_tokens);
/* 4-layer post-norm encoder over packed sequences.  out_cls fp32 [nseq,384] = relu(out[CLS])
 * (prepare_model.py:215-220); out_tokens (optional) = raw encoder output fp32 [total_tokens,384];
 * attention map of the LAST layer as in sais_temporal_attention. */
int sais_temporal_forward(const SaisTemporalWeights* w_host, const float* x_frames, const int32_t* seq_offsets,
                          const uint8_t* key_pad, const int64_t* attn_offsets, int32_t nseq, int32_t total_tokens,
                          int32_t max_S, void* workspace, size_t workspace_bytes, float* out_cls, float* out_tokens,
                          float* attn_out, sais_stream_t stream);

/* Clip head (prepare_model.py:378-382,405-409): out[b] = W · relu(mean_s cls_a[b,s] + mean_s cls_b[b,s]) + bias.
 * cls_a/cls_b fp32 [B*nsnip,384] (cls_b may be NULL for single-modality); W fp32 [256,384]. */
int sais_clip_head(const float* cls_a, const float* cls_b, int32_t B, int32_t nsnip, const float* lin_w,
                   const float* lin_b, float* out, sais_stream_t stream);

/* MIL pathway of fullModel.forward (task='MIL', prepare_model.py:359-363; getClipReps :451-466, MIL_Head :468-488,
 * calcAttention / obtainVideoRep / obtainVideoScore :131-149).
 *   sais_add_pos_rows: out[r,:] = x[r,:] + pos[r % period,:] over fp32 rows of 384 (the clip positional embeddings).
 *   sais_mil_head: enc_out fp32 [B,nsnip,384] = output of the clip-level encoder (batch-major, before the ReLU);
 *     att_a / att_b: the gated-attention pair [256,384] + [256]; att_c_w [ncls,256], att_c_b [ncls]: attentionModules;
 *     final_w [ncls,384], final_b [ncls]: finalModules.  Writes reps_out [B,nsnip,384] = relu(enc_out), logits [B,ncls] and
 *     attn_out [ncls,B,nsnip] (softmax over the snippets).  1 <= nsnip <= 64, 1 <= ncls <= 3. */
int sais_add_pos_rows(const float* x, const float* pos, int64_t rows, int32_t period, float* out, sais_stream_t stream);
int sais_mil_head(const float* enc_out, int32_t B, int32_t nsnip, int32_t ncls, const float* att_a_w, const float* att_a_b,
                  const float* att_b_w, const float* att_b_b, const float* att_c_w, const float* att_c_b,
                  const float* final_w, const float* final_b, float* reps_out, float* logits, float* attn_out,
                  sais_stream_t stream);

/* Prototype scoring (prepare_miscellaneous.py:102-125, process_inference_results.py:76-91):
 * probs = softmax-free exp(cos)/sum exp(cos); pred = argmax.  reps fp32 [B,D], protos fp32 [P,D], P <= 64. */
int sais_prototype_score(const float* reps, const float* protos, int32_t B, int32_t P, int32_t D, float* probs,
                         float* sims, int32_t* pred, sais_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SAIS_B200_H_ */

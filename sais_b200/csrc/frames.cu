// Frame front-end ahead of the ViT: centre crop + Pillow-exact antialiased bilinear resize of
// decoded uint8 frames [N,H,W,3] to [N,224,224,3] — what `transforms.CenterCrop((0.8*h, 0.8*w))`
// (SAIS/scripts/dino-main/main_dino.py:298-301, getCropDims :317-322) followed by `transforms.Resize((224,224))`
// (SAIS/scripts/extract_representations.py:158-162) hands to ToTensor.  For a PIL image that resize is Pillow's
// two-pass separable triangle filter in 8-bit fixed point (libImaging/Resample.c: precompute_coeffs,
// normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc, ImagingResampleVertical_8bpc; PRECISION_BITS = 22): the
// horizontal pass runs first and rounds to uint8, the vertical pass reads that uint8 intermediate.  Everything is
// integer arithmetic, so the kernels are BIT-EXACT against Pillow (tests/test_frames.py).
//
// Both kernels are HBM-bound byte work (no tensor cores): the horizontal pass reads the crop window once with
// 16-byte loads into shared memory (one CTA = kRows crop rows; thread = (output pixel, channel) so a warp's gathers
// stay inside ~220 bytes of the row) and writes the 224-wide intermediate coalesced; the vertical pass reads the
// intermediate with 4-byte loads (thread = 4 consecutive bytes of an output row).  Coefficients are int32 tables
// built on the host exactly like Pillow does (double arithmetic, same operation order) and stored tap-major so that
// consecutive threads read consecutive words.
// Algorithmic bytes per frame: ch*cw*3 (crop window read) + 224*224*3 (result) [+ 2*ch*224*3 for the intermediate].
#include <cmath>
#include <cstdint>

#include "common.cuh"
#include "kernels.h"

namespace sais {

namespace {

constexpr int kOut = 224;            // output side
constexpr int kPrecisionBits = 22;   // Resample.c: 32 - 8 - 2
constexpr int kRows = 4;             // crop rows per CTA iteration of the horizontal pass
constexpr int kRowBytes = kOut * 3;  // one intermediate / output row

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= kPrecisionBits;
  return uint8_t(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// table layout (int32): [kOut][2] = (first source index, tap count), then [ksize][kOut] coefficients (tap-major)
__global__ void __launch_bounds__(kRowBytes) resize_horizontal_kernel(const uint8_t* __restrict__ frames, int64_t total_bytes,
                                                                      int N, int H, int W, int top, int left, int ch, int cw,
                                                                      const int32_t* __restrict__ table, int row_pitch,
                                                                      uint8_t* __restrict__ tmp) {
  pdl_wait();  // (PDL, common.cuh) no global access above this line
  extern __shared__ __align__(16) uint8_t rows_s[];  // [kRows][row_pitch]
  const int t = threadIdx.x;
  const int xx = t / 3, c = t - xx * 3;
  const int xmin = __ldg(table + 2 * xx), cnt = __ldg(table + 2 * xx + 1);
  const int32_t* kk = table + 2 * kOut + xx;
  const int groups_per_frame = (ch + kRows - 1) / kRows;
  const int64_t n_groups = int64_t(N) * groups_per_frame;
  const uintptr_t abs0 = reinterpret_cast<uintptr_t>(frames);
  for (int64_t g = blockIdx.x; g < n_groups; g += gridDim.x) {
    const int n = int(g / groups_per_frame);
    const int r0 = int(g - int64_t(n) * groups_per_frame) * kRows;
    const int nr = (ch - r0 < kRows) ? ch - r0 : kRows;
    __syncthreads();  // previous iteration's readers are done with rows_s
    int lead[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      lead[r] = 0;
      if (r < nr) {
        const int64_t off = ((int64_t(n) * H + top + r0 + r) * W + left) * 3;  // first byte of the crop row
        const uintptr_t a0 = (abs0 + off) & ~uintptr_t(15);                      // 16-byte aligned start (absolute address)
        lead[r] = int((abs0 + off) - a0);
        const int chunks = (lead[r] + cw * 3 + 15) >> 4;
        uint8_t* dst = rows_s + r * row_pitch;
        for (int j = t; j < chunks; j += kRowBytes) {
          const int64_t b = int64_t(a0 - abs0) + int64_t(j) * 16;  // offset of this chunk inside the frames buffer (may be < 0)
          if (b >= 0 && b + 16 <= total_bytes) {
            *reinterpret_cast<uint4*>(dst + j * 16) = __ldg(reinterpret_cast<const uint4*>(frames + b));
          } else {  // chunk straddles the ends of the buffer: stay inside it
            for (int i = 0; i < 16; ++i) dst[j * 16 + i] = (b + i >= 0 && b + i < total_bytes) ? frames[b + i] : uint8_t(0);
          }
        }
      }
    }
    __syncthreads();
    int acc[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) acc[r] = 1 << (kPrecisionBits - 1);
    const uint8_t* src = rows_s + xmin * 3 + c;
    for (int x = 0; x < cnt; ++x) {
      const int k = __ldg(kk + x * kOut);
#pragma unroll
      for (int r = 0; r < kRows; ++r) acc[r] += int(src[r * row_pitch + lead[r] + x * 3]) * k;
    }
    uint8_t* out = tmp + (int64_t(n) * ch + r0) * kRowBytes + t;
#pragma unroll
    for (int r = 0; r < kRows; ++r)
      if (r < nr) out[r * kRowBytes] = clip8(acc[r]);
  }
}

// tmp u8 [N,ch,672] -> out u8 [N,224,672]; thread = 4 consecutive bytes of one output row
__global__ void __launch_bounds__(kRowBytes / 4) resize_vertical_kernel(const uint8_t* __restrict__ tmp, int N, int ch,
                                                                        const int32_t* __restrict__ table,
                                                                        uint8_t* __restrict__ out) {
  pdl_wait();  // (PDL, common.cuh) no global access above this line
  const int t = threadIdx.x;
  for (int64_t row = blockIdx.x; row < int64_t(N) * kOut; row += gridDim.x) {
    const int n = int(row / kOut), yy = int(row - int64_t(n) * kOut);
    const int ymin = __ldg(table + 2 * yy), cnt = __ldg(table + 2 * yy + 1);
    const int32_t* kk = table + 2 * kOut + yy;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(tmp + (int64_t(n) * ch + ymin) * kRowBytes) + t;
    int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0, a3 = a0;
    for (int y = 0; y < cnt; ++y) {
      const int k = __ldg(kk + y * kOut);
      const uint32_t v = __ldg(src + y * (kRowBytes / 4));
      a0 += int(v & 0xff) * k;
      a1 += int((v >> 8) & 0xff) * k;
      a2 += int((v >> 16) & 0xff) * k;
      a3 += int(v >> 24) * k;
    }
    const uint32_t o = uint32_t(clip8(a0)) | (uint32_t(clip8(a1)) << 8) | (uint32_t(clip8(a2)) << 16) | (uint32_t(clip8(a3)) << 24);
    reinterpret_cast<uint32_t*>(out + row * kRowBytes)[t] = o;
  }
}

int resize_ksize(int in_size) {
  double filterscale = double(in_size) / kOut;
  if (filterscale < 1.0) filterscale = 1.0;
  return int(std::ceil(filterscale)) * 2 + 1;
}

}  // namespace

}  // namespace sais

using namespace sais;

extern "C" {

int sais_center_crop_box(int32_t height, int32_t width, double height_frac, double width_frac, int32_t* box4_host) {
  if (height <= 0 || width <= 0 || !(height_frac > 0.0 && height_frac <= 1.0) || !(width_frac > 0.0 && width_frac <= 1.0) ||
      !box4_host) {
    set_last_error("center_crop_box: bad arguments");
    return kErrInvalidArg;
  }
  // torchvision center_crop with a float size + PIL Image.crop: Python round() twice (round-half-even = nearbyint
  // in the default rounding mode)
  const double ch = height_frac * height, cw = width_frac * width;
  const double top = std::nearbyint((height - ch) / 2.0), left = std::nearbyint((width - cw) / 2.0);
  const int x0 = int(std::nearbyint(left)), y0 = int(std::nearbyint(top));
  const int x1 = int(std::nearbyint(left + cw)), y1 = int(std::nearbyint(top + ch));
  box4_host[0] = y0;
  box4_host[1] = x0;
  box4_host[2] = y1 - y0;
  box4_host[3] = x1 - x0;
  return kOk;
}

int64_t sais_resize_table_ints(int32_t in_size) {
  if (in_size <= 0) return 0;
  return int64_t(kOut) * 2 + int64_t(resize_ksize(in_size)) * kOut;
}

int sais_resize_build_table(int32_t in_size, int32_t* table_host) {
  if (in_size <= 0 || !table_host) {
    set_last_error("resize_build_table: bad arguments");
    return kErrInvalidArg;
  }
  // Resample.c precompute_coeffs (bilinear: support 1) + normalize_coeffs_8bpc, same double expressions in the same order
  const double in0 = 0.0, in1 = double(in_size);
  const double scale = (in1 - in0) / kOut;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 1.0 * filterscale;
  const int ksize = int(std::ceil(support)) * 2 + 1;
  const double ss = 1.0 / filterscale;
  int32_t* bounds = table_host;
  int32_t* kk = table_host + 2 * kOut;
  for (int i = 0; i < ksize * kOut; ++i) kk[i] = 0;
  double w[4096];
  if (ksize > 4096) {
    set_last_error("resize_build_table: source side %d too large", in_size);
    return kErrShape;
  }
  for (int xx = 0; xx < kOut; ++xx) {
    const double center = in0 + (xx + 0.5) * scale;
    int xmin = int(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = int(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      double a = (x + xmin - center + 0.5) * ss;
      if (a < 0.0) a = -a;
      w[x] = a < 1.0 ? 1.0 - a : 0.0;
      ww += w[x];
    }
    for (int x = 0; x < xmax; ++x) {
      if (ww != 0.0) w[x] /= ww;
      const double v = w[x] * double(1 << kPrecisionBits);
      kk[x * kOut + xx] = w[x] < 0 ? int(-0.5 + v) : int(0.5 + v);
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
  return kOk;
}

int sais_crop_resize_u8(const uint8_t* frames, int32_t N, int32_t H, int32_t W, int32_t top, int32_t left, int32_t ch,
                        int32_t cw, const int32_t* table_h, const int32_t* table_v, uint8_t* tmp, uint8_t* out,
                        sais_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (N == 0) return kOk;
  if (!frames || !table_h || !table_v || !tmp || !out || N < 0 || H <= 0 || W <= 0 || top < 0 || left < 0 || ch <= 0 ||
      cw <= 0 || top + ch > H || left + cw > W) {
    set_last_error("crop_resize_u8: bad arguments (N=%d H=%d W=%d box=%d,%d,%d,%d)", N, H, W, top, left, ch, cw);
    return kErrInvalidArg;
  }
  if ((reinterpret_cast<uintptr_t>(tmp) | reinterpret_cast<uintptr_t>(out)) & 3) {
    set_last_error("crop_resize_u8: tmp / out must be 4-byte aligned");
    return kErrInvalidArg;
  }
  const int row_pitch = ((cw * 3 + 15 + 15) & ~15) + 16;  // leading misalignment (< 16) + row, rounded up to 16-byte chunks
  const size_t smem = size_t(kRows) * row_pitch;
  if (smem > 200 * 1024) {
    set_last_error("crop_resize_u8: crop width %d too large", cw);
    return kErrShape;
  }
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(resize_horizontal_kernel),
                                   smem > 48 * 1024 ? 200 * 1024 : 48 * 1024, "resize_horizontal"))
    return rc;
  const int64_t total_bytes = int64_t(N) * H * W * 3;
  const int64_t groups = int64_t(N) * ((ch + kRows - 1) / kRows);
  const int64_t cap = int64_t(num_sms()) * 16;
  {
    LaunchScope ls(kClsPatchify, stream, double(N) * (double(ch) * cw * 3 + double(ch) * kRowBytes));
    int rc = check_cuda(launch_pdl(resize_horizontal_kernel, dim3(unsigned(groups < cap ? groups : cap)), dim3(kRowBytes), smem,
                                   stream, 1, frames, total_bytes, N, H, W, top, left, ch, cw, table_h, row_pitch, tmp),
                        "resize_horizontal launch");
    if (rc) return rc;
  }
  const int64_t rows = int64_t(N) * kOut;
  const int64_t capv = int64_t(num_sms()) * 32;
  LaunchScope ls(kClsPatchify, stream, double(N) * (double(ch) * kRowBytes + double(kOut) * kRowBytes));
  return check_cuda(launch_pdl(resize_vertical_kernel, dim3(unsigned(rows < capv ? rows : capv)), dim3(kRowBytes / 4), size_t(0),
                               stream, 1, tmp, N, ch, table_v, out),
                    "resize_vertical launch");
}

}  // extern "C"

// JPEG front-end of the frame pipeline: batched decode of same-sized JPEG frames straight into the
// uint8 [N,H,W,3] device buffer sais_crop_resize_u8 consumes.  The reference decodes every frame on the host with Pillow
// inside a single-threaded DataLoader (dino-main/main_dino.py:295-301,313 -> torchvision ImageFolder -> PIL.Image.open
// .convert('RGB'); extract_representations.py:178) — the true wall-clock bottleneck of its pipeline.  Both the RGB frames
// and the optical-flow frames reach the ViT this way: the RAFT stage writes its flow fields as `flows_%08d.jpg` images
// (extract_representations.py:246-261), so "externally supplied flow frames" are JPEG inputs like any other.
//
// Entropy decoding / IDCT is nvJPEG's (NVIDIA's library, like cuBLAS for a plain GEMM — there is nothing of this path's
// arithmetic in it); what is ours is the binding: one handle + batched state per device, interleaved-RGB output directly at
// frame pitch W*3, all frames of a call in ONE nvjpegDecodeBatched on the caller's stream, no staging copy.
// sais_jpeg_info is a host-only SOF-marker parse (no CUDA context needed).
#include <mutex>
#include <vector>

#include <nvjpeg.h>

#include "common.cuh"
#include "kernels.h"

namespace sais {
namespace {

struct JpegCtx {
  nvjpegHandle_t handle = nullptr;
  nvjpegJpegState_t state = nullptr;
  int batch = 0;  // batch size the state is initialised for
};
constexpr int kMaxDev = 64;
std::mutex g_mu;
JpegCtx g_ctx[kMaxDev];

const char* nvjpeg_err(nvjpegStatus_t s) {
  switch (s) {
    case NVJPEG_STATUS_SUCCESS: return "success";
    case NVJPEG_STATUS_NOT_INITIALIZED: return "not initialized";
    case NVJPEG_STATUS_INVALID_PARAMETER: return "invalid parameter";
    case NVJPEG_STATUS_BAD_JPEG: return "bad jpeg";
    case NVJPEG_STATUS_JPEG_NOT_SUPPORTED: return "jpeg not supported";
    case NVJPEG_STATUS_ALLOCATOR_FAILURE: return "allocator failure";
    case NVJPEG_STATUS_EXECUTION_FAILED: return "execution failed";
    case NVJPEG_STATUS_ARCH_MISMATCH: return "arch mismatch";
    case NVJPEG_STATUS_INTERNAL_ERROR: return "internal error";
    case NVJPEG_STATUS_IMPLEMENTATION_NOT_SUPPORTED: return "implementation not supported";
    default: return "unknown";
  }
}

int check_nvjpeg(nvjpegStatus_t s, const char* what) {
  if (s == NVJPEG_STATUS_SUCCESS) return kOk;
  set_last_error("%s: nvjpeg status %d (%s)", what, int(s), nvjpeg_err(s));
  return kErrCuda;
}

}  // namespace
}  // namespace sais

using namespace sais;

extern "C" {

int sais_jpeg_info(const uint8_t* data, size_t len, int32_t* hw2_host) {
  if (!data || !hw2_host || len < 4 || data[0] != 0xFF || data[1] != 0xD8) {
    set_last_error("jpeg_info: not a JPEG stream");
    return kErrInvalidArg;
  }
  size_t i = 2;
  while (i + 3 < len) {
    if (data[i] != 0xFF) {
      ++i;
      continue;
    }
    const uint8_t m = data[i + 1];
    if (m == 0xFF) {  // fill byte
      ++i;
      continue;
    }
    if (m == 0x01 || (m >= 0xD0 && m <= 0xD9)) {  // stand-alone markers
      i += 2;
      continue;
    }
    const size_t seg = (size_t(data[i + 2]) << 8) | data[i + 3];
    const bool sof = m >= 0xC0 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC;
    if (sof) {
      if (i + 9 > len) break;
      hw2_host[0] = (int32_t(data[i + 5]) << 8) | data[i + 6];  // height
      hw2_host[1] = (int32_t(data[i + 7]) << 8) | data[i + 8];  // width
      return kOk;
    }
    if (m == 0xDA) break;  // start of scan without a frame header
    i += 2 + seg;
  }
  set_last_error("jpeg_info: no frame header found");
  return kErrShape;
}

int sais_jpeg_decode_batch(const uint8_t* const* data_host, const size_t* lengths_host, int32_t n, int32_t H, int32_t W,
                           uint8_t* out_device, sais_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (n == 0) return kOk;
  if (!data_host || !lengths_host || !out_device || n < 0 || H <= 0 || W <= 0) {
    set_last_error("jpeg_decode_batch: bad arguments");
    return kErrInvalidArg;
  }
  for (int i = 0; i < n; ++i) {
    int32_t hw[2];
    if (int rc = sais_jpeg_info(data_host[i], lengths_host[i], hw)) return rc;
    if (hw[0] != H || hw[1] != W) {
      set_last_error("jpeg_decode_batch: frame %d is %dx%d, expected %dx%d (one call decodes same-sized frames)", i, hw[0],
                     hw[1], H, W);
      return kErrShape;
    }
  }
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) dev = 0;
  std::lock_guard<std::mutex> lk(g_mu);
  JpegCtx& c = g_ctx[dev];
  int rc;
  if (!c.handle) {
    // GPU-assisted Huffman decode for batches (baseline streams); nvJPEG falls back internally where it cannot
    nvjpegStatus_t s = nvjpegCreateEx(NVJPEG_BACKEND_GPU_HYBRID, nullptr, nullptr, 0, &c.handle);
    if (s != NVJPEG_STATUS_SUCCESS) s = nvjpegCreateSimple(&c.handle);
    if ((rc = check_nvjpeg(s, "nvjpegCreate"))) return rc;
    if ((rc = check_nvjpeg(nvjpegJpegStateCreate(c.handle, &c.state), "nvjpegJpegStateCreate"))) return rc;
  }
  if (c.batch != n) {
    if ((rc = check_nvjpeg(nvjpegDecodeBatchedInitialize(c.handle, c.state, n, 1, NVJPEG_OUTPUT_RGBI),
                           "nvjpegDecodeBatchedInitialize")))
      return rc;
    c.batch = n;
  }
  std::vector<nvjpegImage_t> dst(static_cast<size_t>(n));
  for (int i = 0; i < n; ++i) {
    for (int k = 0; k < NVJPEG_MAX_COMPONENT; ++k) {
      dst[i].channel[k] = nullptr;
      dst[i].pitch[k] = 0;
    }
    dst[i].channel[0] = out_device + size_t(i) * H * W * 3;
    dst[i].pitch[0] = size_t(W) * 3;
  }
  // (not counted by sais_launch_count: the kernels launched here are nvJPEG's, not this library's)
  return check_nvjpeg(nvjpegDecodeBatched(c.handle, c.state, data_host, lengths_host, dst.data(), stream),
                      "nvjpegDecodeBatched");
}

}  // extern "C"

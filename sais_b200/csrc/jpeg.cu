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
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include <nvjpeg.h>

#include "common.cuh"
#include "kernels.h"

namespace sais {
namespace {

struct JpegCtx {
  nvjpegHandle_t handle = nullptr;
  nvjpegJpegState_t state = nullptr;
  int batch = 0;     // batch size the state is initialised for
  bool tried = false;
  bool ok = false;
};
constexpr int kMaxDev = 64;
std::mutex g_mu;
// Two decoders per device, tried in this order for every call: the GPU's fixed-function JPEG engines
// (NVJPEG_BACKEND_HARDWARE: baseline streams, the common chroma subsamplings) and the CUDA decoder with Huffman decoding on
// the SMs (NVJPEG_BACKEND_GPU_HYBRID; nvJPEG's default backend if that cannot be created either).  A stream the engines
// reject (progressive, exotic sampling, pitch constraints) is decoded by the second one.  SAIS_JPEG_BACKEND=gpu skips the
// engines; sais_jpeg_last_backend() reports which decoder served the last call (1 = hardware engines, 2 = CUDA decoder).
JpegCtx g_ctx[kMaxDev][2];
int g_last_backend[kMaxDev];

// Threaded decoder (backend 3, the default): measured on the B200 boxes, nvjpegDecodeBatched spends its time in ONE host
// thread's Huffman decode whichever backend the handle was created with (6.7 ms per 1080p / 600 KB frame = 150 frames/s for
// GPU_HYBRID, HYBRID with 1 / 8 / 16 "cpu threads" alike; the fixed-function engines are refused by this nvJPEG build on
// sm_100: NVJPEG_STATUS_ARCH_MISMATCH).  Entropy decoding is serial per image but independent across images, so the frames of
// a call are dealt to T host threads, each with its own nvJPEG state and CUDA stream (the handle is shared: it is thread
// safe, states are not); every worker stream first waits for the caller's stream (the output buffer may still be read by
// earlier work) and the caller's stream then waits for every worker stream.  T = SAIS_JPEG_THREADS, default half the host
// cores (at most 16, at most one per frame).
struct JpegWorker {
  nvjpegJpegState_t state = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
};
struct JpegPool {
  nvjpegHandle_t handle = nullptr;
  std::vector<JpegWorker> workers;
  cudaEvent_t start = nullptr;
  bool tried = false, ok = false;
};
JpegPool g_pool[kMaxDev];

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
  static const bool skip_hw = getenv("SAIS_JPEG_BACKEND") != nullptr;
  std::vector<nvjpegImage_t> dst(static_cast<size_t>(n));
  for (int i = 0; i < n; ++i) {
    for (int k = 0; k < NVJPEG_MAX_COMPONENT; ++k) {
      dst[i].channel[k] = nullptr;
      dst[i].pitch[k] = 0;
    }
    dst[i].channel[0] = out_device + size_t(i) * H * W * 3;
    dst[i].pitch[0] = size_t(W) * 3;
  }
  int rc = kErrCuda;
  static const bool batched_api = getenv("SAIS_JPEG_BACKEND") != nullptr;  // gpu / cpu: the single-call batched API below
  if (!batched_api) {
    JpegPool& pl = g_pool[dev];
    if (!pl.tried) {
      pl.tried = true;
      static const int env_t = getenv("SAIS_JPEG_THREADS") ? atoi(getenv("SAIS_JPEG_THREADS")) : 0;
      int T = env_t > 0 ? env_t : int(std::thread::hardware_concurrency() / 2);
      T = T < 1 ? 1 : (T > 16 ? 16 : T);
      bool ok = nvjpegCreateSimple(&pl.handle) == NVJPEG_STATUS_SUCCESS &&
                cudaEventCreateWithFlags(&pl.start, cudaEventDisableTiming) == cudaSuccess;
      pl.workers.resize(size_t(T));
      for (int t = 0; ok && t < T; ++t) {
        JpegWorker& w = pl.workers[size_t(t)];
        ok = nvjpegJpegStateCreate(pl.handle, &w.state) == NVJPEG_STATUS_SUCCESS &&
             cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&w.done, cudaEventDisableTiming) == cudaSuccess;
      }
      pl.ok = ok;
    }
    if (pl.ok) {
      const int T = int(pl.workers.size()) < n ? int(pl.workers.size()) : n;
      if ((rc = check_cuda(cudaEventRecord(pl.start, stream), "jpeg: event record"))) return rc;
      std::vector<nvjpegStatus_t> st(size_t(T), NVJPEG_STATUS_SUCCESS);
      std::vector<std::thread> th;
      th.reserve(size_t(T));
      for (int t = 0; t < T; ++t) {
        th.emplace_back([&, t]() {
          JpegWorker& w = pl.workers[size_t(t)];
          if (cudaSetDevice(dev) != cudaSuccess || cudaStreamWaitEvent(w.stream, pl.start, 0) != cudaSuccess) {
            st[size_t(t)] = NVJPEG_STATUS_EXECUTION_FAILED;
            return;
          }
          for (int i = t; i < n; i += T) {  // interleaved: consecutive frames of a video have similar entropy-coded sizes
            nvjpegStatus_t r = nvjpegDecode(pl.handle, w.state, data_host[i], lengths_host[i], NVJPEG_OUTPUT_RGBI, &dst[size_t(i)],
                                            w.stream);
            if (r != NVJPEG_STATUS_SUCCESS) {
              st[size_t(t)] = r;
              return;
            }
          }
          if (cudaEventRecord(w.done, w.stream) != cudaSuccess) st[size_t(t)] = NVJPEG_STATUS_EXECUTION_FAILED;
        });
      }
      for (auto& x : th) x.join();
      for (int t = 0; t < T; ++t)
        if (st[size_t(t)] != NVJPEG_STATUS_SUCCESS) return check_nvjpeg(st[size_t(t)], "nvjpegDecode (threaded)");
      for (int t = 0; t < T; ++t)
        if ((rc = check_cuda(cudaStreamWaitEvent(stream, pl.workers[size_t(t)].done, 0), "jpeg: join worker stream"))) return rc;
      g_last_backend[dev] = 3;
      return kOk;
    }
  }
  for (int which = skip_hw ? 1 : 0; which < 2; ++which) {
    JpegCtx& c = g_ctx[dev][which];
    if (!c.tried) {
      c.tried = true;
      static const char* be = getenv("SAIS_JPEG_BACKEND");
      const nvjpegBackend_t second = (be && !strcmp(be, "cpu")) ? NVJPEG_BACKEND_HYBRID : NVJPEG_BACKEND_GPU_HYBRID;
      nvjpegStatus_t s = nvjpegCreateEx(which == 0 ? NVJPEG_BACKEND_HARDWARE : second, nullptr, nullptr, 0, &c.handle);
      static const bool dbg = getenv("SAIS_JPEG_DEBUG") != nullptr;
      if (dbg) fprintf(stderr, "[sais jpeg] nvjpegCreateEx(%s) -> %d (%s)\n", which == 0 ? "HARDWARE" : "GPU_HYBRID", int(s), nvjpeg_err(s));
      if (s != NVJPEG_STATUS_SUCCESS && which == 1) s = nvjpegCreateSimple(&c.handle);
      if (s == NVJPEG_STATUS_SUCCESS && nvjpegJpegStateCreate(c.handle, &c.state) == NVJPEG_STATUS_SUCCESS) {
        c.ok = true;
      } else if (which == 1) {
        return check_nvjpeg(s != NVJPEG_STATUS_SUCCESS ? s : NVJPEG_STATUS_INTERNAL_ERROR, "nvjpegCreate");
      }
    }
    if (!c.ok) continue;
    if (c.batch != n) {
      // (max_cpu_threads: host-side Huffman workers of the hybrid path; the engines ignore it)
      static const int threads = getenv("SAIS_JPEG_THREADS") ? atoi(getenv("SAIS_JPEG_THREADS")) : 4;
      nvjpegStatus_t s = nvjpegDecodeBatchedInitialize(c.handle, c.state, n, threads > 0 ? threads : 1, NVJPEG_OUTPUT_RGBI);
      if (s != NVJPEG_STATUS_SUCCESS) {
        rc = check_nvjpeg(s, "nvjpegDecodeBatchedInitialize");
        if (getenv("SAIS_JPEG_DEBUG")) fprintf(stderr, "[sais jpeg] batched init with backend %d failed: %d (%s)\n", which, int(s), nvjpeg_err(s));
        if (which == 0) { c.ok = false; continue; }
        return rc;
      }
      c.batch = n;
    }
    // (not counted by sais_launch_count: the kernels launched here are nvJPEG's, not this library's)
    nvjpegStatus_t s = nvjpegDecodeBatched(c.handle, c.state, data_host, lengths_host, dst.data(), stream);
    if (s == NVJPEG_STATUS_SUCCESS) {
      g_last_backend[dev] = which + 1;
      return kOk;
    }
    rc = check_nvjpeg(s, which == 0 ? "nvjpegDecodeBatched (hardware engines)" : "nvjpegDecodeBatched");
    if (getenv("SAIS_JPEG_DEBUG")) fprintf(stderr, "[sais jpeg] decode with backend %d failed: %d (%s)\n", which, int(s), nvjpeg_err(s));
    c.batch = 0;  // the batched state is undefined after a failed decode: re-initialise before the next use
  }
  return rc;
}

int sais_jpeg_last_backend(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) dev = 0;
  std::lock_guard<std::mutex> lk(g_mu);
  return g_last_backend[dev];
}

}  // extern "C"

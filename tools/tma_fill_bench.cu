// Microbenchmark (dev tool, not part of the library): how many bytes per clock can one SM pull from L2 with TMA,
// unicast vs. cluster multicast?  Decides whether weight tiles should be multicast across CTA pairs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I sais_b200/csrc -o gpurun_out/tma_fill tools/tma_fill_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cudaTypedefs.h>

#include "common.cuh"

using namespace sais;

namespace sais {
void set_last_error(const char*, ...) {}
int check_cuda(cudaError_t e, const char*) { return e == cudaSuccess ? 0 : -3; }
}  // namespace sais

constexpr int kBoxRows = 128;             // 128 rows x 128 B = 16 KB per box
constexpr int kBoxBytes = kBoxRows * 128;

// Each CTA streams `iters` boxes through a ring of `stages` slots.  MC = multicast width (1 = unicast): with MC > 1
// every CTA of the cluster issues 1/MC of each box (kBoxRows / MC rows) and multicasts it to all MC CTAs, so every SM
// still RECEIVES whole boxes while L2 is read once per cluster.
template <int MC>
__global__ void __launch_bounds__(64, 1) fill_kernel(const __grid_constant__ CUtensorMap tmap, int iters, int stages,
                                                     int rows_total, long long* cycles_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + stages * kBoxBytes);
  uint64_t* empty = full + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = MC > 1 ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], MC);  // every CTA of the cluster must have released the slot before anyone refills it
    }
    fence_mbar_init();
  }
  if (MC > 1) cluster_sync_all(); else __syncthreads();
  const long long t0 = clock64();
  const int cluster_id = blockIdx.x / MC;
  if (warp == 0 && lane == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < iters; ++i) {
      mbar_wait(&empty[stage], phase ^ 1);
      mbar_arrive_expect_tx(&full[stage], kBoxBytes);
      const int row0 = ((cluster_id * 977 + i) * kBoxRows) % (rows_total - kBoxRows);
      if (MC == 1) {
        tma_load_2d(smem + stage * kBoxBytes, &tmap, &full[stage], 0, row0);
      } else {
        constexpr int part = kBoxRows / MC;
        tma_load_2d_mcast(smem + stage * kBoxBytes + crank * part * 128, &tmap, &full[stage], 0, row0 + crank * part,
                          uint16_t((1u << MC) - 1));
      }
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < iters; ++i) {
      mbar_wait(&full[stage], phase);
      if (MC == 1) {
        mbar_arrive(&empty[stage]);
      } else {
        for (int r = 0; r < MC; ++r) {  // release the slot in every CTA of the cluster
          uint32_t remote;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(&empty[stage])), "r"(r));
          mbar_arrive_cluster(remote);
        }
      }
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles_out[blockIdx.x] = clock64() - t0;
  if (MC > 1) cluster_sync_all();
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
}

template <int MC>
void run(const CUtensorMap& tm, int stages, int iters, int rows_total, long long* d_cycles, int sm_mhz) {
  const int smem = stages * kBoxBytes + 1024 + 512;
  cudaFuncSetAttribute(fill_kernel<MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148 - 148 % MC);
  cfg.blockDim = dim3(64);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = MC; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    cudaError_t err = cudaLaunchKernelEx(&cfg, fill_kernel<MC>, tm, iters, stages, rows_total, d_cycles);
    cudaEventRecord(e1);
    if (err != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
      printf("MC=%d stages=%d: launch failed: %s\n", MC, stages, cudaGetErrorString(cudaGetLastError()));
      return;
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, d_cycles, sizeof(long long) * cfg.gridDim.x, cudaMemcpyDeviceToHost);
    long long mx = 0; for (unsigned i = 0; i < cfg.gridDim.x; ++i) mx = h[i] > mx ? h[i] : mx;
    const double bytes_per_sm = double(iters) * kBoxBytes;
    if (rep == 2)
      printf("MC=%d stages=%2d (%3d KB in flight/SM): %7.1f us  recv %6.2f TB/s  %5.1f B/clk/SM (clock64)  L2 reads %6.2f TB/s\n",
             MC, stages, stages * kBoxBytes / 1024, ms * 1e3, bytes_per_sm * cfg.gridDim.x / (ms * 1e-3) / 1e12,
             bytes_per_sm / double(mx), bytes_per_sm * cfg.gridDim.x / MC / (ms * 1e-3) / 1e12);
  }
}

int main() {
  const int rows_total = 32 * 1024 * 1024 / 128;  // 32 MB, L2 resident
  void* buf; cudaMalloc(&buf, size_t(rows_total) * 128);
  cudaMemset(buf, 1, size_t(rows_total) * 128);
  long long* d_cycles; cudaMalloc(&d_cycles, 148 * sizeof(long long));
  auto encode = get_encode();
  auto make = [&](int box_rows) {
    CUtensorMap tm;
    const cuuint64_t gdim[2] = {64, cuuint64_t(rows_total)};
    const cuuint64_t gstride[1] = {128};
    const cuuint32_t box[2] = {64, cuuint32_t(box_rows)};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) printf("encode failed %d\n", int(r));
    return tm;
  };
  const int iters = 4000;  // 64 MB per SM
  for (int stages : {2, 4, 8, 12}) run<1>(make(kBoxRows), stages, iters, rows_total, d_cycles, 0);
  for (int stages : {4, 8, 12}) run<2>(make(kBoxRows / 2), stages, iters, rows_total, d_cycles, 0);
  for (int stages : {4, 8, 12}) run<4>(make(kBoxRows / 4), stages, iters, rows_total, d_cycles, 0);
  return 0;
}

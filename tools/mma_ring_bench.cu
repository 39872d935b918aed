// Microbenchmark (dev tool): tcgen05.mma kind::f16 rate when the operands ROTATE through a ring of shared-memory
// stages like in the real GEMM (mma_rate_bench.cu re-reads the same four K slices), optionally with a
// tcgen05.commit after every k-block of 4 MMAs and an accumulator switch every 24 MMAs.  No TMA traffic.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I sais_b200/csrc -o tools/mma_ring tools/mma_ring_bench.cu
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

using namespace sais;
namespace sais {
void set_last_error(const char*, ...) {}
int check_cuda(cudaError_t e, const char*) { return e == cudaSuccess ? 0 : -3; }
bool pdl_enabled() { return false; }
}  // namespace sais

constexpr int kStageBytes = 48 * 1024;  // A 16 KB + B up to 32 KB
constexpr int kStages = 4;

// mode bit 0: rotate stages; bit 1: commit after every 4 MMAs; bit 2: switch accumulator every 24 MMAs;
// bit 3: issue from a single divergent lane (if (lane == 0)) instead of warp-uniform + elect
template <int CG, bool kSingle>
__global__ void __launch_bounds__(128, 1) mma_kernel(int N, int iters, int mode, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, sink;
  __shared__ uint32_t tmem_base_s;
  for (int i = threadIdx.x; i < kStages * kStageBytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i & 1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = CG > 1 ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&sink, 1 << 19); fence_mbar_init(); }
  if (warp == 0) {
    if (CG == 1) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
    else { tmem_alloc_cg2(&tmem_base_s, 512); tmem_relinquish_cg2(); }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  if (CG > 1) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 0 && crank == 0) {
    const uint32_t idesc = umma_idesc_bf16(128 * CG, N);
    const long long t0 = clock64();
    constexpr bool single = kSingle;
    if (!single || lane == 0) {
      int stage = 0, acc = 0;
      for (int kb = 0; kb < iters / 4; ++kb) {
        const uint32_t sa = smem_u32(smem + stage * kStageBytes);
        const uint64_t da = umma_desc_sw128_kmajor(sa), db = umma_desc_sw128_kmajor(sa + 16384);
        const uint32_t d = tmem + acc * 256;
        if (single || elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (CG == 1) umma_f16(d, da + 2 * k, db + 2 * k, idesc, 1);
            else umma_f16_cg2(d, da + 2 * k, db + 2 * k, idesc, 1);
          }
          if (mode & 2) { if (CG == 1) umma_commit(&sink); else umma_commit_cg2(&sink); }
        }
        if (!single) __syncwarp();
        if ((mode & 1) && ++stage == kStages) stage = 0;
        if ((mode & 4) && (kb % 6) == 5) acc ^= 1;
      }
      if (single || elect_one()) { if (CG == 1) umma_commit(&bar); else umma_commit_cg2(&bar); }
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    if (lane == 0) out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  if (CG > 1) cluster_sync_all(); else __syncthreads();
  if (warp == 0) { tc_fence_after(); if (CG == 1) tmem_dealloc(tmem, 512); else tmem_dealloc_cg2(tmem, 512); }
}

template <int CG, bool kSingle>
void run(int N, int mode, long long* d_out) {
  const int iters = 4800;
  const int smem = kStages * kStageBytes + 1024;
  cudaFuncSetAttribute(mma_kernel<CG, kSingle>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaMemset(d_out, 0, 148 * sizeof(long long));
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, mma_kernel<CG, kSingle>, N, iters, mode, d_out);
    if (e != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
      printf("CG=%d N=%3d mode=%d: failed: %s\n", CG, N, mode, cudaGetErrorString(cudaGetLastError()));
      return;
    }
  }
  long long h[148]; cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  const double cyc = double(mx) / iters;
  const double nominal = 128.0 * N / 256.0;
  printf("CG=%d N=%3d rotate=%d commit/kb=%d acc-switch=%d single-lane=%d: %6.1f cycles/MMA (nominal %5.1f, %4.1f%%)\n", CG, N,
         mode & 1, (mode >> 1) & 1, (mode >> 2) & 1, int(kSingle), cyc, nominal, 100.0 * nominal / cyc);
}

int main() {
  long long* d_out; cudaMalloc(&d_out, 148 * sizeof(long long));
  for (int cg = 1; cg <= 2; ++cg)
    for (int N : {128, 192, 256})
      for (int mode : {0, 1, 3, 7}) {
        if (cg == 1) { run<1, false>(N, mode, d_out); run<1, true>(N, mode, d_out); }
        else { run<2, false>(N, mode, d_out); run<2, true>(N, mode, d_out); }
      }
  return 0;
}

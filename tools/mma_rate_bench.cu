// Microbenchmark (dev tool): issue rate of tcgen05.mma kind::f16 as a function of N, CTA group and operand source
// (A from shared memory "SS" vs. A from tensor memory "TS"), with all operands resident (no TMA traffic).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I sais_b200/csrc -o tools/mma_rate tools/mma_rate_bench.cu
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

using namespace sais;
namespace sais {
void set_last_error(const char*, ...) {}
int check_cuda(cudaError_t e, const char*) { return e == cudaSuccess ? 0 : -3; }
}  // namespace sais

// smem: A tile 128 rows x 64 k (16 KB, SW128 K-major) at 0; B tile up to 256 rows x 64 k (32 KB) at 16 KB.
template <int CG, bool TS>
__global__ void __launch_bounds__(128, 1) mma_kernel(int N, int iters, int b_stride_k, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  const int warp = threadIdx.x >> 5;
  const uint32_t crank = CG > 1 ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) {
    if (CG == 1) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
    else { tmem_alloc_cg2(&tmem_base_s, 512); tmem_relinquish_cg2(); }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  if (CG > 1) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  long long dt = 0;
  if (threadIdx.x == 0 && crank == 0) {
    const uint32_t idesc = umma_idesc_bf16(128 * CG, N);
    const uint64_t da = umma_desc_sw128_kmajor(smem_u32(smem));
    const uint64_t db = umma_desc_sw128_kmajor(smem_u32(smem + 16384));
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int k = i & 3;
      if (TS) {
        if (CG == 1) umma_f16_ts(tmem, tmem + 256 + k * 8, db + 2 * k, idesc, 1);
        else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                          "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                          ::"r"(tmem), "r"(tmem + 256 + k * 8), "l"(db + 2 * k), "r"(idesc), "r"(1u) : "memory");
      } else {
        if (CG == 1) umma_f16(tmem, da + 2 * k, db + 2 * k, idesc, 1);
        else umma_f16_cg2(tmem, da + 2 * k, db + 2 * k, idesc, 1);
      }
    }
    if (CG == 1) umma_commit(&bar); else umma_commit_cg2(&bar);
    mbar_wait(&bar, 0);
    dt = clock64() - t0;
    out[blockIdx.x] = dt;
  }
  tc_fence_before();
  if (CG > 1) cluster_sync_all(); else __syncthreads();
  if (warp == 0) { tc_fence_after(); if (CG == 1) tmem_dealloc(tmem, 512); else tmem_dealloc_cg2(tmem, 512); }
}

template <int CG, bool TS>
void run(int N, long long* d_out) {
  const int iters = 4000;
  const int smem = 48 * 1024 + 1024;
  cudaFuncSetAttribute(mma_kernel<CG, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaMemset(d_out, 0, 148 * sizeof(long long));
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, mma_kernel<CG, TS>, N, iters, 0, d_out);
    if (e != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
      printf("CG=%d %s N=%3d: failed: %s\n", CG, TS ? "TS" : "SS", N, cudaGetErrorString(cudaGetLastError()));
      return;
    }
  }
  long long h[148]; cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  const double cyc = double(mx) / iters;
  const double nominal = 128.0 * N / 256.0;  // per SM: 128 rows x N x 16 MACs at 4096 MAC/clk
  printf("CG=%d %s M=%3d N=%3d: %6.1f cycles/instr (nominal %5.1f)  -> %5.0f MAC/clk/SM (%4.1f%% of 4096)\n", CG,
         TS ? "TS" : "SS", 128 * CG, N, cyc, nominal, 128.0 * N * 16 / cyc, 100.0 * nominal / cyc);
}

int main() {
  long long* d_out; cudaMalloc(&d_out, 148 * sizeof(long long));
  for (int N : {32, 64, 128, 192, 256}) run<1, false>(N, d_out);
  for (int N : {32, 64, 128, 192, 256}) run<1, true>(N, d_out);
  for (int N : {32, 64, 128, 192, 256}) run<2, false>(N, d_out);
  for (int N : {64, 128, 256}) run<2, true>(N, d_out);
  return 0;
}

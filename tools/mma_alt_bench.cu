// Microbenchmark (dev tool): tcgen05.mma throughput when two GEMM shapes alternate on one tensor pipe, as in the fused
// MLP kernel: G1 = 24 x (M256 N64 K16) into accumulator S, G2 = 8 x (M256 N192 K16) into accumulator acc.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I sais_b200/csrc -o tools/mma_alt tools/mma_alt_bench.cu
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

using namespace sais;
namespace sais {
void set_last_error(const char*, ...) {}
int check_cuda(cudaError_t e, const char*) { return e == cudaSuccess ? 0 : -3; }
}  // namespace sais

// mode bits: 1 = commit after each group; 2 = G2 uses A from TMEM (TS); 4 = G1 only; 8 = G2 only;
//            16 = background smem traffic from 8 warps; 32 = G1 with N=128 (12 MMAs... 24 MMAs of N128 = chunk 128)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
alt_kernel(int iters, int mode, int n1, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_base_s;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i & 255);
  const int warp = threadIdx.x >> 5;
  const uint32_t crank = cluster_ctarank();
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
  if (warp == 9) { tmem_alloc_cg2(&tmem_base_s, 512); tmem_relinquish_cg2(); }
  fence_proxy_async_smem();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  uint8_t* a_s = smem;                 // 96 KB
  uint8_t* w1_s = smem + 96 * 1024;    // 24 KB (up to 48 KB for n1 = 128)
  uint8_t* w2_s = smem + 144 * 1024;   // 24 KB
  uint8_t* h_s = smem + 168 * 1024;    // 16 KB (+16)
  if (warp == 9 && (threadIdx.x & 31) == 0 && crank == 0) {
    const uint32_t idesc1 = umma_idesc_bf16(256, n1);
    const uint32_t idesc2 = umma_idesc_bf16(256, 192);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (!(mode & 8)) {
        for (int kb = 0; kb < 6; ++kb) {
          const uint64_t da = umma_desc_sw128_kmajor(smem_u32(a_s + kb * 16384));
          const uint64_t db = umma_desc_sw128_kmajor(smem_u32(w1_s + kb * (n1 / 2) * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_cg2(tmem + 384 + (i & 1) * 64 * (n1 == 64), da + 2 * k, db + 2 * k, idesc1, (kb | k) != 0);
        }
        if (mode & 1) umma_commit_cg2_mcast(&bar[0], uint16_t(0b11));
      }
      if (!(mode & 4)) {
        const uint64_t da = umma_desc_sw128_kmajor(smem_u32(h_s + (i & 1) * 16384));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          for (int h = 0; h < 2; ++h) {
            const uint64_t db = umma_desc_sw128_kmajor(smem_u32(w2_s + h * 12288));
            if (mode & 2)
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                           "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                           ::"r"(tmem + h * 192), "r"(tmem + 384 + 64 + k * 8), "l"(db + 2 * k), "r"(idesc2), "r"(1u) : "memory");
            else
              umma_f16_cg2(tmem + h * 192, da + 2 * k, db + 2 * k, idesc2, 1);
          }
        if (mode & 1) umma_commit_cg2_mcast(&bar[1], uint16_t(0b11));
      }
    }
    umma_commit_cg2(&bar[0]);
    // wait for the final completion: poll the barrier phase the last commit flips
    const int commits0 = ((mode & 1) && !(mode & 8) ? iters : 0) + 1;
    mbar_wait(&bar[0], (commits0 - 1) & 1);
    out[blockIdx.x] = clock64() - t0;
  } else if ((mode & 16) && warp < 8) {
    // background smem traffic: every epilogue-like warp streams 16-byte stores over a private 4 KB window
    uint32_t addr = smem_u32(smem + 184 * 1024 + warp * 2048 + (threadIdx.x & 31) * 16);
    for (int i = 0; i < iters * 40; ++i) {
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(addr + ((i & 3) << 9)), "r"(i) : "memory");
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 9) { tc_fence_after(); tmem_dealloc_cg2(tmem, 512); }
}

void run(int mode, int n1, const char* what, long long* d_out) {
  const int iters = 400;
  const int smem = 201 * 1024 + 1024;
  cudaFuncSetAttribute(alt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaMemset(d_out, 0, 148 * sizeof(long long));
  for (int rep = 0; rep < 2; ++rep) {
    alt_kernel<<<148, 320, smem>>>(iters, mode, n1, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: failed: %s\n", what, cudaGetErrorString(e)); return; }
  }
  long long h[148]; cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  const double g1 = (mode & 8) ? 0 : 24 * (n1 == 64 ? 45.5 : n1 / 2.0), g2 = (mode & 4) ? 0 : 8 * 96.0;
  printf("%-46s mode=%2d n1=%3d: %8.1f cycles/round (sum of isolated rates %6.1f)\n", what, mode, n1, double(mx) / iters, g1 + g2);
}

int main() {
  long long* d_out; cudaMalloc(&d_out, 148 * sizeof(long long));
  run(4, 64, "G1 only (24 x N64)", d_out);
  run(8, 64, "G2 only (8 x N192)", d_out);
  run(0, 64, "G1,G2 alternating, no commits", d_out);
  run(1, 64, "G1,G2 alternating, commit per group", d_out);
  run(2, 64, "G1,G2(TS) alternating", d_out);
  run(3, 64, "G1,G2(TS) alternating, commits", d_out);
  run(16, 64, "G1,G2 alternating + smem store traffic", d_out);
  run(4, 128, "G1 only (24 x N128)", d_out);
  run(0, 128, "G1(N128),G2 alternating", d_out);
  run(2, 128, "G1(N128),G2(TS) alternating", d_out);
  return 0;
}

// Microbenchmark (dev tool): per-SM throughput of the instructions the GEMM epilogues lean on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/alu_rate tools/alu_rate_bench.cu
#include <cstdio>
#include <cstdint>
template <int OP>
__global__ void __launch_bounds__(512, 1) k(float* out, int iters, long long* cyc) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i * 0.1f;
  uint64_t p[4];
  for (int i = 0; i < 4; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
      if (OP == 4 && i < 4) asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p[i]));
      if (OP == 5) asm volatile("min.f32 %0, %0, 81.0;" : "+f"(a[i]));
      if (OP == 6) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %1;" : "=r"(r) : "f"(a[i])); a[i] = __uint_as_float(r); }
      if (OP == 7) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(*reinterpret_cast<uint32_t*>(&a[i])));
    }
  }
  const long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  for (int i = 0; i < 4; ++i) s += float(p[i] & 0xff);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP>
void run(const char* name, int per_iter, int elems) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  k<OP><<<148, 512>>>(out, iters, cyc);
  k<OP><<<148, 512>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
  const double instr = double(iters) * per_iter * 16;  // warp-instructions per SM (16 warps)
  printf("%-22s %6.2f cycles per warp-instruction per SMSP  -> %5.1f elements/clk/SM\n", name, double(mx) / (instr / 4),
         instr * 32 * elems / double(mx));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("tanh.approx.f32", 8, 1);
  run<1>("ex2.approx.f32", 8, 1);
  run<2>("rcp.approx.f32", 8, 1);
  run<3>("fma.rn.f32", 8, 1);
  run<4>("fma.rn.f32x2", 4, 2);
  run<5>("min.f32", 8, 1);
  run<6>("cvt.rn.bf16x2.f32", 8, 2);
  run<7>("tanh.approx.bf16x2", 8, 2);
  return 0;
}

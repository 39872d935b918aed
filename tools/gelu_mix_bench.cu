// Microbenchmark (dev tool): do the MUFU and FMA pipes overlap in the GELU epilogue's instruction mix?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/gelu_mix tools/gelu_mix_bench.cu
// OP 0: 4 tanh + 16 fma.f32x2 on independent registers per iteration (overlap -> ~33 cycles per SMSP, serial -> ~65)
// OP 1: the fc1 epilogue's arithmetic per 16 accumulators (2 FFMA2 fold + GELU-from-half + bf16 pack), registers only
// OP 2: OP 1 without the two tanh per pair;  OP 3: only the 16 tanh
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint64_t pack2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
template <int OP>
__global__ void __launch_bounds__(512, 1) k(float* out, int iters, long long* cyc) {
  float a[16];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i * 0.1f;
  uint64_t p[8];
  for (int i = 0; i < 8; ++i) p[i] = pack2(a[2 * i], a[2 * i + 1]);
  const uint64_t rs2 = pack2(0.5f + threadIdx.x * 1e-6f, 0.5f), nm2 = pack2(0.01f, 0.01f), c2 = pack2(0.3f, 0.2f), b2 = pack2(0.1f, 0.05f);
  uint32_t sink = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (OP == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p[i]));
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint64_t h = p[i];
        if (OP != 3) h = fma2(h, rs2, fma2(nm2, c2, b2 + i));
        float s0, s1, u0, u1, t0, t1;
        if (OP != 3) {
          unpack2(mul2(h, h), s0, s1);
          const uint64_t s = pack2(fminf(s0, 20.25f), fminf(s1, 20.25f));
          uint64_t q = fma2(pack2(-0.011248564f, -0.011248564f), s, pack2(0.29604521f, 0.29604521f));
          q = fma2(q, s, pack2(1.5950157f, 1.5950157f));
          unpack2(mul2(h, q), u0, u1);
        } else {
          unpack2(h, u0, u1);
        }
        if (OP == 1 || OP == 3) {
          asm volatile("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(u0));
          asm volatile("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(u1));
        } else {
          t0 = u0; t1 = u1;
        }
        if (OP != 3) {
          float y0, y1;
          unpack2(fma2(h, pack2(t0, t1), h), y0, y1);
          uint32_t r;
          asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(y1), "f"(y0));
          sink ^= r;
          p[i] = pack2(y0, y1);
        } else {
          p[i] = pack2(t0, t1);
        }
      }
    }
  }
  const long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 16; ++i) s += a[i];
  for (int i = 0; i < 8; ++i) s += float(p[i] & 0xff);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + float(sink);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP>
void run(const char* name, int threads) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  k<OP><<<148, threads>>>(out, iters, cyc);
  k<OP><<<148, threads>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
  const double warps_per_smsp = threads / 32 / 4.0;
  printf("%-44s %2d warps/SMSP: %7.1f cycles per iteration per SMSP (per warp-iteration %6.1f)\n", name, int(warps_per_smsp),
         double(mx) / iters, double(mx) / iters / warps_per_smsp);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int th : {128, 256, 512}) {
    run<0>("4 tanh + 16 fma2", th);
    run<1>("16-acc fold + GELU + pack (16 tanh)", th);
    run<2>("16-acc fold + GELU + pack, no tanh", th);
    run<3>("16 tanh only", th);
  }
  return 0;
}

// Micro-benchmark (run on the GPU box: nvcc -arch=sm_100a -O3 scripts/mufu_bench.cu -o /tmp/mufu && /tmp/mufu):
// MUFU throughput of the exp2 / tanh variants the attention and GELU epilogues could use, and tanh.approx accuracy.
#include <cuda_runtime.h>
#include <cstdio>
#include <cmath>
#include <cstdint>

template <int OP>
__device__ __forceinline__ uint32_t op(uint32_t x) {
  uint32_t y;
  if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 1) asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 3) asm volatile("tanh.approx.f32 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 4) asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 5) asm volatile("tanh.approx.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 6) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 7) asm volatile("{.reg .f32 t; mov.b32 t, %1; fma.rn.f32 t, t, 0f3F000000, 0f3F000000; mov.b32 %0, t;}" : "=r"(y) : "r"(x));
  return y;
}

template <int OP>
__global__ void bench(uint32_t* out, int iters, long long* cycles) {
  uint32_t a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 0x3c003c00u + threadIdx.x + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = op<OP>(a[i]);
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__global__ void tanh_err(float* maxerr, float* maxerr_tail) {
  float me = 0.f, mt = 0.f;
  for (int i = threadIdx.x + blockIdx.x * blockDim.x; i < 4000000; i += gridDim.x * blockDim.x) {
    float x = -10.f + 20.f * (float)i / 4000000.f;
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    float e = fabsf(y - tanhf(x));
    me = fmaxf(me, e);
    if (fabsf(x) > 2.f) mt = fmaxf(mt, e);
  }
  atomicMax((int*)maxerr, __float_as_int(me));
  atomicMax((int*)maxerr_tail, __float_as_int(mt));
}

template <int OP>
void run(const char* name, int elems_per_op) {
  uint32_t* out;
  long long* cyc;
  const int blocks = 148, threads = 1024, iters = 4096;
  cudaMalloc(&out, blocks * threads * 4);
  cudaMalloc(&cyc, blocks * 8);
  bench<OP><<<blocks, threads>>>(out, 16, cyc);
  bench<OP><<<blocks, threads>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < blocks; ++i) avg += h[i];
  avg /= blocks;
  double ops = (double)threads * iters * 8;
  printf("%-24s %8.2f instr-lanes/clk/SM  %8.2f elements/clk/SM\n", name, ops / avg, ops * elems_per_op / avg);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.f16x2", 2);
  run<2>("ex2.approx.ftz.bf16x2", 2);
  run<3>("tanh.approx.f32", 1);
  run<4>("tanh.approx.f16x2", 2);
  run<5>("tanh.approx.bf16x2", 2);
  run<6>("rcp.approx.ftz.f32", 1);
  run<7>("fma.rn.f32 (imm)", 1);
  float *e;
  cudaMalloc(&e, 8);
  cudaMemset(e, 0, 8);
  tanh_err<<<148, 256>>>(e, e + 1);
  float h[2];
  cudaMemcpy(h, e, 8, cudaMemcpyDeviceToHost);
  printf("tanh.approx.f32 max abs err on [-10,10]: %.3e ; for |x| > 2: %.3e\n", h[0], h[1]);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

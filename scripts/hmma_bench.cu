// Micro-benchmark (GPU box): throughput of the legacy warp-level mma.sync.m16n8k16 (bf16 / f16 -> f32) on sm_100a, per SM, as a
// function of resident warps — the bound of the mma.sync attention kernels (attn.cuh).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I lam_slide_b200/csrc scripts/hmma_bench.cu -o /tmp/hmma && /tmp/hmma
#include <cstdio>
#include "ptx.cuh"
using namespace lam;

__device__ __forceinline__ void mma_f16_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int F16, int NACC>
__global__ void hmma_loop(float* out, int iters) {
  float c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u};
  uint32_t b0 = threadIdx.x * 5u, b1 = 11u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (F16) mma_f16_16816(c[i], a, b0, b1);
      else mma_bf16_16816(c[i], a, b0, b1);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456f) out[0] = s;
}

// NH HMMA (m16n8k16 bf16) + NK HMMA (m16n8k8) + NM MUFU.EX2 per iteration, all independent: do the tensor and the MUFU pipe overlap?
template <int NH, int NK, int NM>
__global__ void mix_loop(float* out, int iters) {
  float c[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  float e[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) e[i] = -0.001f * (threadIdx.x + i);
  uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u};
  uint32_t b0 = threadIdx.x * 5u, b1 = 11u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < (NH > NM ? NH : NM); ++i) {
      if (i < NH) mma_bf16_16816(c[i & 7], a, b0, b1);
      if (i < NK) mma_bf16_1688(c[(i + 4) & 7], a, b0);
      if (i < NM) e[i & 15] = fast_exp2(e[i & 15]);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
#pragma unroll
  for (int i = 0; i < 16; ++i) s += e[i];
  if (s == 123.456f) out[0] = s;
}

template <int NH, int NK, int NM>
void run_mix(int warps) {
  float* d;
  cudaMalloc(&d, 4);
  const int iters = 20000;
  cudaEvent_t a, b;
  cudaEventCreate(&a), cudaEventCreate(&b);
  mix_loop<NH, NK, NM><<<148, warps * 32>>>(d, 100);
  cudaEventRecord(a);
  mix_loop<NH, NK, NM><<<148, warps * 32>>>(d, iters);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  printf("mix: %2d HMMA.16816 + %2d HMMA.1688 + %2d MUFU.EX2 per iteration, warps/SM=%2d: %7.1f cycles per iteration per SM sub-partition warp (at 1.9 GHz)\n",
         NH, NK, NM, warps, ms * 1e-3 * 1.9e9 / ((double)iters * warps / 4.0));
  cudaFree(d);
}

template <int F16, int NACC>
void run(int warps) {
  float* d;
  cudaMalloc(&d, 4);
  const int iters = 20000;
  cudaEvent_t a, b;
  cudaEventCreate(&a), cudaEventCreate(&b);
  hmma_loop<F16, NACC><<<148, warps * 32>>>(d, 100);
  cudaEventRecord(a);
  hmma_loop<F16, NACC><<<148, warps * 32>>>(d, iters);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double flops = 2.0 * 16 * 8 * 16 * (double)NACC * iters * warps * 148;
  printf("%s warps/SM=%2d independent accumulators=%d: %7.1f TFLOP/s   %5.2f cycles per HMMA per SM sub-partition (at 1.9 GHz)\n",
         F16 ? "f16 " : "bf16", warps, NACC, flops / ms * 1e-9, ms * 1e-3 * 1.9e9 / ((double)NACC * iters * warps / 4.0));
  cudaFree(d);
}

int main() {
  run<0, 4>(4);
  run<0, 8>(4);
  run<0, 8>(8);
  run<0, 8>(16);
  run<0, 8>(32);
  run<1, 8>(16);
  run_mix<10, 0, 0>(16);
  run_mix<0, 4, 0>(16);
  run_mix<10, 4, 0>(16);
  run_mix<0, 0, 16>(16);
  run_mix<10, 4, 16>(16);
  run_mix<10, 4, 8>(16);
  run_mix<6, 2, 16>(16);
  return 0;
}

// Micro-benchmark (GPU box): cycles per tcgen05.mma (kind::f16, bf16 x bf16 -> fp32, M = 128, cta_group::1, K = 16, SS operands in
// 128B-swizzled shared memory) as a function of N, issued back to back by one thread with no loads in the loop.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I lam_slide_b200/csrc scripts/mma_bench.cu -o /tmp/mma && /tmp/mma
#include <cstdio>
#include "ptx.cuh"
using namespace lam;

template <int N>
__global__ void __launch_bounds__(128, 1) mma_loop(long long* cycles, int iters, int same_k) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + N * 128) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1 && lane == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, N);
    const uint64_t a_desc = umma_desc_sw128(smem_u32(smem));
    const uint64_t b_desc = umma_desc_sw128(smem_u32(smem + 16384));
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem + (it & 1) * N * 0, a_desc + (same_k ? 0 : 2 * k), b_desc + (same_k ? 0 : 2 * k), idesc, 1);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

template <int N>
void run(int grid) {
  long long* d;
  cudaMalloc(&d, grid * 8);
  const int iters = 2000;
  const int smem = 16384 + N * 128 + 1024;
  cudaFuncSetAttribute(mma_loop<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  mma_loop<N><<<grid, 128, smem>>>(d, 10, 0);
  mma_loop<N><<<grid, 128, smem>>>(d, iters, 0);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < grid; ++i) avg += h[i];
  avg /= grid;
  printf("N=%3d grid=%3d: %7.1f cycles per MMA (ideal %3d)   %s\n", N, grid, avg / (iters * 4), N / 2, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}

int main() {
  run<64>(1);
  run<128>(1);
  run<192>(1);
  run<256>(1);
  run<64>(148);
  run<128>(148);
  run<192>(148);
  run<256>(148);
  return 0;
}

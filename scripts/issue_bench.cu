// Micro-benchmark (GPU box): cost of the MMA-issue protocol.  One "unit" = 4 tcgen05.mma (M = 128, N, K = 16, SS) as in the
// persistent GEMM kernels; modes add the pieces of the real issue loop one at a time:
//   0  one thread, MMAs only                            1  + tcgen05.commit per unit (ring of barriers, nobody waits)
//   2  + full/empty handshake with a producer warp       3  as 2 but warp-uniform loop with elect_one() (the shipped form)
//   4  as 3 with descriptors advanced by additions only (no per-unit descriptor rebuild)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I lam_slide_b200/csrc scripts/issue_bench.cu -o /tmp/issue && /tmp/issue
#include <cstdio>
#include "ptx.cuh"
using namespace lam;

constexpr int kStages = 4;

template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) issue_loop(long long* cycles, int iters) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_bar[kStages], empty_bar[kStages], done_bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int unit_bytes = N * 128;
  for (int i = threadIdx.x; i < (16384 + kStages * unit_bytes) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&full_bar[s], 1), mbar_init(&empty_bar[s], 1);
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = slot;
  uint8_t* ring = smem + 16384;
  if (warp == 2 && MODE >= 2) {  // producer: hands the ring slots over without loading anything
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&empty_bar[s], ph ^ 1);
      if (elect_one()) mbar_arrive(&full_bar[s]);
      __syncwarp();
      if (++s == kStages) s = 0, ph ^= 1;
    }
  }
  if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16(128, N);
    const uint64_t a_desc = umma_desc_sw128(smem_u32(smem));
    long long t0 = clock64();
    if (MODE <= 2) {
      if (lane == 0) {
        int s = 0;
        uint32_t ph = 0;
        for (int it = 0; it < iters; ++it) {
          if (MODE >= 2) {
            mbar_wait(&full_bar[s], ph);
            tcgen05_fence_after();
          }
          const uint64_t b_desc = umma_desc_sw128(smem_u32(ring + s * unit_bytes));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, 1);
          if (MODE >= 1) umma_commit(&empty_bar[s]);
          if (++s == kStages) s = 0, ph ^= 1;
        }
      }
    } else if (MODE == 3) {
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(&full_bar[s], ph);
        tcgen05_fence_after();
        const uint64_t b_desc = umma_desc_sw128(smem_u32(ring + s * unit_bytes));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, 1);
          umma_commit(&empty_bar[s]);
        }
        __syncwarp();
        if (++s == kStages) s = 0, ph ^= 1;
      }
    } else {
      int s = 0;
      uint32_t ph = 0;
      const uint64_t b_desc0 = umma_desc_sw128(smem_u32(ring));
      uint64_t b_desc = b_desc0;
      uint32_t full_a = smem_u32(&full_bar[0]), empty_a = smem_u32(&empty_bar[0]);
      const bool leader = elect_one();
      for (int it = 0; it < iters; ++it) {
        mbar_wait(reinterpret_cast<uint64_t*>(__cvta_shared_to_generic(full_a + s * 8)), ph);
        tcgen05_fence_after();
        if (leader) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, 1);
          umma_commit(reinterpret_cast<uint64_t*>(__cvta_shared_to_generic(empty_a + s * 8)));
        }
        b_desc += unit_bytes >> 4;
        if (++s == kStages) s = 0, ph ^= 1, b_desc = b_desc0;
      }
      __syncwarp();
    }
    if (lane == 0) {
      umma_commit(&done_bar);
      mbar_wait(&done_bar, 0);
      cycles[blockIdx.x] = clock64() - t0;
    }
    __syncwarp();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

template <int N, int MODE>
void run(int grid) {
  long long* d;
  cudaMalloc(&d, grid * 8);
  const int iters = 4000;
  const int smem = 16384 + kStages * N * 128 + 1024;
  cudaFuncSetAttribute(issue_loop<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  issue_loop<N, MODE><<<grid, 128, smem>>>(d, 10);
  issue_loop<N, MODE><<<grid, 128, smem>>>(d, iters);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < grid; ++i) avg += h[i];
  avg /= grid;
  printf("N=%3d mode=%d grid=%3d: %7.1f cycles per unit of 4 MMAs (tensor pipe alone %3d)   %s\n", N, MODE, grid, avg / iters, 2 * N,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}

template <int N>
void run_all() {
  run<N, 0>(148);
  run<N, 1>(148);
  run<N, 2>(148);
  run<N, 3>(148);
  run<N, 4>(148);
}

int main() {
  run_all<64>();
  run_all<128>();
  run_all<192>();
  run_all<256>();
  return 0;
}

// Micro-benchmark (GPU box): tcgen05.ld / tcgen05.st throughput per SM as a function of the number of warps issuing them.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I lam_slide_b200/csrc scripts/ldtm_bench.cu -o /tmp/ldtm && /tmp/ldtm
#include <cstdio>
#include "ptx.cuh"
using namespace lam;

template <int MODE>  // 0: ld x32, 1: ld x16, 2: st x16 (via inline asm below)
__global__ void __launch_bounds__(512, 1) ldtm_loop(long long* cycles, uint32_t* sink, int iters, int nwarps) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t taddr = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 64;
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  if (warp < nwarps) {
    uint32_t r[32];
    for (int i = 0; i < 32; ++i) r[i] = i;
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (MODE == 0) {
        tmem_ld32(taddr, r);
        tmem_ld32(taddr + 32, r);
        tmem_ld_wait();
      } else if (MODE == 1) {
        tmem_ld16(taddr, r);
        tmem_ld16(taddr + 16, r + 16);
        tmem_ld16(taddr + 32, r);
        tmem_ld16(taddr + 48, r + 16);
        tmem_ld_wait();
      } else {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
                     "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
                     "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr + 16),
                     "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
                     "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      acc += r[it & 31];
    }
    t1 = clock64();
  }
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc<512>(slot);
  }
}

template <int MODE>
void run(const char* name, int nwarps, int bytes_per_iter_per_thread) {
  long long* d;
  uint32_t* sink;
  const int grid = 148, iters = 4000;
  cudaMalloc(&d, grid * 8);
  cudaMalloc(&sink, grid * 512 * 4);
  ldtm_loop<MODE><<<grid, 512>>>(d, sink, 10, nwarps);
  ldtm_loop<MODE><<<grid, 512>>>(d, sink, iters, nwarps);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < grid; ++i) avg += h[i];
  avg /= grid;
  printf("%-10s %2d warps: %8.1f B/clk/SM   %s\n", name, nwarps, (double)bytes_per_iter_per_thread * 32 * nwarps * iters / avg,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
  cudaFree(sink);
}

int main() {
  for (int w : {4, 8, 16}) run<0>("ld.x32", w, 256);
  for (int w : {4, 8, 16}) run<1>("ld.x16", w, 256);
  for (int w : {4, 8, 16}) run<2>("st.x16", w, 128);
  return 0;
}

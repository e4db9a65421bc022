// Micro-benchmark of global store patterns (run on the GPU box): what write bandwidth does a tile epilogue get when a warp
// instruction writes SEG contiguous bytes in each of several rows (row pitch PITCH bytes), vs fully contiguous stores.
// nvcc -arch=sm_100a -O3 scripts/store_bench.cu -o /tmp/sb && /tmp/sb
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

// grid = 148 CTAs x 512 threads (16 warps).  The output is [rows, pitch] bytes.  A CTA walks 128-row blocks; inside a block it
// walks column tiles of TILE bytes; warp w writes rows 32*(w%4).. of the column quarter w/4 (TILE/4 bytes wide) in units of
// SEG bytes per row per instruction group.
template <int SEG>
__global__ void __launch_bounds__(512) pattern_kernel(uint8_t* out, int rows, int pitch, int tile_bytes) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, cq = warp >> 2;
  const int qw = tile_bytes / 4;           // bytes per warp per row
  constexpr int UNITS = SEG / 16;          // 16-byte units per row segment
  const uint4 val = make_uint4(lane, warp, blockIdx.x, 7);
  for (int mb = blockIdx.x; mb * 128 < rows; mb += gridDim.x) {
    const int row0 = mb * 128 + q * 32;
    for (int t0 = 0; t0 + tile_bytes <= pitch; t0 += tile_bytes) {
      for (int s0 = 0; s0 < qw; s0 += SEG) {
#pragma unroll
        for (int k = 0; k < UNITS; ++k) {
          const int id = lane + 32 * k;
          const int r = id / UNITS, ch = id % UNITS;
          if (row0 + r < rows) *reinterpret_cast<uint4*>(out + (size_t)(row0 + r) * pitch + t0 + cq * qw + s0 + ch * 16) = val;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(512) linear_kernel(uint4* out, size_t n16) {
  const uint4 val = make_uint4(threadIdx.x, 1, blockIdx.x, 7);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) out[i] = val;
}

template <class F>
float time_ms(F f, int iters = 10) {
  cudaEvent_t a, b;
  cudaEventCreate(&a), cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < iters; ++i) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms / iters;
}

int main() {
  const int rows = 128000, pitch = 5376;  // 2688 bf16 columns
  const size_t bytes = (size_t)rows * pitch;
  uint8_t* out;
  cudaMalloc(&out, bytes);
  cudaMemset(out, 0, bytes);
  float ms = time_ms([&] { linear_kernel<<<148 * 4, 512>>>((uint4*)out, bytes / 16); });
  printf("linear fill                       : %7.1f us  %6.2f TB/s\n", ms * 1e3, bytes / ms * 1e-9);
  ms = time_ms([&] { pattern_kernel<48><<<148, 512>>>(out, rows, pitch, 384); });
  printf("tile 384 B, 48 B row segments     : %7.1f us  %6.2f TB/s\n", ms * 1e3, bytes / ms * 1e-9);
  ms = time_ms([&] { pattern_kernel<96><<<148, 512>>>(out, rows, pitch, 384); });
  printf("tile 384 B, 96 B row segments     : %7.1f us  %6.2f TB/s\n", ms * 1e3, bytes / ms * 1e-9);
  ms = time_ms([&] { pattern_kernel<128><<<148, 512>>>(out, rows, pitch, 512); });
  printf("tile 512 B, 128 B row segments    : %7.1f us  %6.2f TB/s\n", ms * 1e3, (size_t)rows * (pitch / 512 * 512) / ms * 1e-9);
  ms = time_ms([&] { pattern_kernel<256><<<148, 512>>>(out, rows, pitch, 1024); });
  printf("tile 1024 B, 256 B row segments   : %7.1f us  %6.2f TB/s\n", ms * 1e3, (size_t)rows * (pitch / 1024 * 1024) / ms * 1e-9);
  ms = time_ms([&] { pattern_kernel<96><<<148 * 2, 512>>>(out, rows, pitch, 384); });
  printf("tile 384 B, 96 B segs, 296 CTAs   : %7.1f us  %6.2f TB/s\n", ms * 1e3, bytes / ms * 1e-9);
  ms = time_ms([&] { pattern_kernel<96><<<148 * 4, 512>>>(out, rows, pitch, 384); });
  printf("tile 384 B, 96 B segs, 592 CTAs   : %7.1f us  %6.2f TB/s\n", ms * 1e3, bytes / ms * 1e-9);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

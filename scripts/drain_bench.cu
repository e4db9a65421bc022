// Micro-benchmark (GPU box): how fast can 16 warps of one CTA add a [128 x 384] fp32 tile into global memory (h += tile)?
// This is the "drain" at the end of an m-block of the fused MLP kernel.  Variants:
//   0  TMA reduce-add,  8-column boxes (32-byte rows, SWIZZLE_32B), two staging boxes per warp
//   1  TMA reduce-add, 16-column boxes (64-byte rows, SWIZZLE_64B), one staging box per warp
//   2  TMA reduce-add, 32-column boxes (128-byte rows, SWIZZLE_128B), one staging box per warp
//   3  TMA store (no add), 16-column boxes
//   4  red.global.add.v4.f32 straight from registers (thread = row, 64 contiguous bytes per thread and box)
//   5  ld.global.v4 + add + st.global.v4 straight from registers (rows are owned by the CTA: no atomicity needed)
//   6  as 5 with all loads of the warp's 96 columns issued before the first store (6 x 16 registers)
//   7  as 1 with two staging boxes per warp (reduce-adds of two boxes in flight)
//   8  as 1 with all 16 warps' boxes of one column group issued by ONE thread after a CTA barrier (fewer, back-to-back TMA ops)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I lam_slide_b200/csrc scripts/drain_bench.cu -lcuda -o /tmp/drain && /tmp/drain
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include "ptx.cuh"
using namespace lam;

constexpr int H = 384;

__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int V>
__global__ void __launch_bounds__(512, 1) drain_kernel(const __grid_constant__ CUtensorMap tm, float* h, int num_m_blocks) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int q = warp & 3, cq = warp >> 2;
  const bool issuer = elect_one();
  const float4 val = make_float4(1.f, 2.f, 3.f, 4.f);
  for (int mb = blockIdx.x; mb < num_m_blocks; mb += gridDim.x) {
    const int row0 = mb * 128 + q * 32;
    const int col0 = cq * 96;
    if (V == 0) {
      const uint32_t st = smem_u32(smem) + warp * 2048;
      for (int bx = 0; bx < 12; ++bx) {
        if (issuer) bulk_wait_read<1>();
        __syncwarp();
        const uint32_t s = st + (bx & 1) * 1024;
        st_shared_v4(s + lane * 32 + (((lane >> 2) & 1) << 4), 1, 2, 3, 4);
        st_shared_v4(s + lane * 32 + ((1 ^ ((lane >> 2) & 1)) << 4), 1, 2, 3, 4);
        fence_proxy_async();
        __syncwarp();
        if (issuer) {
          tma_reduce_add_2d_s(&tm, s, col0 + bx * 8, row0);
          bulk_commit();
        }
      }
    } else if (V == 1 || V == 3) {
      const uint32_t s = smem_u32(smem) + warp * 2048;
      for (int bx = 0; bx < 6; ++bx) {
        if (issuer) bulk_wait_read<0>();
        __syncwarp();
        for (int ch = 0; ch < 4; ++ch) st_shared_v4(s + lane * 64 + ((ch ^ ((lane >> 1) & 3)) << 4), 1, 2, 3, 4);
        fence_proxy_async();
        __syncwarp();
        if (issuer) {
          if (V == 1) tma_reduce_add_2d_s(&tm, s, col0 + bx * 16, row0);
          else tma_store_2d_s(&tm, s, col0 + bx * 16, row0);
          bulk_commit();
        }
      }
    } else if (V == 7) {
      const uint32_t st = smem_u32(smem) + warp * 4096;
      for (int bx = 0; bx < 6; ++bx) {
        if (issuer) bulk_wait_read<1>();
        __syncwarp();
        const uint32_t s = st + (bx & 1) * 2048;
        for (int ch = 0; ch < 4; ++ch) st_shared_v4(s + lane * 64 + ((ch ^ ((lane >> 1) & 3)) << 4), 1, 2, 3, 4);
        fence_proxy_async();
        __syncwarp();
        if (issuer) {
          tma_reduce_add_2d_s(&tm, s, col0 + bx * 16, row0);
          bulk_commit();
        }
      }
    } else if (V == 2) {
      const uint32_t s = smem_u32(smem) + warp * 4096;
      for (int bx = 0; bx < 3; ++bx) {
        if (issuer) bulk_wait_read<0>();
        __syncwarp();
        for (int ch = 0; ch < 8; ++ch) st_shared_v4(s + lane * 128 + ((ch ^ (lane & 7)) << 4), 1, 2, 3, 4);
        fence_proxy_async();
        __syncwarp();
        if (issuer) {
          tma_reduce_add_2d_s(&tm, s, col0 + bx * 32, row0);
          bulk_commit();
        }
      }
    } else if (V == 4) {
      float* p = h + (size_t)(row0 + lane) * H + col0;
      for (int bx = 0; bx < 6; ++bx)
        for (int ch = 0; ch < 4; ++ch) red_add_v4(p + bx * 16 + ch * 4, val);
    } else if (V == 5) {
      float4* p = reinterpret_cast<float4*>(h + (size_t)(row0 + lane) * H + col0);
      for (int bx = 0; bx < 6; ++bx) {
        float4 x[4];
        for (int ch = 0; ch < 4; ++ch) x[ch] = p[bx * 4 + ch];
        for (int ch = 0; ch < 4; ++ch) p[bx * 4 + ch] = make_float4(x[ch].x + 1.f, x[ch].y + 2.f, x[ch].z + 3.f, x[ch].w + 4.f);
      }
    } else {
      float4* p = reinterpret_cast<float4*>(h + (size_t)(row0 + lane) * H + col0);
      float4 x[24];
#pragma unroll
      for (int i = 0; i < 24; ++i) x[i] = p[i];
#pragma unroll
      for (int i = 0; i < 24; ++i) p[i] = make_float4(x[i].x + 1.f, x[i].y + 2.f, x[i].z + 3.f, x[i].w + 4.f);
    }
  }
  if (V <= 3 && issuer) bulk_wait_read<0>();
}

static CUtensorMap make_map(float* h, int rows, int box_cols, CUtensorMapSwizzle sw) {
  CUtensorMap m;
  cuuint64_t gdim[2] = {(cuuint64_t)H, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {H * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, h, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                                      CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
  return m;
}

template <int V>
void run(float* h, int rows, int grid, const char* name) {
  const int box_cols = V == 0 ? 8 : (V == 2 ? 32 : 16);  // V == 7: 16
  const CUtensorMapSwizzle sw = V == 0 ? CU_TENSOR_MAP_SWIZZLE_32B : (V == 2 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
  CUtensorMap tm = make_map(h, rows, box_cols, sw);
  const int smem = 65536 + 1024;
  cudaFuncSetAttribute(drain_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int mblocks = rows / 128;
  cudaEvent_t a, b;
  cudaEventCreate(&a), cudaEventCreate(&b);
  drain_kernel<V><<<grid, 512, smem>>>(tm, h, mblocks);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < 5; ++i) drain_kernel<V><<<grid, 512, smem>>>(tm, h, mblocks);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  ms /= 5;
  const double per_cta_blocks = (double)((mblocks + grid - 1) / grid);
  printf("%-58s grid=%3d: %8.1f us  %6.2f TB/s of tile bytes  %7.0f ns per m-block per CTA   %s\n", name, grid, ms * 1e3,
         (double)rows * H * 4 / ms * 1e-9, ms * 1e6 / per_cta_blocks, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  cuInit(0);
  const int rows = 128000;
  float* h;
  cudaMalloc(&h, (size_t)rows * H * 4);
  cudaMemset(h, 0, (size_t)rows * H * 4);
  for (int grid : {148, 2}) {
    const int r = grid == 2 ? 128 * 2 * 50 : rows;
    run<0>(h, r, grid, "0 TMA reduce  8-col boxes x2");
    run<1>(h, r, grid, "1 TMA reduce 16-col boxes");
    run<2>(h, r, grid, "2 TMA reduce 32-col boxes");
    run<3>(h, r, grid, "3 TMA store  16-col boxes");
    run<4>(h, r, grid, "4 red.global.add.v4.f32 from registers");
    run<5>(h, r, grid, "5 ld + add + st from registers");
    run<6>(h, r, grid, "6 ld x24, then st x24");
    run<7>(h, r, grid, "7 TMA reduce 16-col boxes x2");
  }
  return 0;
}

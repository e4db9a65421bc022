// fp32-accurate linear layers of the first stage on the 5th-generation tensor cores: 3xTF32 with tcgen05.mma.kind::tf32.
//
//   Y[r, n] = epi( sum_k X[r, k] W[n, k] + bias[n] )       (same contract and epilogue options as LinearArgs / linear_f32_tc_kernel)
//
// x = hi + lo with hi = x rounded to TF32 and lo = x - hi (exact in fp32); x.w is accumulated as lo.hi + hi.lo + hi.hi in the fp32
// TMEM accumulator, the dropped lo.lo term is 2^-22 relative: as good as an fp32 FMA chain (the reference's first stage is fp32 and
// its outputs are checked to 1e-4; measured 1.4e-6).  The weights are split once at *_create (w_hi / w_lo, fp32 arrays the tensor
// core reads as TF32); the activations are split on the fly in shared memory.
//
//   warp 0      TMA producer: per 32-wide k-block the raw fp32 A tile (128 rows x 128 B, 128-byte swizzle) and the W_hi / W_lo tiles
//   warps 2..5  splitters: rewrite the landed A tile as hi in place and write lo next to it (element-wise, so the swizzle is
//               irrelevant), fence.proxy.async, arrive
//   warp 1      MMA issuer: 3 tcgen05.mma (M128 x BN x K8) per 8-wide k step into one of two TMEM accumulators
//   warps 6..13 epilogue, thread = (row, column half): tcgen05.ld -> bias / GELU / row-add / residual / SiLU -> global fp32
//
// Persistent (one CTA per SM, tiles n-fastest so the CTAs that share an A tile run together and HBM sees it once); the epilogue of
// tile i overlaps the main loop of tile i + 1 through the second accumulator.  The mma.sync 3xTF32 kernel this replaces reached
// ~100-120 TFLOP/s of tensor work (every fragment split in registers, 24 MMAs per k8 step per warp); the layers here are
// 2.4 - 34 GFLOP x 3 over 0.1 - 0.5 GB, i.e. HBM-bound once the tensor work runs at tcgen05 rate.
#pragma once
#include <cuda.h>

#include "first_stage.cuh"
#include "ptx.cuh"

namespace lam {

constexpr int kL5Threads = 32 * 14;
constexpr int kL5BK = 32;                       // fp32 per k-block = one 128-byte swizzle row
constexpr int kL5ABytes = 128 * kL5BK * 4;      // 16 KB

template <int BN>
struct L5Cfg {
  static constexpr int kBBytes = BN * kL5BK * 4;
  static constexpr int kStageBytes = 2 * kL5ABytes + 2 * kBBytes;  // A (-> hi), A lo, W hi, W lo
  static constexpr int kStagesRaw = (200 * 1024) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 4 ? 4 : kStagesRaw;
  static constexpr int kAccStride = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  static constexpr uint32_t kTmemCols = 2 * kAccStride;
  static constexpr int kSmem = kStages * kStageBytes + 256 + 1024;
};

// kind::tf32, TF32 x TF32 -> FP32, both operands K-major (cute::UMMA::InstrDescriptor: a_format = b_format = 2)
__host__ __device__ constexpr uint32_t umma_idesc_tf32(uint32_t M, uint32_t N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// round-to-nearest TF32 head (ties away from zero in magnitude) and the exact remainder
__host__ __device__ __forceinline__ float tf32_head(float x) {
#ifdef __CUDA_ARCH__
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
#else
  uint32_t u;
  memcpy(&u, &x, 4);
  u = (u + 0x1000u) & 0xffffe000u;
  float r;
  memcpy(&r, &u, 4);
  return r;
#endif
}

template <int BN>
__global__ void __launch_bounds__(kL5Threads, 1)
linear_tc5_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_whi,
                  const __grid_constant__ CUtensorMap tm_wlo, LinearArgs a, int m_tiles, int n_tiles) {
  using C = L5Cfg<BN>;
  static_assert(BN % 32 == 0 && BN <= 256, "two column halves of 16-column chunks");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* split_bar = full_bar + C::kStages;
  uint64_t* empty_bar = split_bar + C::kStages;
  uint64_t* acc_full = empty_bar + C::kStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (a.K + kL5BK - 1) / kL5BK;
  const int tiles = m_tiles * n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_whi);
    tma_prefetch_desc(&tm_wlo);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&split_bar[s], 4);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::kTmemCols>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const int m0 = (tile / n_tiles) * 128, n0 = (tile % n_tiles) * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        if (elect_one()) {
          uint8_t* st = smem + s * C::kStageBytes;
          mbar_arrive_expect_tx(&full_bar[s], kL5ABytes + 2 * C::kBBytes);
          tma_load_2d(&tm_a, &full_bar[s], st, kb * kL5BK, m0);
          tma_load_2d(&tm_whi, &full_bar[s], st + 2 * kL5ABytes, kb * kL5BK, n0);
          tma_load_2d(&tm_wlo, &full_bar[s], st + 2 * kL5ABytes + C::kBBytes, kb * kL5BK, n0);
        }
        __syncwarp();
        if (++s == C::kStages) s = 0, ph ^= 1;
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = umma_idesc_tf32(128, BN);
    int s = 0, it = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      mbar_wait(&acc_empty[acc], ((it >> 1) & 1) ^ 1);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + acc * C::kAccStride;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[s], ph);   // W tiles (and the raw A tile)
        mbar_wait(&split_bar[s], ph);  // A hi / lo written and fenced
        tcgen05_fence_after();
        const uint32_t st = smem_u32(smem + s * C::kStageBytes);
        const uint64_t a_hi = umma_desc_sw128(st), a_lo = umma_desc_sw128(st + kL5ABytes);
        const uint64_t w_hi = umma_desc_sw128(st + 2 * kL5ABytes), w_lo = umma_desc_sw128(st + 2 * kL5ABytes + C::kBBytes);
        const int rem = a.K - kb * kL5BK;
        const int ksteps = rem >= kL5BK ? 4 : (rem + 7) >> 3;  // the TMA zero-fills past K; skip the all-zero k steps
        if (elect_one()) {
          for (int k = 0; k < ksteps; ++k) {
            // 8 fp32 = 32 B inside the 128 B swizzle atom: +2 in the (addr >> 4) field; small terms first
            umma_tf32_ss(d_tmem, a_lo + 2 * k, w_hi + 2 * k, idesc, (kb | k) != 0);
            umma_tf32_ss(d_tmem, a_hi + 2 * k, w_lo + 2 * k, idesc, 1);
            umma_tf32_ss(d_tmem, a_hi + 2 * k, w_hi + 2 * k, idesc, 1);
          }
          umma_commit(&empty_bar[s]);
          if (kb == num_kb - 1) umma_commit(&acc_full[acc]);
        }
        __syncwarp();
        if (++s == C::kStages) s = 0, ph ^= 1;
      }
    }
  } else if (warp < 6) {
    // ===== splitters: A -> (hi in place, lo) =====
    const int t = threadIdx.x - 64;
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[s], ph);
        float4* hi = reinterpret_cast<float4*>(smem + s * C::kStageBytes);
        float4* lo = reinterpret_cast<float4*>(smem + s * C::kStageBytes + kL5ABytes);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int i = t + 128 * j;
          const float4 v = hi[i];
          float4 h, l;
          h.x = tf32_head(v.x), h.y = tf32_head(v.y), h.z = tf32_head(v.z), h.w = tf32_head(v.w);
          l.x = v.x - h.x, l.y = v.y - h.y, l.z = v.z - h.z, l.w = v.w - h.w;
          hi[i] = h;
          lo[i] = l;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&split_bar[s]);
        if (++s == C::kStages) s = 0, ph ^= 1;
      }
    }
  } else {
    // ===== epilogue: thread = (row, column half) =====
    const int quarter = warp & 3, half = (warp - 6) >> 2;
    constexpr int HW = BN / 2;
    const bool vec = (a.N & 3) == 0 && (a.ldy & 3) == 0 && (reinterpret_cast<uintptr_t>(a.Y) & 15) == 0 &&
                     (!a.res || ((a.ldr & 3) == 0 && (reinterpret_cast<uintptr_t>(a.res) & 15) == 0)) &&
                     (!a.rowadd || ((a.ldra & 3) == 0 && (reinterpret_cast<uintptr_t>(a.rowadd) & 15) == 0)) &&
                     (!a.bias || (reinterpret_cast<uintptr_t>(a.bias) & 15) == 0);
    int it = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const int m0 = (tile / n_tiles) * 128, n0 = (tile % n_tiles) * BN + half * HW;
      const int row = m0 + quarter * 32 + lane;
      mbar_wait(&acc_full[acc], (it >> 1) & 1);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + acc * C::kAccStride + half * HW + (static_cast<uint32_t>(quarter * 32) << 16);
      const float* res_row = a.res ? a.res + (size_t)row * a.ldr : nullptr;
      const float* add_row = a.rowadd ? a.rowadd + (size_t)(row % a.rowadd_period) * a.ldra : nullptr;
      float* y_row = a.Y + (size_t)row * a.ldy;
#pragma unroll 1
      for (int c = 0; c < HW; c += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + c, r);
        tmem_ld_wait();
        const int n = n0 + c;
        if (row >= a.rows || n >= a.N) continue;
        if (vec && n + 16 <= a.N) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
            if (a.bias) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + n + j));
              v.x += b.x, v.y += b.y, v.z += b.z, v.w += b.w;
            }
            if (a.gelu == 1) v.x = gelu_erf(v.x), v.y = gelu_erf(v.y), v.z = gelu_erf(v.z), v.w = gelu_erf(v.w);
            if (add_row) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(add_row + n + j));
              v.x += b.x, v.y += b.y, v.z += b.z, v.w += b.w;
            }
            if (res_row) {
              const float4 b = *reinterpret_cast<const float4*>(res_row + n + j);
              v.x += b.x, v.y += b.y, v.z += b.z, v.w += b.w;
            }
            if (a.gelu == 2) {
              v.x = v.x / (1.0f + expf(-v.x)), v.y = v.y / (1.0f + expf(-v.y));
              v.z = v.z / (1.0f + expf(-v.z)), v.w = v.w / (1.0f + expf(-v.w));
            }
            *reinterpret_cast<float4*>(y_row + n + j) = v;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (n + j >= a.N) continue;
            float v = __uint_as_float(r[j]) + (a.bias ? __ldg(a.bias + n + j) : 0.f);
            if (a.gelu == 1) v = gelu_erf(v);
            if (add_row) v += __ldg(add_row + n + j);
            if (res_row) v += res_row[n + j];
            if (a.gelu == 2) v = v / (1.0f + expf(-v));
            y_row[n + j] = v;
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

}  // namespace lam

// fp32-accurate linear layers of the first stage on the 5th-generation tensor cores: 3xTF32 with tcgen05.mma.kind::tf32.
//
//   Y[r, n] = epi( sum_k X[r, k] W[n, k] + bias[n] )       (same contract and epilogue options as LinearArgs / linear_f32_tc_kernel)
//
// x = hi + lo with hi = x rounded to TF32 and lo = x - hi (exact in fp32); x.w is accumulated as lo.hi + hi.lo + hi.hi in the fp32
// TMEM accumulator, the dropped lo.lo term is 2^-22 relative: as good as an fp32 FMA chain (the reference's first stage is fp32 and
// its outputs are checked to 1e-4; measured <= 3.8e-6 on the latents of all golden cases).  The weights are split once at *_create (w_hi / w_lo, fp32 arrays the tensor
// core reads as TF32); the activations are split on the fly in shared memory.
//
//   warp 0      TMA producer: per 32-wide k-block the raw fp32 A tile (128 rows x 128 B, 128-byte swizzle) and the W_hi / W_lo tiles
//   warps 2..5  splitters: rewrite the landed A tile as hi in place and write lo next to it (element-wise, so the swizzle is
//               irrelevant), fence.proxy.async, arrive
//   warp 1      MMA issuer: 3 tcgen05.mma (M128 x BN x K8) per 8-wide k step into one of two TMEM accumulators
//   warps 6..13 epilogue, thread = (row, column half): tcgen05.ld -> bias / GELU / row-add / residual / SiLU -> global fp32
//
// Persistent (one CTA per SM, tiles n-fastest so the CTAs that share an A tile run together and HBM sees it once); the epilogue of
// tile i overlaps the main loop of tile i + 1 through the second accumulator.  An optional LayerNorm of the result (the next
// sub-layer's pre-norm) runs in the epilogue: the finished values go back into the accumulator columns (tcgen05.st), a thread — or
// the two threads that share a row, through shared memory — merges shifted one-pass statistics, and a second pass over TMEM writes
// the normalised row.
// Measured (B200, 4AA first stage, 256 k entity rows / 128 k latent rows; scripts/gpu_l5var.sh switches parts of the kernel off):
//   the mma.sync 3xTF32 kernel this replaces reached ~100 - 120 TFLOP/s of tensor work (every fragment split in registers):
//   net_merge.2 [256 -> 256] 682 us, encoder.mlp.0 [384 -> 96] 466 us, decoder layers [128 -> 128] 195 us;
//   this kernel: 250 / 149 / 125 us in the step (194 / 142 / 93 us alone).  With the epilogue reduced to releasing the accumulator:
//   142 / 126 / 47 us; additionally without the splitters and the weight tiles (A stream + MMAs only): 127 / 111 / 44 us — i.e. the
//   A stream of 2 - 3 stages x 16 KB per SM sustains ~3 TB/s, and the epilogue costs as much again on the narrow layers.  Both are
//   latency, not bandwidth: tensor pipe 22 % busy, DRAM 25 %, LSU 30 %.  Next: weights resident in shared memory for the narrow
//   layers (frees the ring for 4 - 6 A stages), 16 epilogue warps.
#pragma once
#include <cuda.h>

#include "first_stage.cuh"
#include "gemm_ws.cuh"
#include "ptx.cuh"

namespace lam {

constexpr int kL5Threads = 32 * 14;
constexpr int kL5BK = 32;                       // fp32 per k-block = one 128-byte swizzle row
constexpr int kL5ABytes = 128 * kL5BK * 4;      // 16 KB

template <int BN>
struct L5Cfg {
  static constexpr int kBBytes = BN * kL5BK * 4;
  static constexpr int kStageBytes = 2 * kL5ABytes + 2 * kBBytes;  // A (-> hi), A lo, W hi, W lo
  static constexpr int kStagesRaw = (192 * 1024) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 4 ? 4 : kStagesRaw;
  static constexpr int kAccStride = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  static constexpr uint32_t kTmemCols = 2 * kAccStride;
  // stages | barriers | LN statistics [2 halves][128 rows] | staging boxes of the epilogue warps (8 x 4 KB)
  static constexpr int kSmem = kStages * kStageBytes + 256 + 2048 + 8 * 4096;
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
  extern __shared__ __align__(1024) uint8_t smem_l5[];
  uint8_t* smem = smem_l5;
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // the 128-byte swizzle atoms need 1024-byte aligned tiles
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* split_bar = full_bar + C::kStages;
  uint64_t* empty_bar = split_bar + C::kStages;
  uint64_t* acc_full = empty_bar + C::kStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float2* ln_stats = reinterpret_cast<float2*>(smem + C::kStages * C::kStageBytes + 256);
  uint8_t* stage_box = smem + C::kStages * C::kStageBytes + 256 + 2048;  // 8 epilogue warps x (output box | residual box) of 2 KB

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
          mbar_arrive_expect_tx(&full_bar[s], (a.debug & 8) ? kL5ABytes : kL5ABytes + 2 * C::kBBytes);
          tma_load_2d(&tm_a, &full_bar[s], st, kb * kL5BK, m0);
          if (!(a.debug & 8)) {
            tma_load_2d(&tm_whi, &full_bar[s], st + 2 * kL5ABytes, kb * kL5BK, n0);
            tma_load_2d(&tm_wlo, &full_bar[s], st + 2 * kL5ABytes + C::kBBytes, kb * kL5BK, n0);
          }
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
          if (a.debug & 4) break;
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
    // Global memory is touched in row-contiguous 64-byte segments only: a 32-row x 16-column box goes through the warp's staging
    // box in shared memory, written by thread = row and read back by lane = (row % 8, 16-byte piece), 8 rows per instruction (and the
    // other way round for the residual).  With thread = row accesses straight to global memory every instruction touched 32 rows
    // (32 L1 wavefronts for 512 bytes) and the LSU, not HBM, bounded the narrow layers.
    const int quarter = warp & 3, half = (warp - 6) >> 2;
    constexpr int HW = BN / 2;
    const bool ln = a.ln_out != nullptr;
    const bool vec = (a.N & 3) == 0 && (!a.Y || ((a.ldy & 3) == 0 && (reinterpret_cast<uintptr_t>(a.Y) & 15) == 0)) &&
                     (!a.res || ((a.ldr & 3) == 0 && (reinterpret_cast<uintptr_t>(a.res) & 15) == 0)) &&
                     (!a.rowadd || ((a.ldra & 3) == 0 && (reinterpret_cast<uintptr_t>(a.rowadd) & 15) == 0)) &&
                     (!a.bias || (reinterpret_cast<uintptr_t>(a.bias) & 15) == 0) &&
                     (!ln || ((a.ld_ln & 3) == 0 && (reinterpret_cast<uintptr_t>(a.ln_out) & 15) == 0 &&
                              (!a.ln_w || ((reinterpret_cast<uintptr_t>(a.ln_w) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.ln_b) & 15) == 0 &&
                                           (a.ln_group & 3) == 0))));
    const uint32_t box_st = smem_u32(stage_box) + (warp - 6) * 4096, box_rs = box_st + 2048;
    const int rr0 = lane >> 2, pc = lane & 3;  // coalesced mapping: rows rr0 + 8 i (i < 4), 16-byte piece pc of the 64-byte row segment
    int it = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const int m0 = (tile / n_tiles) * 128, n0 = (tile % n_tiles) * BN + half * HW;
      const int row0w = m0 + quarter * 32;
      const int row = row0w + lane;
      const uint32_t taddr = tmem_base + acc * C::kAccStride + half * HW + (static_cast<uint32_t>(quarter * 32) << 16);
      const float* res_row = a.res ? a.res + (size_t)(a.res_idx ? (row < a.rows ? a.res_idx[row] : 0) : row) * a.ldr : nullptr;
      const float* add_row = a.rowadd ? a.rowadd + (size_t)(row % a.rowadd_period) * a.ldra : nullptr;
      float* y_row = a.Y ? a.Y + (size_t)row * a.ldy : nullptr;
      // rows of the coalesced mapping
      bool rok[4];
      const float* rsrc[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rw = row0w + rr0 + 8 * i;
        rok[i] = rw < a.rows;
        rsrc[i] = (a.res && rok[i]) ? a.res + (size_t)(a.res_idx ? a.res_idx[rw] : rw) * a.ldr + pc * 4 : nullptr;
      }
      float4 rp[4];
      auto res_fetch = [&](int n) {
#pragma unroll
        for (int i = 0; i < 4; ++i) rp[i] = rsrc[i] ? *reinterpret_cast<const float4*>(rsrc[i] + n) : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      // bias of a chunk: fetched one chunk ahead (with ~220 KB of shared memory per CTA there is no L1 to speak of: every fetch is an
      // L2 round trip that would otherwise sit between the accumulator load and the first add of every chunk)
      float4 bp[4];
      auto bias_fetch = [&](int n) {
        const bool ok = a.bias && vec && n + 16 <= a.N;
#pragma unroll
        for (int j = 0; j < 4; ++j) bp[j] = ok ? __ldg(reinterpret_cast<const float4*>(a.bias + n) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      if (a.res && vec && n0 + 16 <= a.N) res_fetch(n0);  // in flight while the accumulator is being finished
      bias_fetch(n0);
      mbar_wait(&acc_full[acc], (it >> 1) & 1);
      tcgen05_fence_after();
      float kshift = 0.f, s1 = 0.f, s2 = 0.f;  // shifted one-pass statistics of this thread's columns (LN mode)
      int cnt = 0;
      uint32_t rn[16];
      tmem_ld16(taddr, rn);
#pragma unroll 1
      for (int c = 0; c < ((a.debug & 2) ? 0 : HW); c += 16) {
        uint32_t r[16];
        const int n = n0 + c;
        const bool full = vec && n + 16 <= a.N;  // warp-uniform
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = rn[j];
        if (c + 16 < HW) tmem_ld16(taddr + c + 16, rn);  // next chunk's accumulator columns in flight under this chunk's work
        if (full && a.res) {
          __syncwarp();  // the previous chunk's reads of the residual box are done
#pragma unroll
          for (int i = 0; i < 4; ++i)
            st_shared_v4(box_rs + stage_off64(rr0 + 8 * i, pc), __float_as_uint(rp[i].x), __float_as_uint(rp[i].y), __float_as_uint(rp[i].z),
                         __float_as_uint(rp[i].w));
          __syncwarp();
          if (c + 16 < HW && n + 32 <= a.N) res_fetch(n + 16);
        }
        if (full) {
          if (a.Y) __syncwarp();  // the previous chunk's reads of the output box are done
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
            {
              const float4 b = bp[j >> 2];
              v.x += b.x, v.y += b.y, v.z += b.z, v.w += b.w;
              if (j == 12 && c + 16 < HW) bias_fetch(n + 16);  // the next chunk's bias: in flight under the rest of this chunk
            }
            if (a.gelu == 1) v.x = gelu_erf(v.x), v.y = gelu_erf(v.y), v.z = gelu_erf(v.z), v.w = gelu_erf(v.w);
            if (add_row) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(add_row + n + j));
              v.x += b.x, v.y += b.y, v.z += b.z, v.w += b.w;
            }
            if (a.res) {
              const float4 b = ld_shared_f4(box_rs + stage_off64(lane, j >> 2));
              v.x += b.x, v.y += b.y, v.z += b.z, v.w += b.w;
            }
            if (a.gelu == 2) {
              v.x = v.x / (1.0f + expf(-v.x)), v.y = v.y / (1.0f + expf(-v.y));
              v.z = v.z / (1.0f + expf(-v.z)), v.w = v.w / (1.0f + expf(-v.w));
            }
            if (a.Y) st_shared_v4(box_st + stage_off64(lane, j >> 2), __float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w));
            if (ln) {
              if (cnt == 0 && j == 0) kshift = v.x;
              const float d0 = v.x - kshift, d1 = v.y - kshift, d2 = v.z - kshift, d3 = v.w - kshift;
              s1 += (d0 + d1) + (d2 + d3);
              s2 = fmaf(d0, d0, s2), s2 = fmaf(d1, d1, s2), s2 = fmaf(d2, d2, s2), s2 = fmaf(d3, d3, s2);
              r[j] = __float_as_uint(v.x), r[j + 1] = __float_as_uint(v.y), r[j + 2] = __float_as_uint(v.z), r[j + 3] = __float_as_uint(v.w);
            }
          }
          if (ln) cnt += 16;
          if (a.Y) {
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 w = ld_shared_f4(box_st + stage_off64(rr0 + 8 * i, pc));
              if (rok[i] && !(a.debug & 1)) *reinterpret_cast<float4*>(a.Y + (size_t)(row0w + rr0 + 8 * i) * a.ldy + n + pc * 4) = w;
            }
          }
        } else if (row < a.rows && n < a.N) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (n + j >= a.N) continue;
            float v = __uint_as_float(r[j]) + (a.bias ? __ldg(a.bias + n + j) : 0.f);
            if (a.gelu == 1) v = gelu_erf(v);
            if (add_row) v += __ldg(add_row + n + j);
            if (res_row) v += res_row[n + j];
            if (a.gelu == 2) v = v / (1.0f + expf(-v));
            if (y_row) y_row[n + j] = v;
            if (ln) {
              if (cnt == 0) kshift = v;
              const float d = v - kshift;
              s1 += d, s2 = fmaf(d, d, s2), ++cnt;
              r[j] = __float_as_uint(v);
            }
          }
        }
        if (ln) tmem_st16(taddr + c, r);  // the finished values go back into the accumulator columns for the second pass
      }
      if (a.debug & 2) tmem_ld_wait();
      if (ln) {
        // ---- LayerNorm of the values just produced.  Group = this thread's column half: local statistics.  Group = the row: the two
        // halves merge (mean, M2, count) through shared memory (Chan et al.).
        tmem_st_wait();
        float mean = 0.f, m2 = 0.f;
        if (cnt > 0) {
          mean = kshift + s1 / (float)cnt;
          m2 = fmaxf(s2 - s1 * s1 / (float)cnt, 0.f);
        }
        float tot = (float)cnt;
        if (a.ln_group == 0) {
          ln_stats[half * 128 + quarter * 32 + lane] = make_float2(mean, m2);
          asm volatile("bar.sync 1, 256;" ::: "memory");
          const float2 o = ln_stats[(half ^ 1) * 128 + quarter * 32 + lane];
          const int n_other = half == 0 ? min(max(a.N - HW, 0), HW) : min(a.N, HW);  // single n-tile: the other half's column count
          const float ca = (float)cnt, cb = (float)n_other;
          tot = ca + cb;
          if (tot > 0.f) {
            const float delta = o.x - mean;
            const float new_mean = mean + delta * (cb / tot);
            m2 = m2 + o.y + delta * delta * (ca * cb / tot);
            mean = new_mean;
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");  // both halves have read before the next tile's statistics are written
        }
        const float rstd = rsqrtf(m2 / fmaxf(tot, 1.f) + a.ln_eps);
        float* o_row = a.ln_out + (size_t)row * a.ld_ln;
        uint32_t r0buf[16];
        tmem_ld16(taddr, r0buf);
#pragma unroll 1
        for (int c = 0; c < HW; c += 16) {
          uint32_t r[16];
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) r[j] = r0buf[j];
          if (c + 16 < HW) tmem_ld16(taddr + c + 16, r0buf);
          const int n = n0 + c;
          const bool full = vec && n + 16 <= a.N;
          const int wbase = a.ln_group ? n % a.ln_group : n;  // 16-column chunks never straddle a group (group % 16 == 0 or one chunk)
          if (full) {
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              float4 y;
              y.x = (__uint_as_float(r[j]) - mean) * rstd, y.y = (__uint_as_float(r[j + 1]) - mean) * rstd;
              y.z = (__uint_as_float(r[j + 2]) - mean) * rstd, y.w = (__uint_as_float(r[j + 3]) - mean) * rstd;
              if (a.ln_w) {
                const float4 w = __ldg(reinterpret_cast<const float4*>(a.ln_w + wbase + j)), b = __ldg(reinterpret_cast<const float4*>(a.ln_b + wbase + j));
                y.x = fmaf(y.x, w.x, b.x), y.y = fmaf(y.y, w.y, b.y), y.z = fmaf(y.z, w.z, b.z), y.w = fmaf(y.w, w.w, b.w);
              }
              st_shared_v4(box_st + stage_off64(lane, j >> 2), __float_as_uint(y.x), __float_as_uint(y.y), __float_as_uint(y.z), __float_as_uint(y.w));
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 w = ld_shared_f4(box_st + stage_off64(rr0 + 8 * i, pc));
              if (rok[i]) *reinterpret_cast<float4*>(a.ln_out + (size_t)(row0w + rr0 + 8 * i) * a.ld_ln + n + pc * 4) = w;
            }
          } else if (row < a.rows && n < a.N) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (n + j >= a.N) continue;
              const float t = (__uint_as_float(r[j]) - mean) * rstd;
              o_row[n + j] = a.ln_w ? fmaf(t, __ldg(a.ln_w + wbase + j), __ldg(a.ln_b + wbase + j)) : t;
            }
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

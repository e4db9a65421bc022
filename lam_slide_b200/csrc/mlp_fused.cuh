// Fused "second half" of a ParallelMLPAttentionV2 block (mmdit.py:241-248, latent_si_v31.py:54,61):
//
//     h[rows, H] += gate[b] * ( [ attn | gelu(u W1m^T + b1m) ] W2^T + b2 )
//
// i.e. the MLP half of linear1, the GELU and ALL of linear2 in one persistent tcgen05 kernel: the 4H-wide MLP activation lives
// only in TMEM / shared memory (128 hidden columns at a time) and never goes to HBM — per 128 k tokens that removes 393 MB of
// writes and 393 MB of reads from a block that is otherwise HBM-bound in linear2.  W1m = linear1.weight[3H:], the attention
// output `attn` [rows, H] bf16 comes from the attention kernels, `u` is the LN + modulate output (A operand of linear1).
//
// One CTA per SM, m-blocks of 128 rows.  TMEM (512 columns): OUT accumulator 128 x H fp32 (H <= 384 columns) | ACC1 128 x 128.
// Shared memory: u tile resident (H/64 k-blocks of 16 KB), G = two 128 x 64 bf16 tiles (A operand of the second GEMM: first the
// attention tile k-blocks by TMA, then the GELU output of each hidden chunk written by the epilogue warps in the 128-byte
// swizzled layout), a ring of 16 KB weight units ([128 rows x 64 k] of W1m or W2), constants, barriers.
// Per m-block the MMA thread issues:
//     A-phase :  OUT  = attn W2[:, :H]^T                       (H/64 k-blocks x H/128 units)
//     B-phase :  for each hidden chunk j (128 columns):  ACC1 = u W1m_j^T  (G1, H/64 units);   OUT += gelu(ACC1 + b) W2[:, H+128j..]^T  (G2, 2 x H/128 units)
//                issued as G1(0) G1(1) G2(0) G1(2) G2(1) ... so the GELU of chunk j (16 epilogue warps, registers) overlaps G1(j+1).
// Warp roles as in gemm_ws.cuh: warp 0 TMA producer, warp 1 MMA issuer (both warp-uniform, elect around the issue), warps 2..17
// epilogue (TMEM lane quarter = warp % 4, column quarter = (warp - 2) / 4).
#pragma once
#include "gemm_ws.cuh"

namespace lam {

struct FusedMlpParams {
  const float* b1m;   // [M]   linear1.bias[3H:]
  const float* b2;    // [H]   linear2.bias
  const float* gate;  // gate of sample b at gate + b * gate_stride, [H]
  int gate_stride;
  int rows_per_sample;  // T * L
  int H, M, rows;
};

struct FusedMlpSmem {
  int u_bytes, g_bytes, ring_bytes, const_bytes, bar_bytes, total;
};
static inline __host__ __device__ FusedMlpSmem fused_mlp_smem(int H, int M, int stages) {
  FusedMlpSmem s;
  s.u_bytes = (H / 64) * 16384;
  s.g_bytes = 2 * 16384;
  s.ring_bytes = stages * 16384;
  s.const_bytes = ((M + H) * 4 + 15) / 16 * 16;
  s.bar_bytes = 512;
  s.total = s.u_bytes + s.g_bytes + s.ring_bytes + s.const_bytes + s.bar_bytes;
  return s;
}

__global__ void __launch_bounds__(kWsThreads, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tmap_u, const __grid_constant__ CUtensorMap tmap_attn,
                 const __grid_constant__ CUtensorMap tmap_w1, const __grid_constant__ CUtensorMap tmap_w2,
                 const __grid_constant__ CUtensorMap tmap_h, int num_m_blocks, int stages, FusedMlpParams p) {
  constexpr int kUnit = 16384;  // one [128 x 64] bf16 tile
  const int H = p.H, M = p.M;
  const int KB = H / 64;    // k-blocks of u / attn
  const int NI = H / 128;   // 128-row units of W2 (output column groups)
  const int NJ = M / 128;   // hidden chunks

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const FusedMlpSmem plan = fused_mlp_smem(H, M, stages);
  uint8_t* u_res = smem;
  uint8_t* g_buf = u_res + plan.u_bytes;   // two 16 KB tiles (also the staging area of the final epilogue)
  uint8_t* ring = g_buf + plan.g_bytes;
  float* smf = reinterpret_cast<float*>(ring + plan.ring_bytes);  // [M] b1m | [H] b2
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(smf) + plan.const_bytes);
  uint64_t* full_bar = bars;         // [8] ring
  uint64_t* empty_bar = bars + 8;    // [8]
  uint64_t* a_full = bars + 16;      // [8] u k-blocks
  uint64_t* a_empty = bars + 24;     // [8]
  uint64_t* ga_full = bars + 32;     // [2] attention tile in G slot (TMA)
  uint64_t* gg_full = bars + 34;     // [2] GELU output in G slot (4 epilogue warps each... see counts below)
  uint64_t* g_empty = bars + 36;     // [2] G slot consumed by the MMAs
  uint64_t* acc1_full = bars + 38;
  uint64_t* acc1_empty = bars + 39;
  uint64_t* out_full = bars + 40;
  uint64_t* out_free = bars + 41;
  uint64_t* stage_free = bars + 42;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 43);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_u);
    tma_prefetch_desc(&tmap_attn);
    tma_prefetch_desc(&tmap_w1);
    tma_prefetch_desc(&tmap_w2);
    for (int s = 0; s < 8; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ga_full[s], 1);
      mbar_init(&gg_full[s], kWsEpiWarps / 2);  // the 8 warps whose hidden columns fall into this 64-column tile
      mbar_init(&g_empty[s], 1);
    }
    mbar_init(acc1_full, 1);
    mbar_init(acc1_empty, kWsEpiWarps);
    mbar_init(out_full, 1);
    mbar_init(out_free, kWsEpiWarps);
    mbar_init(stage_free, kWsEpiWarps);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < M; i += kWsThreads - 64) smf[i] = p.b1m[i];
    for (int i = threadIdx.x - 64; i < H; i += kWsThreads - 64) smf[M + i] = p.b2[i];
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_out = tmem_base;            // H columns
  const uint32_t tmem_acc1 = tmem_base + 384;     // 128 columns

  if (warp == 0) {
    // ===== TMA producer: u k-blocks (once per m-block), attention tiles into G, weight units into the ring — in exactly the
    // order the MMA warp consumes them =====
    int s = 0;
    uint32_t ph = 0, it = 0, g_use0 = 0, g_use1 = 0;  // writes into G slot 0 / 1 so far (TMA + epilogue)
    auto unit = [&](const CUtensorMap* tm, int row, int col) {
      mbar_wait(&empty_bar[s], ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&full_bar[s], kUnit);
        tma_load_2d(tm, &full_bar[s], ring + s * kUnit, col, row);
      }
      __syncwarp();
      if (++s == stages) s = 0, ph ^= 1;
    };
    for (int mb = blockIdx.x; mb < num_m_blocks; mb += gridDim.x, ++it) {
      const int m0 = mb * kBlockM;
      // the final epilogue of the previous m-block stages its output boxes in G: wait until it is done with them
      mbar_wait(stage_free, (it & 1) ^ 1);
      // A-phase: attention tile k-block kk -> G slot kk & 1, then its NI weight units
      for (int kk = 0; kk < KB; ++kk) {
        const int slot = kk & 1;
        mbar_wait(&g_empty[slot], ((slot ? g_use1 : g_use0) & 1) ^ 1);
        if (slot) ++g_use1;
        else ++g_use0;
        if (elect_one()) {
          mbar_arrive_expect_tx(&ga_full[slot], kUnit);
          tma_load_2d(&tmap_attn, &ga_full[slot], g_buf + slot * kUnit, kk * 64, m0);
        }
        __syncwarp();
        for (int i = 0; i < NI; ++i) unit(&tmap_w2, i * 128, kk * 64);
        if (kk == 0) {  // u for this m-block (its buffers are released by the last G1 of the previous m-block)
          for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(&a_empty[kb], (it & 1) ^ 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(&a_full[kb], kUnit);
              tma_load_2d(&tmap_u, &a_full[kb], u_res + kb * kUnit, kb * 64, m0);
            }
            __syncwarp();
          }
        }
      }
      // B-phase (the G slots are now written by the epilogue warps: count their uses so the parities stay in step)
      auto g1_units = [&](int j) {
        for (int kb = 0; kb < KB; ++kb) unit(&tmap_w1, 3 * H + j * 128, kb * 64);
      };
      auto g2_units = [&](int j) {
        for (int kk = 0; kk < 2; ++kk)
          for (int i = 0; i < NI; ++i) unit(&tmap_w2, i * 128, H + j * 128 + kk * 64);
      };
      g1_units(0);
      for (int j = 0; j < NJ; ++j) {
        if (j + 1 < NJ) g1_units(j + 1);
        g2_units(j);
        ++g_use0, ++g_use1;
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
    int s = 0;
    uint32_t ph = 0, it = 0, n_acc1 = 0, ga_use0 = 0, ga_use1 = 0, gg_use0 = 0, gg_use1 = 0;
    // 4 MMAs (one 64-wide k-block): D[tmem] (+)= A[a_addr] * B[ring slot]^T; frees the ring slot
    auto mma_unit = [&](uint32_t d_tmem, uint32_t a_addr, bool first_zero) {
      mbar_wait(&full_bar[s], ph);
      tcgen05_fence_after();
      const uint64_t a_desc = umma_desc_sw128(a_addr);
      const uint64_t b_desc = umma_desc_sw128(smem_u32(ring + s * kUnit));
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (first_zero && k == 0) ? 0u : 1u);
        umma_commit(&empty_bar[s]);
      }
      __syncwarp();
      if (++s == stages) s = 0, ph ^= 1;
    };
    for (int mb = blockIdx.x; mb < num_m_blocks; mb += gridDim.x, ++it) {
      mbar_wait(out_free, (it & 1) ^ 1);  // the previous m-block's OUT has been drained
      tcgen05_fence_after();
      // A-phase
      for (int kk = 0; kk < KB; ++kk) {
        const int slot = kk & 1;
        mbar_wait(&ga_full[slot], (slot ? ga_use1 : ga_use0) & 1);
        if (slot) ++ga_use1;
        else ++ga_use0;
        tcgen05_fence_after();
        for (int i = 0; i < NI; ++i) mma_unit(tmem_out + i * 128, smem_u32(g_buf + slot * kUnit), kk == 0);
        if (elect_one()) umma_commit(&g_empty[slot]);
        __syncwarp();
      }
      // B-phase
      auto g1 = [&](int j) {
        mbar_wait(acc1_empty, (n_acc1 & 1) ^ 1);
        tcgen05_fence_after();
        for (int kb = 0; kb < KB; ++kb) {
          if (j == 0) mbar_wait(&a_full[kb], it & 1);
          mma_unit(tmem_acc1, smem_u32(u_res + kb * kUnit), kb == 0);
          if (j == NJ - 1) {
            if (elect_one()) umma_commit(&a_empty[kb]);
            __syncwarp();
          }
        }
        if (elect_one()) umma_commit(acc1_full);
        __syncwarp();
        ++n_acc1;
      };
      auto g2 = [&](int j) {
        for (int kk = 0; kk < 2; ++kk) {
          mbar_wait(&gg_full[kk], (kk ? gg_use1 : gg_use0) & 1);
          if (kk) ++gg_use1;
          else ++gg_use0;
          tcgen05_fence_after();
          for (int i = 0; i < NI; ++i) mma_unit(tmem_out + i * 128, smem_u32(g_buf + kk * kUnit), false);
          if (elect_one()) umma_commit(&g_empty[kk]);
          __syncwarp();
        }
        if (j == NJ - 1) {
          if (elect_one()) umma_commit(out_full);
          __syncwarp();
        }
      };
      g1(0);
      for (int j = 0; j < NJ; ++j) {
        if (j + 1 < NJ) g1(j + 1);
        g2(j);
      }
    }
  } else {
    // ===== epilogue warps =====
    const int q = warp & 3;
    const int cq = (warp - 2) >> 2;
    const uint32_t lane_t = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t smf_s = smem_u32(smf);
    const uint32_t g_s = smem_u32(g_buf);
    const int r_in_tile = q * 32 + lane;
    uint32_t it = 0, n_acc1 = 0, g_use_mine = 0;  // writes so far into the G tile this warp writes (TMA + epilogue)
    for (int mb = blockIdx.x; mb < num_m_blocks; mb += gridDim.x, ++it) {
      const int row0 = mb * kBlockM + q * 32;
      const int row = row0 + lane;
      // ---- hidden chunks: ACC1 -> +bias -> GELU -> bf16 -> G (A operand layout: 128-byte swizzle, K-major)
      const int kk = cq >> 1;  // which 64-column tile of the chunk this warp's 32 columns fall into
      g_use_mine += kk ? KB / 2 : (KB + 1) / 2;  // the A-phase's (TMA) writes into that tile
      for (int j = 0; j < NJ; ++j, ++n_acc1) {
        mbar_wait(acc1_full, n_acc1 & 1);
        tcgen05_fence_after();
        uint32_t v[32];
        tmem_ld32(lane_t + 384 + cq * 32, v);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc1_empty);
        uint32_t w[16];
        const uint32_t bias_s = smf_s + (j * 128 + cq * 32) * 4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 bv = ld_shared_f4(bias_s + i * 16);
          const float y0 = gelu_fast(__uint_as_float(v[4 * i + 0]) + bv.x);
          const float y1 = gelu_fast(__uint_as_float(v[4 * i + 1]) + bv.y);
          const float y2 = gelu_fast(__uint_as_float(v[4 * i + 2]) + bv.z);
          const float y3 = gelu_fast(__uint_as_float(v[4 * i + 3]) + bv.w);
          w[2 * i] = pack_bf16x2(y0, y1);
          w[2 * i + 1] = pack_bf16x2(y2, y3);
        }
        mbar_wait(&g_empty[kk], (g_use_mine & 1) ^ 1);  // the MMAs that read the previous content of this tile are done
        ++g_use_mine;
        const uint32_t tile = g_s + kk * kUnit + r_in_tile * 128;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int chunk = (cq & 1) * 4 + c;
          st_shared_v4(tile + ((chunk ^ (r_in_tile & 7)) << 4), w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
        }
        fence_proxy_async();  // generic-proxy writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&gg_full[kk]);
      }
      // ---- OUT: h += gate * (acc + b2) as TMA f32 reduce-add boxes of 16 columns, staged in G (free once out_full fired)
      mbar_wait(out_full, it & 1);
      tcgen05_fence_after();
      const int b = (row < p.rows ? row : p.rows - 1) / p.rows_per_sample;
      const float* gate = p.gate + (size_t)b * p.gate_stride;
      const int wcols = H / 4;  // this warp's output columns
      const uint32_t stage_s = g_s + (warp - 2) * 2048;
      for (int bx = 0; bx < wcols / 16; ++bx) {
        const int col = cq * wcols + bx * 16;
        uint32_t v[16];
        tmem_ld16(lane_t + col, v);
        tmem_ld_wait();
        if (bx == wcols / 16 - 1) {  // OUT is in registers: the next m-block's A-phase may overwrite it
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(out_free);
        }
        float4 o[4];
#pragma unroll
        for (int jx = 0; jx < 4; ++jx) {
          const float4 gv = __ldg(reinterpret_cast<const float4*>(gate + col) + jx);
          const float4 bv = ld_shared_f4(smf_s + (M + col) * 4 + jx * 16);
          o[jx].x = gv.x * (__uint_as_float(v[4 * jx + 0]) + bv.x);
          o[jx].y = gv.y * (__uint_as_float(v[4 * jx + 1]) + bv.y);
          o[jx].z = gv.z * (__uint_as_float(v[4 * jx + 2]) + bv.z);
          o[jx].w = gv.w * (__uint_as_float(v[4 * jx + 3]) + bv.w);
        }
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
          st_shared_v4(stage_s + stage_off64(lane, ch), __float_as_uint(o[ch].x), __float_as_uint(o[ch].y), __float_as_uint(o[ch].z),
                       __float_as_uint(o[ch].w));
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_reduce_add_2d_s(&tmap_h, stage_s, col, row0);
          bulk_commit();
        }
      }
      if (lane == 0) {
        bulk_wait_read<0>();
        mbar_arrive(stage_free);  // G is free for the next m-block's attention tiles
      }
      __syncwarp();
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace lam

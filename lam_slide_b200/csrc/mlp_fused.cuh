// Fused "second half" of a ParallelMLPAttentionV2 block (mmdit.py:241-248, latent_si_v31.py:54,61):
//
//     h[rows, H] += gate[b] * ( [ attn | gelu(u W1m^T + b1m) ] W2^T + b2 )
//
// i.e. the MLP half of linear1, the GELU and ALL of linear2 in one persistent tcgen05 kernel: the 4H-wide MLP activation lives
// only in TMEM / shared memory (128 hidden columns at a time) and never goes to HBM — per 128 k tokens that removes 393 MB of
// writes and 393 MB of reads from a block that is otherwise HBM-bound in linear2.  W1m = linear1.weight[3H:], the attention
// output `attn` [rows, H] bf16 comes from the attention kernels, `u` is the LN + modulate output (A operand of linear1).
//
// CTA pairs (2-CTA clusters, tcgen05 cta_group::2, M = 256): each CTA owns 128 rows (its u tile, its G tiles, its half of TMEM)
// and stages HALF of every weight unit, so per SM the weight bytes crossing the L2 -> SM port are halved (one CTA per m-block
// needs 64 B/clk/SM of weights at full tensor rate; the L2 delivers ~42).
// TMEM (512 columns): OUT accumulator 128 x H fp32 (H <= 384 columns) | ACC1 128 x 128 at column 384.
// Shared memory: u tile resident (H/64 k-blocks of 16 KB), G = two 128 x 64 bf16 tiles (A operand of the second GEMM: first the
// attention tile k-blocks by TMA, then the GELU output of each hidden chunk written by the epilogue warps in the 128-byte
// swizzled layout), ring 1 of W1m half-units [64 rows x 64 k], ring 2 of W2 half-units [NU/2 rows x 64 k], constants, barriers.
//
// Two independent MMA streams, each with its own producer warp, ring and issuing warp (measured, scripts/issue_bench.cu: the
// wait / elect / 4 x UTCHMMA / commit sequence costs ~350 cycles of the issuing warp per unit, more than the 256 cycles an
// N = 128 unit occupies the tensor pipe — one issuer for both streams left the pipe at 40 %):
//     stream 1 (warps 0,1):  for each hidden chunk j (128 columns):  ACC1 = u W1m_j^T                (H/64 units, N = 128)
//     stream 2 (warps 2,3):  OUT = attn W2[:, :H]^T ;  then per chunk j:  OUT += gelu(ACC1 + b1m_j) W2[:, H+128j ..]^T
//                            (units of NU = 192 output columns for H = 384, else min(H, 256))
// Warps 4..19: epilogue (TMEM lane quarter = warp % 4, column quarter = (warp - 4) / 4): ACC1 -> + bias -> GELU -> bf16 -> G;
// at the end of the m-block OUT -> gate * (acc + b2) -> TMA f32 reduce-add into h (staged in G).
// Barrier protocol in the pair (same offsets in both CTAs): "full" barriers live on the leader (both producers' TMA loads
// complete_tx there, the leader arms 2x the bytes; both CTAs' epilogue warps arrive there), "empty" barriers are per CTA (the
// leader's issuing threads commit to both CTAs).
#pragma once
#include "gemm_ws.cuh"

namespace lam {

constexpr int kFusedThreads = 128 + 32 * kWsEpiWarps;  // 2 producer + 2 issuer + 16 epilogue warps

struct FusedMlpParams {
  const float* b1m;   // [M]   linear1.bias[3H:]
  const float* b2;    // [H]   linear2.bias
  const float* gate;  // gate of sample b at gate + b * gate_stride, [H]
  int gate_stride;
  int rows_per_sample;  // T * L
  int H, M, rows;
  int debug;  // profiling aid: 1 skip the GELU math, 2 skip the output reduce-add, 4 skip the G tile writes, 8 skip the weight TMA loads
};

// output columns per stream-2 MMA (N of the instruction; each CTA of the pair stages N/2 weight rows)
static inline __host__ __device__ int fused_mlp_out_unit(int H) { return H % 192 == 0 && H > 256 ? 192 : (H <= 256 ? H : 128); }

struct FusedMlpSmem {
  int u_bytes, g_bytes, ring1_bytes, ring2_bytes, const_bytes, bar_bytes, total;
};
static inline __host__ __device__ FusedMlpSmem fused_mlp_smem(int H, int M, int stages1, int stages2) {
  FusedMlpSmem s;
  s.u_bytes = (H / 64) * 16384;
  s.g_bytes = 2 * 16384;
  s.ring1_bytes = stages1 * 8192;
  s.ring2_bytes = (stages2 * fused_mlp_out_unit(H) * 64 + 1023) / 1024 * 1024;
  s.const_bytes = ((M + H) * 4 + 15) / 16 * 16;
  s.bar_bytes = 512;
  s.total = s.u_bytes + s.g_bytes + s.ring1_bytes + s.ring2_bytes + s.const_bytes + s.bar_bytes;
  return s;
}

__global__ void __launch_bounds__(kFusedThreads, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tmap_u, const __grid_constant__ CUtensorMap tmap_attn,
                 const __grid_constant__ CUtensorMap tmap_w1, const __grid_constant__ CUtensorMap tmap_w2,
                 const __grid_constant__ CUtensorMap tmap_h, int num_m_blocks, int stages1, int stages2, FusedMlpParams p) {
  constexpr int kTile = 16384;   // one [128 x 64] bf16 tile (u k-block, G tile)
  constexpr int kUnit1 = 8192;   // this CTA's half of a W1m unit: [64 rows x 64 k]
  const int H = p.H, M = p.M;
  const int KB = H / 64;               // k-blocks of u / attn
  const int NU = fused_mlp_out_unit(H);
  const int NI = H / NU;               // stream-2 units per k-block
  const int NJ = M / 128;              // hidden chunks
  const int kUnit2 = NU * 64;          // this CTA's half of a W2 unit: [NU/2 rows x 64 k]

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const FusedMlpSmem plan = fused_mlp_smem(H, M, stages1, stages2);
  uint8_t* u_res = smem;
  uint8_t* g_buf = u_res + plan.u_bytes;   // two 16 KB tiles (also the staging area of the final epilogue)
  uint8_t* ring1 = g_buf + plan.g_bytes;
  uint8_t* ring2 = ring1 + plan.ring1_bytes;
  float* smf = reinterpret_cast<float*>(ring2 + plan.ring2_bytes);  // [M] b1m | [H] b2
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(smf) + plan.const_bytes);
  uint64_t* full1 = bars;            // [8] ring 1                      (leader)
  uint64_t* empty1 = bars + 8;       // [8]                             (per CTA)
  uint64_t* full2 = bars + 16;       // [8] ring 2                      (leader)
  uint64_t* empty2 = bars + 24;      // [8]                             (per CTA)
  uint64_t* a_full = bars + 32;      // [8] u k-blocks                  (leader)
  uint64_t* a_empty = bars + 40;     // [8]                             (per CTA)
  uint64_t* ga_full = bars + 48;     // [2] attention tile in G slot    (leader; TMA of both CTAs)
  uint64_t* gg_full = bars + 50;     // [2] GELU output in G slot       (leader; 8 warps of each CTA)
  uint64_t* g_empty = bars + 52;     // [2] G slot consumed by the MMAs (per CTA)
  uint64_t* acc1_full = bars + 54;   //                                 (per CTA)
  uint64_t* acc1_empty = bars + 55;  //                                 (leader; 16 warps of each CTA)
  uint64_t* out_full = bars + 56;    //                                 (per CTA)
  uint64_t* out_free = bars + 57;    //                                 (leader; 16 warps of each CTA)
  uint64_t* stage_free = bars + 58;  // G no longer used as staging     (per CTA; its 16 epilogue warps)
  uint64_t* attn_done = bars + 59;   // OUT = attn W2a^T has consumed the last attention tile: G belongs to the GELU chunks (per CTA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 60);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta_rank = (int)cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int m_first = (blockIdx.x / 2) * 2;
  const int m_step = (gridDim.x / 2) * 2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_u);
    tma_prefetch_desc(&tmap_attn);
    tma_prefetch_desc(&tmap_w1);
    tma_prefetch_desc(&tmap_w2);
    for (int s = 0; s < 8; ++s) {
      mbar_init(&full1[s], 1);
      mbar_init(&empty1[s], 1);
      mbar_init(&full2[s], 1);
      mbar_init(&empty2[s], 1);
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ga_full[s], 1);
      mbar_init(&gg_full[s], 2 * (kWsEpiWarps / 2));  // the 8 warps of each CTA whose hidden columns fall into this 64-column tile
      mbar_init(&g_empty[s], 1);
    }
    mbar_init(acc1_full, 1);
    mbar_init(acc1_empty, 2 * kWsEpiWarps);
    mbar_init(out_full, 1);
    mbar_init(out_free, 2 * kWsEpiWarps);
    mbar_init(stage_free, kWsEpiWarps);
    mbar_init(attn_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair<512>(tmem_slot);
  if (warp >= 4) {
    for (int i = threadIdx.x - 128; i < M; i += kFusedThreads - 128) smf[i] = p.b1m[i];
    for (int i = threadIdx.x - 128; i < H; i += kFusedThreads - 128) smf[M + i] = p.b2[i];
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before any TMA completion / remote arrive
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_out = tmem_base;            // H columns
  const uint32_t tmem_acc1 = tmem_base + 384;     // 128 columns

  if (warp == 0) {
    // ===== producer 1: u k-blocks (once per m-block) and W1m half-units, in the order stream 1 consumes them =====
    int s = 0;
    uint32_t ph = 0, it = 0;
    for (int mbase = m_first; mbase < num_m_blocks; mbase += m_step, ++it) {
      const int m0 = (mbase + cta_rank) * kBlockM;  // may be past the end in the last sweep: the TMA zero-fills, nothing is stored
      for (int j = 0; j < NJ; ++j) {
        for (int kb = 0; kb < KB; ++kb) {
          if (j == 0) {  // u k-block kb: its buffer is released by the last chunk of the previous m-block
            mbar_wait(&a_empty[kb], (it & 1) ^ 1);
            if (elect_one()) {
              if (leader) mbar_arrive_expect_tx(&a_full[kb], 2 * kTile);
              tma_load_2d_pair(&tmap_u, mapa_u32(smem_u32(&a_full[kb]), 0), u_res + kb * kTile, kb * 64, m0);
            }
            __syncwarp();
          }
          mbar_wait(&empty1[s], ph ^ 1);
          if (elect_one()) {
            if (p.debug & 8) {
              if (leader) mbar_arrive(&full1[s]);
            } else {
              if (leader) mbar_arrive_expect_tx(&full1[s], 2 * kUnit1);
              tma_load_2d_pair(&tmap_w1, mapa_u32(smem_u32(&full1[s]), 0), ring1 + s * kUnit1, kb * 64, 3 * H + j * 128 + cta_rank * 64);
            }
          }
          __syncwarp();
          if (++s == stages1) s = 0, ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===== issuer 1 (leader): ACC1 = u W1m_j^T for every hidden chunk =====
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, 128);
      int s = 0;
      uint32_t ph = 0, it = 0, n_acc1 = 0;
      for (int mbase = m_first; mbase < num_m_blocks; mbase += m_step, ++it) {
        for (int j = 0; j < NJ; ++j, ++n_acc1) {
          mbar_wait(acc1_empty, (n_acc1 & 1) ^ 1);  // the epilogue warps of both CTAs have read the previous chunk
          tcgen05_fence_after();
          for (int kb = 0; kb < KB; ++kb) {
            if (j == 0) mbar_wait(&a_full[kb], it & 1);
            mbar_wait(&full1[s], ph);
            tcgen05_fence_after();
            const uint64_t a_desc = umma_desc_sw128(smem_u32(u_res + kb * kTile));
            const uint64_t b_desc = umma_desc_sw128(smem_u32(ring1 + s * kUnit1));
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_bf16_ss_pair(tmem_acc1, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
              umma_commit_pair(&empty1[s]);
              if (j == NJ - 1) umma_commit_pair(&a_empty[kb]);
              if (kb == KB - 1) umma_commit_pair(acc1_full);
            }
            __syncwarp();
            if (++s == stages1) s = 0, ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===== producer 2: attention tiles into G and W2 half-units, in the order stream 2 consumes them =====
    int s = 0;
    uint32_t ph = 0, it = 0, g_use0 = 0, g_use1 = 0;  // writes into G slot 0 / 1 so far (TMA + epilogue)
    auto unit = [&](int row, int col) {
      mbar_wait(&empty2[s], ph ^ 1);
      if (elect_one()) {
        if (p.debug & 8) {
          if (leader) mbar_arrive(&full2[s]);
        } else {
          if (leader) mbar_arrive_expect_tx(&full2[s], 2 * kUnit2);
          tma_load_2d_pair(&tmap_w2, mapa_u32(smem_u32(&full2[s]), 0), ring2 + s * kUnit2, col, row + cta_rank * (NU / 2));
        }
      }
      __syncwarp();
      if (++s == stages2) s = 0, ph ^= 1;
    };
    for (int mbase = m_first; mbase < num_m_blocks; mbase += m_step, ++it) {
      const int m0 = (mbase + cta_rank) * kBlockM;
      // the final epilogue of the previous m-block stages its output boxes in G: wait until it is done with them
      mbar_wait(stage_free, (it & 1) ^ 1);
      for (int kk = 0; kk < KB; ++kk) {  // attention tile k-block kk -> G slot kk & 1, then its NI weight units
        const int slot = kk & 1;
        mbar_wait(&g_empty[slot], ((slot ? g_use1 : g_use0) & 1) ^ 1);
        if (slot) ++g_use1;
        else ++g_use0;
        if (elect_one()) {
          if (leader) mbar_arrive_expect_tx(&ga_full[slot], 2 * kTile);
          tma_load_2d_pair(&tmap_attn, mapa_u32(smem_u32(&ga_full[slot]), 0), g_buf + slot * kTile, kk * 64, m0);
        }
        __syncwarp();
        for (int i = 0; i < NI; ++i) unit(i * NU, kk * 64);
      }
      for (int j = 0; j < NJ; ++j) {
        for (int kk = 0; kk < 2; ++kk)
          for (int i = 0; i < NI; ++i) unit(i * NU, H + j * 128 + kk * 64);
        ++g_use0, ++g_use1;  // the epilogue warps' writes of chunk j
      }
    }
  } else if (warp == 3) {
    // ===== issuer 2 (leader): OUT = attn W2a^T, then OUT += gelu chunk j W2[:, H + 128 j ..]^T =====
    if (leader) {
      const uint32_t idesc = umma_idesc_bf16(256, NU);
      int s = 0;
      uint32_t ph = 0, it = 0, ga_use0 = 0, ga_use1 = 0, gg_use0 = 0, gg_use1 = 0;
      auto mma_unit = [&](uint32_t d_tmem, uint32_t a_addr, bool first_zero) {
        mbar_wait(&full2[s], ph);
        tcgen05_fence_after();
        const uint64_t a_desc = umma_desc_sw128(a_addr);
        const uint64_t b_desc = umma_desc_sw128(smem_u32(ring2 + s * kUnit2));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (first_zero && k == 0) ? 0u : 1u);
          umma_commit_pair(&empty2[s]);
        }
        __syncwarp();
        if (++s == stages2) s = 0, ph ^= 1;
      };
      for (int mbase = m_first; mbase < num_m_blocks; mbase += m_step, ++it) {
        mbar_wait(out_free, (it & 1) ^ 1);  // the previous m-block's OUT has been drained (both CTAs)
        tcgen05_fence_after();
        for (int kk = 0; kk < KB; ++kk) {
          const int slot = kk & 1;
          mbar_wait(&ga_full[slot], (slot ? ga_use1 : ga_use0) & 1);
          if (slot) ++ga_use1;
          else ++ga_use0;
          tcgen05_fence_after();
          for (int i = 0; i < NI; ++i) mma_unit(tmem_out + i * NU, smem_u32(g_buf + slot * kTile), kk == 0);
          if (elect_one()) {
            umma_commit_pair(&g_empty[slot]);
            if (kk == KB - 1) umma_commit_pair(attn_done);
          }
          __syncwarp();
        }
        for (int j = 0; j < NJ; ++j) {
          for (int kk = 0; kk < 2; ++kk) {
            mbar_wait(&gg_full[kk], (kk ? gg_use1 : gg_use0) & 1);
            if (kk) ++gg_use1;
            else ++gg_use0;
            tcgen05_fence_after();
            for (int i = 0; i < NI; ++i) mma_unit(tmem_out + i * NU, smem_u32(g_buf + kk * kTile), false);
            if (elect_one()) umma_commit_pair(&g_empty[kk]);
            __syncwarp();
          }
        }
        if (elect_one()) umma_commit_pair(out_full);
        __syncwarp();
      }
    }
  } else {
    // ===== epilogue warps =====
    const int q = warp & 3;
    const int cq = (warp - 4) >> 2;
    const uint32_t lane_t = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t smf_s = smem_u32(smf);
    const uint32_t g_s = smem_u32(g_buf);
    const int r_in_tile = q * 32 + lane;
    const int kk = cq >> 1;  // which 64-column tile of a hidden chunk this warp's 32 columns fall into
    const uint32_t acc1_empty_l = mapa_u32(smem_u32(acc1_empty), 0);  // the leader's barriers
    const uint32_t out_free_l = mapa_u32(smem_u32(out_free), 0);
    const uint32_t gg_full_l = mapa_u32(smem_u32(&gg_full[kk]), 0);
    uint32_t it = 0, n_acc1 = 0, g_use_mine = 0;  // writes so far into the G tile this warp writes (TMA + epilogue)
    for (int mbase = m_first; mbase < num_m_blocks; mbase += m_step, ++it) {
      const int row0 = (mbase + cta_rank) * kBlockM + q * 32;
      const int row = row0 + lane;
      // ---- hidden chunks: ACC1 -> +bias -> GELU -> bf16 -> G (A operand layout: 128-byte swizzle, K-major)
      g_use_mine += kk ? KB / 2 : (KB + 1) / 2;  // the attention tiles (TMA) written into that G tile first
      for (int j = 0; j < NJ; ++j, ++n_acc1) {
        mbar_wait(acc1_full, n_acc1 & 1);
        tcgen05_fence_after();
        uint32_t v[32];
        tmem_ld32(lane_t + 384 + cq * 32, v);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc1_empty_l);
        uint32_t w[16];
        const uint32_t bias_s = smf_s + (j * 128 + cq * 32) * 4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (p.debug & 1) {
            w[2 * i] = v[4 * i], w[2 * i + 1] = v[4 * i + 2];
            continue;
          }
          const float4 bv = ld_shared_f4(bias_s + i * 16);
          const float y0 = gelu_fast(__uint_as_float(v[4 * i + 0]) + bv.x);
          const float y1 = gelu_fast(__uint_as_float(v[4 * i + 1]) + bv.y);
          const float y2 = gelu_fast(__uint_as_float(v[4 * i + 2]) + bv.z);
          const float y3 = gelu_fast(__uint_as_float(v[4 * i + 3]) + bv.w);
          w[2 * i] = pack_bf16x2(y0, y1);
          w[2 * i + 1] = pack_bf16x2(y2, y3);
        }
        // Stream 1 runs ahead of stream 2: chunk 0 may be ready while the attention tiles still cycle through G.  A parity wait
        // only tells adjacent phases apart, so first wait for the end of the attention phase as such ...
        if (j == 0) mbar_wait(attn_done, it & 1);
        mbar_wait(&g_empty[kk], (g_use_mine & 1) ^ 1);  // ... then: the MMAs that read the previous content of this tile are done
        ++g_use_mine;
        const uint32_t tile = g_s + kk * kTile + r_in_tile * 128;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (p.debug & 4) break;
          const int chunk = (cq & 1) * 4 + c;
          st_shared_v4(tile + ((chunk ^ (r_in_tile & 7)) << 4), w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
        }
        fence_proxy_async();  // generic-proxy writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(gg_full_l);
      }
      // ---- OUT: h += gate * (acc + b2) as TMA f32 reduce-add boxes of 16 columns, staged in G (free once out_full fired)
      mbar_wait(out_full, it & 1);
      tcgen05_fence_after();
      const int b = (row < p.rows ? row : p.rows - 1) / p.rows_per_sample;
      const float* gate = p.gate + (size_t)b * p.gate_stride;
      const int wcols = H / 4;  // this warp's output columns
      const uint32_t stage_s = g_s + (warp - 4) * 2048;
      for (int bx = 0; bx < wcols / 16; ++bx) {
        const int col = cq * wcols + bx * 16;
        uint32_t v[16];
        tmem_ld16(lane_t + col, v);
        tmem_ld_wait();
        if (bx == wcols / 16 - 1) {  // OUT is in registers: the next m-block's first MMAs may overwrite it
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(out_free_l);
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
        if (lane == 0 && !(p.debug & 2)) {
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
  cluster_sync_all();  // no CTA leaves (or frees TMEM) while the pair still works
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc_pair<512>(tmem_base);
  }
}

}  // namespace lam

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
// Shared memory: u tile resident (H/64 k-blocks of 16 KB) | G = two 128 x 64 bf16 tiles (A operand of the second GEMM: the GELU
// output of a hidden chunk, written by the epilogue warps in the 128-byte swizzled layout; also the staging area of the output
// boxes) | ring 1 (16 KB stages: W1m half-units [64 rows x 128 k], and the attention tile k-blocks [128 x 64]) | ring 2 (W2
// half-units [NU/2 rows x 64 k]) | constants (b1m, b2, the gate rows of this m-block) | barriers.
//
// Two independent MMA streams, each with its own producer warp, ring and issuing warp (measured, scripts/issue_bench.cu: the
// wait / elect / 4 x UTCHMMA / commit sequence costs ~350 cycles of the issuing warp per unit, more than the 256 cycles four
// N = 128 MMAs occupy the tensor pipe — one issuer for both streams left the pipe at 40 %):
//     stream 1 (warps 0,1):  for each hidden chunk j (128 columns):  ACC1 = u W1m_j^T     (units of 2 k-blocks = 8 MMAs, N = 128)
//     stream 2 (warps 2,3):  OUT = attn W2[:, :H]^T ;  then per chunk j:  OUT += gelu(ACC1 + b1m_j) W2[:, H+128j ..]^T
//                            (units of NU = 192 output columns for H = 384, else min(H, 256); 4 MMAs each)
// The attention tiles travel through ring 1: half of them right behind the W1m units of chunk 0 (prefetched while the previous
// m-block drains), the other half behind the units of the last chunk (prefetched during the last chunks), so neither group
// waits for a load and G stays free for the GELU chunks.  (All six at the start cost 7.9k cycles for 4.6k of MMA: with three ring
// stages the later tiles were loaded only after the earlier ones had been consumed.)
// Warps 4..19: epilogue (TMEM lane quarter = warp % 4, column quarter = (warp - 4) / 4): ACC1 -> + bias -> GELU -> bf16 -> G;
// at the end of the m-block OUT -> gate * (acc + b2) -> TMA f32 reduce-add into h (16-column x 32-row boxes staged in G; measured,
// scripts/drain_bench.cu: 4.8 us per m-block alone, 12 us when all CTAs drain at once (HBM read-modify-write), against 6.4 / 16 us
// for 8-column boxes and worse for red.global or ld+st from registers; sending 1..6 of a warp's 6 boxes through red.global.add.v4
// next to the TMA path made the kernel 1..11 % slower).
// Barrier protocol in the pair (same offsets in both CTAs): "full" barriers live on the leader (both producers' TMA loads
// complete_tx there, the leader arms 2x the bytes; both CTAs' epilogue warps arrive there), "empty" barriers are per CTA (the
// leader's issuing threads commit to both CTAs).  Every waiter follows its barrier phase by phase (a parity wait cannot tell
// phases two apart).
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
  // LN mode (ln_scale != nullptr): the drain also applies the NEXT block's pre-norm + modulate (latent_si_v31.py:50,57:
  // u = LayerNorm(h) * (1 + scale) + shift, no affine, eps 1e-6) to the rows it has just updated and writes them to u_out as bf16 —
  // the rows' new residual values are in registers anyway, which deletes one ln_modulate launch (a 270 MB pass over h and u) per block.
  // h is then updated with plain loads / stores of the rows (each thread owns its row x column quarter) instead of TMA reduce-adds.
  float* h;                  // [rows, H] residual stream
  __nv_bfloat16* u_out;      // [rows, H]; may be the buffer tmap_u reads (an m-block's u tile is in shared memory long before its drain)
  const float* ln_shift;     // shift / scale of sample b at + b * gate_stride, [H] each
  const float* ln_scale;
  long long* trace;  // profiling aid: when set, CTA 0 records (tag << 48 | clock) events, 4096 slots per role (see TRACE below)
  int debug;  // profiling aid: 1 skip the GELU math, 2 skip the output reduce-add, 4 skip the G tile writes, 8 skip the weight TMA loads
};

// output columns per stream-2 MMA (N of the instruction; each CTA of the pair stages N/2 weight rows)
static inline __host__ __device__ int fused_mlp_out_unit(int H) { return H % 192 == 0 && H > 256 ? 192 : (H <= 256 ? H : 128); }

struct FusedMlpSmem {
  int u_bytes, g_bytes, ring1_bytes, ring2_bytes, const_bytes, bar_bytes, stat_bytes, total;
};
static inline __host__ __device__ FusedMlpSmem fused_mlp_smem(int H, int M, int stages1, int stages2) {
  FusedMlpSmem s;
  s.u_bytes = (H / 64) * 16384;
  s.g_bytes = 2 * 16384;
  s.ring1_bytes = stages1 * 16384;
  s.ring2_bytes = (stages2 * fused_mlp_out_unit(H) * 64 + 1023) / 1024 * 1024;
  s.const_bytes = ((M + 3 * H) * 4 + 15) / 16 * 16;  // b1m | b2 | gate rows of two samples
  s.bar_bytes = 512;
  s.stat_bytes = 4096;  // LN mode: (mean, M2) of [4 column quarters][128 rows]
  s.total = s.u_bytes + s.g_bytes + s.ring1_bytes + s.ring2_bytes + s.const_bytes + s.bar_bytes + s.stat_bytes;
  return s;
}

// byte offset of (row r, 16-byte chunk c) inside a TMA-staged box with 32-byte rows (SWIZZLE_32B on the host side)
__device__ __forceinline__ uint32_t stage_off32(int r, int c) { return r * 32 + ((c ^ ((r >> 2) & 1)) << 4); }

__global__ void __launch_bounds__(kFusedThreads, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tmap_u, const __grid_constant__ CUtensorMap tmap_attn,
                 const __grid_constant__ CUtensorMap tmap_w1, const __grid_constant__ CUtensorMap tmap_w2,
                 const __grid_constant__ CUtensorMap tmap_h, const __grid_constant__ CUtensorMap tmap_u_st, int num_m_blocks, int stages1,
                 int stages2, FusedMlpParams p) {
  constexpr int kTile = 16384;   // one [128 x 64] bf16 tile (u k-block, G tile, attention tile) = one ring-1 stage
  constexpr int kUnit1 = 16384;  // this CTA's half of a W1m unit: [64 rows x 128 k] as two 128-byte-swizzled [64 x 64] boxes
  const int H = p.H, M = p.M;
  const int KB = H / 64;               // k-blocks of u / attn
  const int NU = fused_mlp_out_unit(H);
  const int NI = H / NU;               // stream-2 units per k-block
  const int NJ = M / 128;              // hidden chunks
  const int kUnit2 = NU * 64;          // this CTA's half of a W2 unit: [NU/2 rows x 64 k]
  const int KB1 = KB / 2;              // attention k-blocks consumed at the start of an m-block; the other KB - KB1 at its end

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const FusedMlpSmem plan = fused_mlp_smem(H, M, stages1, stages2);
  uint8_t* u_res = smem;
  uint8_t* g_buf = u_res + plan.u_bytes;
  uint8_t* ring1 = g_buf + plan.g_bytes;
  uint8_t* ring2 = ring1 + plan.ring1_bytes;
  float* smf = reinterpret_cast<float*>(ring2 + plan.ring2_bytes);  // [M] b1m | [H] b2 | [2][H] gate
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(smf) + plan.const_bytes);
  uint64_t* full1 = bars;            // [8] ring 1                      (leader)
  uint64_t* empty1 = bars + 8;       // [8]                             (per CTA)
  uint64_t* full2 = bars + 16;       // [8] ring 2                      (leader)
  uint64_t* empty2 = bars + 24;      // [8]                             (per CTA)
  uint64_t* a_full = bars + 32;      // [8] u k-blocks                  (leader)
  uint64_t* a_empty = bars + 40;     // [8]                             (per CTA)
  uint64_t* gg_full = bars + 48;     // [2] GELU output in G tile       (leader; 8 warps of each CTA)
  uint64_t* g_empty = bars + 50;     // [2] G tile consumed by the MMAs (per CTA)
  uint64_t* acc1_full = bars + 52;   //                                 (per CTA)
  uint64_t* acc1_empty = bars + 53;  //                                 (leader; 16 warps of each CTA)
  uint64_t* out_full = bars + 54;    //                                 (per CTA)
  uint64_t* out_free = bars + 55;    //                                 (leader; 16 warps of each CTA)
  uint64_t* attn_done = bars + 56;   // the attention tiles of this m-block have left ring 1 (per CTA; issuer 1 waits on the leader)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 57);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform: TMA / MMA operands stay in uniform registers
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
      mbar_init(&gg_full[s], 2 * (kWsEpiWarps / 2));  // the 8 warps of each CTA whose hidden columns fall into this 64-column tile
      mbar_init(&g_empty[s], 1);
    }
    mbar_init(acc1_full, 1);
    mbar_init(acc1_empty, 2 * kWsEpiWarps);
    mbar_init(out_full, 1);
    mbar_init(out_free, 2 * kWsEpiWarps);
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
  // event trace of CTA 0 (roles: 0 epilogue warp 4, 1 issuer 1, 2 issuer 2)
  int trace_n = 0;
  auto TRACE = [&](int role, int tag) {
    if (p.trace != nullptr && blockIdx.x == 0 && lane == 0 && trace_n < 4095) {
      p.trace[role * 4096 + 1 + trace_n] = (static_cast<long long>(tag) << 48) | (clock64() & 0xffffffffffffll);
      p.trace[role * 4096] = ++trace_n;
    }
  };

  // Ring 1 carries, per m-block and in this order: the W1m units of chunk 0 (KB/2 stages, consumed by issuer 1), the attention
  // tile k-blocks (KB stages, consumed by issuer 2), the W1m units of chunks 1.. (consumed by issuer 1).  Each consumer skips
  // the other's stages — but a parity wait cannot tell phases two apart, so neither may wait on a stage whose previous use (by
  // the other consumer) is still outstanding: issuer 2 first waits for chunk 0's ACC1 (its units have then left the ring),
  // issuer 1 waits for attn_done before it goes on to chunk 1.
  auto ring1_skip = [&](int& s, uint32_t& ph, int n) {
    s += n;
    while (s >= stages1) s -= stages1, ph ^= 1;
  };

  if (warp == 0) {
    // ===== producer 1: u k-blocks, W1m half-units and attention tiles, in ring-1 order =====
    int s = 0;
    uint32_t ph = 0, it = 0;
    for (int mbase = m_first; mbase < num_m_blocks; mbase += m_step, ++it) {
      const int m0 = (mbase + cta_rank) * kBlockM;  // may be past the end in the last sweep: the TMA zero-fills, nothing is stored
      for (int kb = 0; kb < KB; ++kb) {  // u: its buffers are released by the last chunk of the previous m-block
        mbar_wait(&a_empty[kb], (it & 1) ^ 1);
        if (elect_one()) {
          if (leader) mbar_arrive_expect_tx(&a_full[kb], 2 * kTile);
          tma_load_2d_pair(&tmap_u, mapa_u32(smem_u32(&a_full[kb]), 0), u_res + kb * kTile, kb * 64, m0);
        }
        __syncwarp();
      }
      for (int j = 0; j < NJ; ++j) {
        for (int kb = 0; kb < KB; kb += 2) {
          mbar_wait(&empty1[s], ph ^ 1);
          if (elect_one()) {
            if (p.debug & 8) {
              if (leader) mbar_arrive(&full1[s]);
            } else {
              if (leader) mbar_arrive_expect_tx(&full1[s], 2 * kUnit1);
              const uint32_t bar = mapa_u32(smem_u32(&full1[s]), 0);
              const int w_row = 3 * H + j * 128 + cta_rank * 64;
              tma_load_2d_pair(&tmap_w1, bar, ring1 + s * kUnit1, kb * 64, w_row);
              tma_load_2d_pair(&tmap_w1, bar, ring1 + s * kUnit1 + kUnit1 / 2, kb * 64 + 64, w_row);
            }
          }
          __syncwarp();
          if (++s == stages1) s = 0, ph ^= 1;
        }
        if (j == 0 || j == NJ - 1) {  // attention k-blocks: the first KB1 after chunk 0's units, the rest after the last chunk's
          const int k_lo = j == 0 ? 0 : KB1, k_hi = j == NJ - 1 ? KB : KB1;
          for (int kk = k_lo; kk < k_hi; ++kk) {
            mbar_wait(&empty1[s], ph ^ 1);
            if (elect_one()) {
              if (leader) mbar_arrive_expect_tx(&full1[s], 2 * kTile);
              tma_load_2d_pair(&tmap_attn, mapa_u32(smem_u32(&full1[s]), 0), ring1 + s * kTile, kk * 64, m0);
            }
            __syncwarp();
            if (++s == stages1) s = 0, ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== issuer 1 (leader): ACC1 = u W1m_j^T for every hidden chunk =====
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, 128);
      const uint64_t a_desc0 = umma_desc_sw128(smem_u32(u_res));  // the 14-bit address field (bytes >> 4) cannot carry: smem < 256 KB
      const uint64_t b_desc0 = umma_desc_sw128(smem_u32(ring1));
      int s = 0;
      uint32_t ph = 0, it = 0, n_acc1 = 0;
      for (int mbase = m_first; mbase < num_m_blocks; mbase += m_step, ++it) {
        // attn_done completes twice per m-block (first / second group of attention tiles): wait for each in turn
        if (it > 0) mbar_wait(attn_done, 1);  // the previous m-block's trailing attention tiles have left ring 1
        for (int j = 0; j < NJ; ++j, ++n_acc1) {
          mbar_wait(acc1_empty, (n_acc1 & 1) ^ 1);  // the epilogue warps of both CTAs have read the previous chunk
          tcgen05_fence_after();
          TRACE(1, 1);
          for (int kb = 0; kb < KB; kb += 2) {  // one unit = two k-blocks = 8 MMAs (N = 128: 512 tensor-pipe cycles per wait/commit)
            if (j == 0) {
              mbar_wait(&a_full[kb], it & 1);
              mbar_wait(&a_full[kb + 1], it & 1);
            }
            mbar_wait(&full1[s], ph);
            tcgen05_fence_after();
            const uint64_t a_desc = a_desc0 + static_cast<uint64_t>(kb * (kTile >> 4));
            const uint64_t b_desc = b_desc0 + static_cast<uint64_t>(s * (kUnit1 >> 4));
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 8; ++k)
                umma_bf16_ss_pair(tmem_acc1, a_desc + (k >> 2) * (kTile >> 4) + 2 * (k & 3), b_desc + (k >> 2) * (kUnit1 >> 5) + 2 * (k & 3),
                                  idesc, (kb | k) != 0);
              umma_commit_pair(&empty1[s]);
              if (j == NJ - 1) {
                umma_commit_pair(&a_empty[kb]);
                umma_commit_pair(&a_empty[kb + 1]);
              }
              if (kb == KB - 2) umma_commit_pair(acc1_full);
            }
            __syncwarp();
            if (++s == stages1) s = 0, ph ^= 1;
          }
          TRACE(1, 2);
          if (j == 0) {  // the leading attention tiles (issuer 2)
            mbar_wait(attn_done, 0);
            TRACE(1, 3);
            ring1_skip(s, ph, KB1);
          }
          if (j == NJ - 1) ring1_skip(s, ph, KB - KB1);  // the trailing ones (waited for at the top of the next m-block)
        }
      }
    }
  } else if (warp == 2) {
    // ===== producer 2: W2 half-units, in the order stream 2 consumes them =====
    int s = 0;
    uint32_t ph = 0;
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
    for (int mbase = m_first; mbase < num_m_blocks; mbase += m_step) {
      for (int kk = 0; kk < KB1; ++kk)
        for (int i = 0; i < NI; ++i) unit(i * NU, kk * 64);
      for (int j = 0; j < NJ; ++j)
        for (int kk = 0; kk < 2; ++kk)
          for (int i = 0; i < NI; ++i) unit(i * NU, H + j * 128 + kk * 64);
      for (int kk = KB1; kk < KB; ++kk)
        for (int i = 0; i < NI; ++i) unit(i * NU, kk * 64);
    }
  } else if (warp == 3) {
    // ===== issuer 2 (leader): OUT = attn W2a^T, then OUT += gelu chunk j W2[:, H + 128 j ..]^T =====
    if (leader) {
      const uint32_t idesc = umma_idesc_bf16(256, NU);
      const uint64_t g_desc0 = umma_desc_sw128(smem_u32(g_buf));
      const uint64_t t_desc0 = umma_desc_sw128(smem_u32(ring1));
      const uint64_t b_desc0 = umma_desc_sw128(smem_u32(ring2));
      int s = 0, s1 = 0;
      uint32_t ph = 0, ph1 = 0, it = 0, n_chunk = 0;
      auto mma_unit = [&](uint32_t d_tmem, uint64_t a_desc, bool first_zero) {
        mbar_wait(&full2[s], ph);
        tcgen05_fence_after();
        const uint64_t b_desc = b_desc0 + static_cast<uint64_t>(s * (kUnit2 >> 4));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (first_zero && k == 0) ? 0u : 1u);
          umma_commit_pair(&empty2[s]);
        }
        __syncwarp();
        if (++s == stages2) s = 0, ph ^= 1;
      };
      for (int mbase = m_first; mbase < num_m_blocks; mbase += m_step, ++it) {
        TRACE(2, 1);
        mbar_wait(out_free, (it & 1) ^ 1);  // the previous m-block's OUT has been drained (both CTAs)
        tcgen05_fence_after();
        TRACE(2, 2);
        mbar_wait(acc1_full, n_chunk & 1);  // chunk 0 of this m-block is complete: its W1m units (issuer 1) have left ring 1
        ring1_skip(s1, ph1, KB / 2);
        auto attn_tiles = [&](int k_lo, int k_hi) {
          for (int kk = k_lo; kk < k_hi; ++kk) {
            mbar_wait(&full1[s1], ph1);
            tcgen05_fence_after();
            const uint64_t a_desc = t_desc0 + static_cast<uint64_t>(s1 * (kTile >> 4));
            for (int i = 0; i < NI; ++i) mma_unit(tmem_out + i * NU, a_desc, kk == 0);
            if (elect_one()) {
              umma_commit_pair(&empty1[s1]);
              if (kk == k_hi - 1) umma_commit_pair(attn_done);
            }
            __syncwarp();
            if (++s1 == stages1) s1 = 0, ph1 ^= 1;
          }
        };
        attn_tiles(0, KB1);  // prefetched into ring 1 while the previous m-block drained
        TRACE(2, 3);
        ring1_skip(s1, ph1, (NJ - 1) * (KB / 2));  // the W1m units of chunks 1.. (issuer 1)
        for (int j = 0; j < NJ; ++j, ++n_chunk) {
          for (int kk = 0; kk < 2; ++kk) {
            mbar_wait(&gg_full[kk], n_chunk & 1);
            tcgen05_fence_after();
            TRACE(2, 4 + kk);
            const uint64_t a_desc = g_desc0 + static_cast<uint64_t>(kk * (kTile >> 4));
            for (int i = 0; i < NI; ++i) mma_unit(tmem_out + i * NU, a_desc, false);
            if (elect_one()) umma_commit_pair(&g_empty[kk]);
            __syncwarp();
          }
        }
        // the trailing attention tiles were loaded behind the last chunk's W1m units, i.e. during the last chunks' compute (the last
        // G2 above implies chunk NJ - 1 is complete, so those ring-1 stages have been through issuer 1)
        attn_tiles(KB1, KB);
        TRACE(2, 6);
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
    const int etid = threadIdx.x - 128;
    const int b_max = (p.rows - 1) / p.rows_per_sample;
    uint32_t it = 0, n_chunk = 0, n_hload = 0;
    for (int mbase = m_first; mbase < num_m_blocks; mbase += m_step, ++it) {
      const int m0 = (mbase + cta_rank) * kBlockM;
      const int row0 = m0 + q * 32;
      const int row = row0 + lane;
      // gate rows of the (at most two, unless rows_per_sample < 128) samples this m-block touches -> shared memory.  The named
      // barrier also orders the previous m-block's staging boxes (each warp has waited for its own TMA reads) before G is rewritten.
      const int b0 = (m0 < p.rows ? m0 : p.rows - 1) / p.rows_per_sample;
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kWsEpiWarps) : "memory");
      if (p.ln_scale && elect_one()) {  // LN mode: this warp's h boxes of the m-block -> L2, a whole m-block ahead of the drain that reads them
        for (int bx = 0; bx < H / 64; ++bx) tma_prefetch_2d(&tmap_h, cq * (H / 4) + bx * 16, row0);
      }
      __syncwarp();
      for (int i = etid; i < 2 * H; i += 32 * kWsEpiWarps) {
        const int bb = b0 + (i >= H ? 1 : 0);
        smf[M + H + i] = p.gate[(size_t)(bb < b_max ? bb : b_max) * p.gate_stride + (i >= H ? i - H : i)];
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kWsEpiWarps) : "memory");
      // ---- hidden chunks: ACC1 -> +bias -> GELU -> bf16 -> G (A operand layout: 128-byte swizzle, K-major)
      for (int j = 0; j < NJ; ++j, ++n_chunk) {
        if (warp == 4) TRACE(0, 1);
        mbar_wait(acc1_full, n_chunk & 1);
        tcgen05_fence_after();
        if (warp == 4) TRACE(0, 2);
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
        if (warp == 4) TRACE(0, 3);
        mbar_wait(&g_empty[kk], (n_chunk & 1) ^ 1);  // the MMAs that read the previous chunk out of this tile are done
        if (warp == 4) TRACE(0, 4);
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
      if (p.ln_scale) {
        // ---- OUT, LN mode: thread = (row, column quarter).  Pass 1, per 16-column box: the h box (prefetched into L2 at the top of the
        // m-block) arrives by TMA in the warp's staging box, new h = h + gate * (acc + b2) leaves through the same box as a plain TMA
        // store and goes back into the OUT columns of TMEM; shifted one-pass statistics of the thread's columns.  The four quarters
        // of a row merge their (mean, M2) through shared memory.  Pass 2: u = (h - mean) rstd (1 + scale) + shift -> bf16 -> 1 KB
        // boxes -> TMA store.  Measured (4AA, B200): 348 us per launch against 280 us for the reduce-add drain + 67 us for the
        // ln_modulate launch it replaces — time-neutral (the drain sits on the epilogue warps' critical path and only 32 KB of staging
        // are left, i.e. 6 serial load -> add -> store rounds per warp), 270 MB less HBM traffic per block.  Variants: rows read and
        // written with 64-byte-per-thread global accesses 435 us (LSU-bound); TMA load + coalesced st.global through the box 358 us;
        // without the L2 prefetch and with per-thread u stores 361 us.
        const int wcols = H / 4, nb = wcols / 16, col0 = cq * wcols;
        const bool row_ok = row < p.rows;
        const uint32_t stage_s = g_s + (warp - 4) * 2048;
        const int e = warp - 4;  // this warp's TMA-load barrier: one of the never-used slots of the ring barrier arrays
        uint64_t* hbar = e < 4 ? &full1[4 + e] : e < 8 ? &empty1[e] : e < 10 ? &full2[e - 2] : e < 12 ? &empty2[e - 4] : e < 14 ? &a_full[e - 6] : &a_empty[e - 8];
        const bool issuer = elect_one();
        if (warp == 4) TRACE(0, 5);
        mbar_wait(out_full, it & 1);  // also: every MMA that reads G has completed, the staging boxes are free
        tcgen05_fence_after();
        if (warp == 4) TRACE(0, 6);
        if (issuer) {
          mbar_arrive_expect_tx(hbar, 2048);
          tma_load_2d(&tmap_h, hbar, g_buf + (warp - 4) * 2048, col0, row0);
        }
        const int b = (row_ok ? row : p.rows - 1) / p.rows_per_sample;
        const bool gate_in_smem = b - b0 <= 1;
        const uint32_t gate_s = smf_s + (M + H + (b - b0) * H) * 4;
        const float* gate_g = p.gate + (size_t)b * p.gate_stride;
        uint32_t v[16];
        tmem_ld16(lane_t + col0, v);
        float kshift = 0.f, s1 = 0.f, s2 = 0.f;
        for (int bx = 0; bx < nb; ++bx, ++n_hload) {
          const int col = col0 + bx * 16;
          tmem_ld_wait();
          mbar_wait(hbar, n_hload & 1);
          uint32_t nv[16];
#pragma unroll
          for (int jx = 0; jx < 4; ++jx) {
            const float4 gv = gate_in_smem ? ld_shared_f4(gate_s + col * 4 + jx * 16) : __ldg(reinterpret_cast<const float4*>(gate_g + col) + jx);
            const float4 bv = ld_shared_f4(smf_s + (M + col) * 4 + jx * 16);
            const float4 hv = ld_shared_f4(stage_s + stage_off64(lane, jx));
            float4 o;
            o.x = fmaf(gv.x, __uint_as_float(v[4 * jx + 0]) + bv.x, hv.x);
            o.y = fmaf(gv.y, __uint_as_float(v[4 * jx + 1]) + bv.y, hv.y);
            o.z = fmaf(gv.z, __uint_as_float(v[4 * jx + 2]) + bv.z, hv.z);
            o.w = fmaf(gv.w, __uint_as_float(v[4 * jx + 3]) + bv.w, hv.w);
            if (bx == 0 && jx == 0) kshift = o.x;
            const float d0 = o.x - kshift, d1 = o.y - kshift, d2 = o.z - kshift, d3 = o.w - kshift;
            s1 += (d0 + d1) + (d2 + d3);
            s2 = fmaf(d0, d0, s2), s2 = fmaf(d1, d1, s2), s2 = fmaf(d2, d2, s2), s2 = fmaf(d3, d3, s2);
            st_shared_v4(stage_s + stage_off64(lane, jx), __float_as_uint(o.x), __float_as_uint(o.y), __float_as_uint(o.z), __float_as_uint(o.w));
            nv[4 * jx + 0] = __float_as_uint(o.x), nv[4 * jx + 1] = __float_as_uint(o.y);
            nv[4 * jx + 2] = __float_as_uint(o.z), nv[4 * jx + 3] = __float_as_uint(o.w);
          }
          fence_proxy_async();
          __syncwarp();
          if (issuer && !(p.debug & 2)) {
            tma_store_2d_s(&tmap_h, stage_s, col, row0);
            bulk_commit();
          }
          tmem_st16(lane_t + col, nv);
          if (bx + 1 < nb) tmem_ld16(lane_t + col + 16, v);
          if (issuer) {
            bulk_wait_read<0>();  // the store has read the box: the next load (or the u boxes of pass 2) may overwrite it
            if (bx + 1 < nb) {
              mbar_arrive_expect_tx(hbar, 2048);
              tma_load_2d(&tmap_h, hbar, g_buf + (warp - 4) * 2048, col + 16, row0);
            }
          }
          __syncwarp();
        }
        tmem_st_wait();
        float2* stats = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(bars) + plan.bar_bytes);
        {
          const float inv_w = 1.0f / (float)wcols;
          const float mean_i = kshift + s1 * inv_w;
          const float m2_i = fmaxf(s2 - s1 * s1 * inv_w, 0.f);
          stats[cq * 128 + r_in_tile] = make_float2(mean_i, m2_i);
        }
        asm volatile("bar.sync 2, %0;" ::"n"(32 * kWsEpiWarps) : "memory");
        float mean, rstd;
        {
          const float2 a0 = stats[r_in_tile], a1 = stats[128 + r_in_tile], a2 = stats[256 + r_in_tile], a3 = stats[384 + r_in_tile];
          mean = 0.25f * ((a0.x + a1.x) + (a2.x + a3.x));
          const float e0 = a0.x - mean, e1 = a1.x - mean, e2 = a2.x - mean, e3 = a3.x - mean;
          const float m2 = ((a0.y + a1.y) + (a2.y + a3.y)) + (float)wcols * ((e0 * e0 + e1 * e1) + (e2 * e2 + e3 * e3));
          rstd = rsqrtf(m2 * (1.0f / (float)H) + 1e-6f);
        }
        const float4* sh_g = reinterpret_cast<const float4*>(p.ln_shift + (size_t)b * p.gate_stride + col0);
        const float4* sc_g = reinterpret_cast<const float4*>(p.ln_scale + (size_t)b * p.gate_stride + col0);
        // u boxes: [32 rows x 16 columns] bf16 = 1 KB, two of them in the warp's staging box (32-byte rows, SWIZZLE_32B)
        tmem_ld16(lane_t + col0, v);
        for (int bx = 0; bx < nb; ++bx) {
          tmem_ld_wait();
          uint32_t pk[8];
#pragma unroll
          for (int jx = 0; jx < 4; ++jx) {
            const float4 sc = __ldg(sc_g + bx * 4 + jx), sh = __ldg(sh_g + bx * 4 + jx);
            const float y0 = fmaf((__uint_as_float(v[4 * jx + 0]) - mean) * rstd, 1.0f + sc.x, sh.x);
            const float y1 = fmaf((__uint_as_float(v[4 * jx + 1]) - mean) * rstd, 1.0f + sc.y, sh.y);
            const float y2 = fmaf((__uint_as_float(v[4 * jx + 2]) - mean) * rstd, 1.0f + sc.z, sh.z);
            const float y3 = fmaf((__uint_as_float(v[4 * jx + 3]) - mean) * rstd, 1.0f + sc.w, sh.w);
            pk[2 * jx] = pack_bf16x2(y0, y1), pk[2 * jx + 1] = pack_bf16x2(y2, y3);
          }
          if (bx + 1 < nb) {
            tmem_ld16(lane_t + col0 + (bx + 1) * 16, v);
          } else {  // OUT is no longer needed: the next m-block's first MMAs may overwrite it
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(out_free_l);
          }
          if (issuer) bulk_wait_read<1>();  // the store two boxes back has read this half of the staging box
          __syncwarp();
          const uint32_t ub = stage_s + (bx & 1) * 1024;
          st_shared_v4(ub + stage_off32(lane, 0), pk[0], pk[1], pk[2], pk[3]);
          st_shared_v4(ub + stage_off32(lane, 1), pk[4], pk[5], pk[6], pk[7]);
          fence_proxy_async();
          __syncwarp();
          if (issuer && !(p.debug & 2)) {
            tma_store_2d_s(&tmap_u_st, ub, col0 + bx * 16, row0);
            bulk_commit();
          }
        }
        if (issuer) bulk_wait_read<0>();  // before the named barrier at the top of the next m-block lets G be rewritten
        __syncwarp();
      } else {
        // ---- OUT: h += gate * (acc + b2) as TMA f32 reduce-add boxes of 16 columns x 32 rows, one staging box per warp in G
        // (free once out_full fired: every MMA that reads G has completed)
        if (warp == 4) TRACE(0, 5);
        mbar_wait(out_full, it & 1);
        tcgen05_fence_after();
        if (warp == 4) TRACE(0, 6);
        const int b = (row < p.rows ? row : p.rows - 1) / p.rows_per_sample;
        const bool gate_in_smem = b - b0 <= 1;
        const uint32_t gate_s = smf_s + (M + H + (b - b0) * H) * 4;
        const float* gate_g = p.gate + (size_t)b * p.gate_stride;
        const int wcols = H / 4;  // this warp's output columns
        const uint32_t stage_s = g_s + (warp - 4) * 2048;
        const int col0 = cq * wcols;
        const bool issuer = elect_one();  // TMA issue, commit and wait all by the same (elected) lane
        uint32_t v[16];
        tmem_ld16(lane_t + col0, v);
        for (int bx = 0; bx < wcols / 16; ++bx) {
          const int col = col0 + bx * 16;
          tmem_ld_wait();
          float4 o[4];
#pragma unroll
          for (int jx = 0; jx < 4; ++jx) {
            const float4 gv = gate_in_smem ? ld_shared_f4(gate_s + col * 4 + jx * 16) : __ldg(reinterpret_cast<const float4*>(gate_g + col) + jx);
            const float4 bv = ld_shared_f4(smf_s + (M + col) * 4 + jx * 16);
            o[jx].x = gv.x * (__uint_as_float(v[4 * jx + 0]) + bv.x);
            o[jx].y = gv.y * (__uint_as_float(v[4 * jx + 1]) + bv.y);
            o[jx].z = gv.z * (__uint_as_float(v[4 * jx + 2]) + bv.z);
            o[jx].w = gv.w * (__uint_as_float(v[4 * jx + 3]) + bv.w);
          }
          if (bx + 1 < wcols / 16) {  // next 16 columns: TMEM load in flight while these are staged and stored
            tmem_ld16(lane_t + col + 16, v);
          } else {  // OUT is in registers: the next m-block's first MMAs may overwrite it
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(out_free_l);
          }
          if (issuer) bulk_wait_read<0>();  // the previous reduce-add has read the staging box
          __syncwarp();
#pragma unroll
          for (int ch = 0; ch < 4; ++ch)
            st_shared_v4(stage_s + stage_off64(lane, ch), __float_as_uint(o[ch].x), __float_as_uint(o[ch].y), __float_as_uint(o[ch].z),
                         __float_as_uint(o[ch].w));
          fence_proxy_async();
          __syncwarp();
          if (issuer && !(p.debug & 2)) {
            tma_reduce_add_2d_s(&tmap_h, stage_s, col, row0);
            bulk_commit();
          }
        }
        if (issuer) bulk_wait_read<0>();  // before the named barrier at the top of the next m-block lets G be rewritten
        __syncwarp();
      }
      if (warp == 4) TRACE(0, 7);
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

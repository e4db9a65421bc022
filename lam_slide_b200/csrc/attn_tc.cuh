// tcgen05 / TMEM attention for long sequences (temporal axis of the second stage; mmdit.py:42-55: softmax(q k^T / sqrt(hd)) v,
// no mask).  q and k arrive RMS-normalised + rotated and q pre-multiplied by hd^-0.5 * log2(e) from the linear1 epilogue, so the
// logits are bounded (see attn.cuh: attn_seq_kernel) and softmax is evaluated as exp2(s) / sum exp2(s) without a running maximum.
//
// One CTA per (sequence, head); TWO CTAs per SM (each ~112 KB of shared memory and 256 TMEM columns at S = 1000), so one CTA's
// K / V load phase and MMA round trips hide behind the other's exponentials.  All of K and V of the (sequence, head) live in
// shared memory in the canonical NO-SWIZZLE UMMA layout (8 x 16-byte "core matrices"): element (key, d) at
//     (key / 8) * (hd / 8) * 128 + (d / 8) * 128 + (key % 8) * 16 + (d % 8) * 2          bytes,
// which serves BOTH MMAs from one image each: K as the K-major B operand of S = Q K^T (N = keys, K = d), V as the MN-major
// B operand of O = P V (N = d, K = keys).  The contraction of S runs over d padded to a multiple of 16: the Q tile carries zero
// chunks there, so whatever (finite) bytes the K image has at those offsets do not matter.
//
// Warp roles (384 threads):
//   warp 0      : TMEM allocator + MMA issuer (warp-uniform loop, one elected lane issues)
//   warp 1      : Q loader (cp.async of a 128-row query tile into the same core-matrix layout, double buffered)
//   warps 2, 3  : idle after the K / V load
//   warps 4..11 : softmax: thread = (query row, half of the chunk's keys) — 8 warps per CTA, 4 per scheduler with two CTAs per SM,
//                 which is what it takes to keep the 16-lane MUFU unit busy.  Per 64-key chunk: tcgen05.ld S (fp32) -> exp2 ->
//                 partial row sum -> bf16 -> tcgen05.st P over the same TMEM columns (P aliases S) -> mbarrier.
// TMEM (256 columns): S0 | S1 (64 fp32 columns each, P over columns [0,16) and [32,48): DOUBLE BUFFERED so S of chunk n + 1 is computed while the
// softmax warps work on chunk n and the MMA round trip is off their critical path) | O (32 columns) | 2 columns for the row-sum exchange.
// Tensor work per chunk: S = Q K^T as 1-2 MMAs (M 128, N 64, K 16), O += P V as 4 MMAs (M 128, N 32, K 16, A from TMEM):
// ~250 tensor cycles against 512 MUFU cycles for the 8 k exponentials — the kernel is bound by the exponentials.
#pragma once
#include "attn.cuh"

namespace lam {

constexpr int kAtcThreads = 384;
constexpr int kAtcChunk = 64;   // keys per S tile

__host__ __device__ inline int atc_s128(int S) { return (S + 127) & ~127; }
template <int HD>
struct AtcCfg {
  static constexpr int KG = (HD / 8) * 128;      // bytes per 8-key group of the K / V images
  static constexpr int KSTEPS = (HD + 15) / 16;  // k16 steps of S = Q K^T
  static __host__ __device__ size_t kv_bytes(int S) { return (size_t)atc_s128(S) / 8 * KG; }
  static __host__ __device__ size_t q_bytes() { return 2 * 8192; }
  // K image | V image | 128-byte zero tail (the N = 32 P V MMA reads 4 d-chunks per key group) | 2 Q tiles | barriers
  static __host__ __device__ size_t smem_bytes(int S) { return 2 * kv_bytes(S) + 128 + q_bytes() + 128; }
};

// no-swizzle UMMA shared-memory descriptor: start address, leading-dimension byte offset, stride-dimension byte offset
__device__ __forceinline__ uint64_t umma_desc_nosw(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
               "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
               : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr));
  return v;
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// One 128-key chunk of one query row: S (fp32, TMEM) -> exp2 -> row-sum partials -> bf16 P written over the same columns.
// MASKED: keys >= key_lim get probability 0 (instantiated for the last chunk of a sequence only).
template <int POLY, bool MASKED>
__device__ __forceinline__ void atc_softmax_sub(const uint32_t* sv, uint32_t p_col, int key0, int key_lim, float& l0, float& l1, float& l2,
                                                float& l3) {
  uint32_t pk[8];
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    const float x0 = __uint_as_float(sv[i]), x1 = __uint_as_float(sv[i + 1]);
    float e0 = ((i / 2) % 8 < POLY) ? poly_exp2(x0) : fast_exp2(x0);
    float e1 = ((i / 2) % 8 < POLY) ? poly_exp2(x1) : fast_exp2(x1);
    if (MASKED) {
      if (key0 + i >= key_lim) e0 = 0.f;
      if (key0 + i + 1 >= key_lim) e1 = 0.f;
    }
    if ((i & 4) == 0) l0 += e0, l1 += e1;
    else l2 += e0, l3 += e1;
    pk[i >> 1] = pack_bf16x2(e0, e1);
  }
  tmem_st8(p_col, pk);  // P over S: these columns only cover sub-chunks that are already in registers
}
// POLY: of every 8 exponentials, POLY are evaluated on the FMA pipe (poly_exp2) instead of MUFU.EX2.
// variant bits (debug aid, lamslide_debug_attention mode 3 + 4 * variant): 1 = swap LBO / SBO of the Q / K descriptors,
// 2 = swap LBO / SBO of the V descriptor.  0 is the layout derived from the canonical UMMA layouts (verified on B200).
template <int HD, int POLY>
__global__ void __launch_bounds__(kAtcThreads, 2)
attn_tc_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int H, int ldo, SeqMap sm, int heads, int variant) {
  using Cfg = AtcCfg<HD>;
  constexpr int CH = HD / 8;
  constexpr int KG = Cfg::KG;
  constexpr int KSTEPS = Cfg::KSTEPS;
  extern __shared__ __align__(1024) uint8_t atc_smem[];
  const int S = sm.S;
  const int S128 = atc_s128(S);
  uint8_t* k_img = atc_smem;
  uint8_t* v_img = k_img + Cfg::kv_bytes(S);
  uint8_t* q_img = v_img + Cfg::kv_bytes(S) + 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(q_img + Cfg::q_bytes());
  uint64_t* s_full = bars;        // [2] MMA -> softmax (S buffer n & 1)
  uint64_t* p_full = bars + 2;    // [2] softmax -> MMA
  uint64_t* o_done = bars + 4;    // MMA -> softmax (last P V of a tile)
  uint64_t* q_full = bars + 5;    // [2] loader -> MMA
  uint64_t* q_empty = bars + 7;   // [2] MMA -> loader
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int z = blockIdx.x / heads;
  const int hh = blockIdx.x % heads;
  const long long base = sm.base(z);
  const size_t ldq = (size_t)3 * H;
  const __nv_bfloat16* qptr = qkv + hh * HD;
  const __nv_bfloat16* kptr = qkv + H + hh * HD;
  const __nv_bfloat16* vptr = qkv + 2 * H + hh * HD;
  const int ntiles = (S + 127) / 128;
  const int nchunks = (S + kAtcChunk - 1) / kAtcChunk;

  if (tid == 0) {
    mbar_init(o_done, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 8);
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<256>(tmem_slot);

  // ---- K, V images (whole sequence); keys >= S are zero rows
  for (int idx = tid; idx < S128 * CH; idx += kAtcThreads) {
    const int key = idx / CH, c = idx % CH;
    const uint32_t off = (key >> 3) * KG + c * 128 + (key & 7) * 16;
    if (key < S) {
      const size_t tok = (size_t)(base + (long long)key * sm.seq_stride) * ldq + c * 8;
      cp_async16(k_img + off, kptr + tok, true);
      cp_async16(v_img + off, vptr + tok, true);
    } else {
      *reinterpret_cast<uint4*>(k_img + off) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(v_img + off) = make_uint4(0, 0, 0, 0);
    }
  }
  if (tid < 8) *reinterpret_cast<uint4*>(v_img + Cfg::kv_bytes(S) + tid * 16) = make_uint4(0, 0, 0, 0);
  for (int idx = tid; idx < 2 * 128; idx += kAtcThreads) {  // zero d-chunks CH .. 2 KSTEPS of the two Q buffers once (never overwritten)
    const int b = idx >> 7, row = idx & 127;
    for (int c = CH; c < 2 * KSTEPS; ++c) *reinterpret_cast<uint4*>(q_img + b * 8192 + (row >> 3) * 512 + c * 128 + (row & 7) * 16) = make_uint4(0, 0, 0, 0);
  }
  cp_async_commit();
  cp_async_wait<0>();
  fence_proxy_async();  // generic-proxy / cp.async writes -> visible to the tensor core (async proxy)
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t q_lbo = (variant & 1) ? 512 : 128, q_sbo = (variant & 1) ? 128 : 512;
  const uint32_t k_lbo = (variant & 1) ? KG : 128, k_sbo = (variant & 1) ? 128 : KG;
  const uint32_t v_lbo = (variant & 2) ? 128 : KG, v_sbo = (variant & 2) ? KG : 128;

  if (warp == 0) {
    // ===== MMA issuer.  Chunk n = (tile t, chunk c) uses S buffer n & 1: S(n) = Q_t K_c^T;  O_t += P(n) V_c.  S(n + 2) is issued
    // right behind P V(n) (in-order tensor pipe: the P it overwrites has been consumed), so two S tiles are always ahead. =====
    constexpr uint32_t idesc_qk = umma_idesc_bf16(128, kAtcChunk);               // A, B K-major
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 32) | (1u << 16);         // B (= V) MN-major
    const uint32_t k_addr = smem_u32(k_img), v_addr = smem_u32(v_img), q_addr = smem_u32(q_img);
    const int total = ntiles * nchunks;
    auto issue_qk = [&](int n) {
      const int t = n / nchunks, c = n % nchunks;
      if (c == 0) {
        mbar_wait(&q_full[t & 1], (t >> 1) & 1);
        tcgen05_fence_after();
      }
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < KSTEPS; ++j) {
          const uint64_t a = umma_desc_nosw(q_addr + (t & 1) * 8192 + j * 256, q_lbo, q_sbo);
          const uint64_t b = umma_desc_nosw(k_addr + c * (kAtcChunk / 8) * KG + j * 256, k_lbo, k_sbo);
          umma_bf16_ss(tmem_base + (n & 1) * kAtcChunk, a, b, idesc_qk, j);
        }
        umma_commit(&s_full[n & 1]);
        if (c == nchunks - 1) umma_commit(&q_empty[t & 1]);  // the query tile is no longer read
      }
      __syncwarp();
    };
    issue_qk(0);
    if (total > 1) issue_qk(1);
    for (int n = 0; n < total; ++n) {
      const int c = n % nchunks;
      mbar_wait(&p_full[n & 1], (n >> 1) & 1);
      tcgen05_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int s = 0; s < kAtcChunk / 16; ++s) {
          const uint64_t b = umma_desc_nosw(v_addr + (c * (kAtcChunk / 8) + 2 * s) * KG, v_lbo, v_sbo);
          umma_bf16_ts(tmem_base + 2 * kAtcChunk, tmem_base + (n & 1) * kAtcChunk + (s >> 1) * 32 + (s & 1) * 8, b, idesc_pv, (c | s) != 0);
        }
        if (c == nchunks - 1) umma_commit(o_done);
      }
      __syncwarp();
      if (n + 2 < total) issue_qk(n + 2);
    }
  } else if (warp == 1) {
    // ===== Q loader: tile t -> buffer t & 1 =====
    for (int t = 0; t < ntiles; ++t) {
      mbar_wait(&q_empty[t & 1], ((t >> 1) & 1) ^ 1);
      for (int idx = lane; idx < 128 * CH; idx += 32) {
        const int row = idx / CH, c = idx % CH;
        const int qrow = t * 128 + row;
        const bool ok = qrow < S;
        uint8_t* dst = q_img + (t & 1) * 8192 + (row >> 3) * 512 + c * 128 + (row & 7) * 16;
        cp_async16(dst, qptr + (size_t)(base + (long long)(ok ? qrow : 0) * sm.seq_stride) * ldq + c * 8, ok);
      }
      cp_async_commit();
      cp_async_wait<0>();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&q_full[t & 1]);
    }
  } else if (warp >= 4) {
    // ===== softmax: thread = (query row, half of the chunk's keys) =====
    const int quarter = warp & 3;
    const int half = (warp - 4) >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_col = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    uint32_t n = 0;
    for (int t = 0; t < ntiles; ++t) {
      const int qrow = t * 128 + row;
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
      for (int c = 0; c < nchunks; ++c, ++n) {
        mbar_wait(&s_full[n & 1], (n >> 1) & 1);
        tcgen05_fence_after();
        // each half overwrites only ITS OWN S columns with P (columns [32 h, 32 h + 16)), so the two warps that share a TMEM lane
        // quarter never touch each other's data and need no synchronisation inside a chunk
        const uint32_t buf = lane_col + (n & 1) * kAtcChunk + 32 * half;
        uint32_t sa[16], sb[16];
        tmem_ld16(buf, sa);
        tmem_ld16(buf + 16, sb);
        if (c > 0) {  // hand chunk n - 1's P to the MMA warp now: its stores have long drained, so nothing stalls here
          tmem_st_wait();
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[(n - 1) & 1]);
        }
        tmem_ld_wait();
        const int key_lim = S - c * kAtcChunk;  // keys >= S (zero K rows -> exp2(0) = 1) must not count: last chunk only
        if (key_lim >= kAtcChunk) {
          atc_softmax_sub<POLY, false>(sa, buf, 32 * half, key_lim, l0, l1, l2, l3);
          atc_softmax_sub<POLY, false>(sb, buf + 8, 32 * half + 16, key_lim, l0, l1, l2, l3);
        } else {
          atc_softmax_sub<POLY, true>(sa, buf, 32 * half, key_lim, l0, l1, l2, l3);
          atc_softmax_sub<POLY, true>(sb, buf + 8, 32 * half + 16, key_lim, l0, l1, l2, l3);
        }
      }
      tmem_st_wait();  // last chunk of the tile
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[(n - 1) & 1]);
      // ---- O of this tile: exchange the partial row sums, normalise, store bf16 (half h writes d in [16 h, 16 h + 16))
      // (through two spare TMEM columns of this lane: the two halves of a row are the same lane of two different warps)
      const float l_mine = (l0 + l1) + (l2 + l3);
      const uint32_t l_col = lane_col + 2 * kAtcChunk + 32;
      tmem_st1(l_col + half, __float_as_uint(l_mine));
      tmem_st_wait();
      tcgen05_fence_before();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      tcgen05_fence_after();
      const float l_other = __uint_as_float(tmem_ld1(l_col + (half ^ 1)));
      mbar_wait(o_done, t & 1);
      tcgen05_fence_after();
      uint32_t ov[16];
      tmem_ld16(lane_col + 2 * kAtcChunk + 16 * half, ov);
      tmem_ld_wait();
      const float inv = 1.f / (l_mine + l_other);
      if (qrow < S) {
        __nv_bfloat16* op = out + (size_t)(base + (long long)qrow * sm.seq_stride) * ldo + hh * HD + 16 * half;
#pragma unroll
        for (int d = 0; d < 16; d += 8) {
          if (16 * half + d < HD) {
            uint4 o4;
            o4.x = pack_bf16x2(__uint_as_float(ov[d + 0]) * inv, __uint_as_float(ov[d + 1]) * inv);
            o4.y = pack_bf16x2(__uint_as_float(ov[d + 2]) * inv, __uint_as_float(ov[d + 3]) * inv);
            o4.z = pack_bf16x2(__uint_as_float(ov[d + 4]) * inv, __uint_as_float(ov[d + 5]) * inv);
            o4.w = pack_bf16x2(__uint_as_float(ov[d + 6]) * inv, __uint_as_float(ov[d + 7]) * inv);
            *reinterpret_cast<uint4*>(op + d) = o4;
          }
        }
      }
      tcgen05_fence_before();  // O is read: the next tile's first P V (ordered behind this thread's next P) may overwrite it
      asm volatile("bar.sync 1, 256;" ::: "memory");  // both halves have read the exchanged sums before they are rewritten
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

}  // namespace lam

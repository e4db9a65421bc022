// tcgen05 / TMEM attention for long sequences — the temporal axis of the second stage (mmdit.py:42-55: softmax(q k^T / sqrt(hd)) v,
// no mask; 4AA: 128 sequences x 16 heads, S = 1000, hd = 24).  q and k arrive RMS-normalised + rotated and q pre-multiplied by
// hd^-0.5 * log2(e) from the linear1 epilogue, so the logits are bounded (checked on the host from the QK-norm scales) and softmax is
// evaluated as exp2(s) / sum exp2(s) without a running maximum.
//
// What bounds this op is the exponential, not the tensor pipe: hd = 24 gives 96 tensor FLOPs per exp2, and MUFU.EX2 delivers 16 per
// clock and SM — 0.47 ms for the 2.05 G exponentials of one 4AA launch at 1.85 GHz.  Measured on B200 (scripts/attn_ubench.cu):
//   tcgen05.ld              ~900 B/clk/SM with >= 8 warps or >= 4 loads in flight (not a bound: S needs ~100 B/clk/SM)
//   MUFU.EX2 alone          15.9 exp/clk/SM
//   + 3 or 4 of every 8 pairs on an FMA-pipe polynomial (Cody-Waite split + degree-3 minimax, packed f32x2): 20 - 21 exp/clk/SM inside
//     a tcgen05.ld -> exp2 -> row sum -> bf16 -> tcgen05.st loop (17.6 / 19.9 with ONE warp per scheduler); more loses again.
// So the kernel is a "softmax engine" that keeps MUFU and the FMA pipe busy side by side, with everything else off its critical path:
//
// Persistent kernel, ONE CTA per SM, 20 warps.  A CTA walks (sequence, head) items  blockIdx.x, + gridDim.x, ...  K and V of the item
// live in shared memory for the whole item in the canonical NO-SWIZZLE UMMA layout (8 x 16-byte core matrices): element (key, d) at
//     (key / 8) * (hd / 8) * 128 + (d / 8) * 128 + (key % 8) * 16 + (d % 8) * 2          bytes,
// which serves BOTH MMAs from one image each (K as the K-major B operand of S = Q K^T, V as the MN-major B operand of O = P V), and are
// DOUBLE BUFFERED across items: the loader warp fills the next item's images while this one is being computed.
// Two 128-row query tiles are in flight (groups A and B: alternate tiles of the CTA's tile stream).  TMEM columns of a group: S (128
// keys = 128 fp32 columns), P (bf16 pairs: keys 0 .. 63 double buffered, keys 64 .. 127 single, 96 columns), O (32 columns).
//   warps 0, 1   : MMA issuer of group A / B:   logits of S_g(n) in registers -> S_g(n + 1) = Q_g K_c'^T (KSTEPS x M128 N128 K16);
//                                               P_g(n) complete -> O_g += P_g(n) V_c (8 x M128 N32 K16, A from TMEM) -> commits
//   warp 2       : K / V loader (cp.async into the core-matrix layout)     warp 3 : Q-tile loader
//   warps 4..11  : softmax of group A, thread = (query row, half of the chunk's keys); warps 12..19 : group B.  Per 128-key chunk a
//                  thread takes 64 logits: 4 x (tcgen05.ld 16 columns, double buffered) -> exp2 (MUFU / polynomial mix) -> packed
//                  row sum -> bf16x2 -> tcgen05.st 8 columns of P; the two halves of a row add their sums through shared memory.
// Every hand-shake through an mbarrier, tcgen05.wait or fence costs the issuing warp 50 - 100 cycles in series with its exponentials
// (~450 cycles per chunk step to see S and pull the first logits, ~100 to hand P over, ~300 amortised for the O epilogue), and a
// round trip through the MMA warp ~700 (a small-N tcgen05.mma costs its issuer ~48 cycles whatever the tensor pipe needs: 350 - 450
// cycles to issue a chunk's Q K^T or P V with the commits).  The protocol keeps the round trips off the softmax warps' path —
//   * the S buffer is released when a warp's last logits are in registers, half way through its exponentials, and Q K^T of the next
//     chunk executes under the rest;
//   * P has its own columns, the lower half double buffered: only the warps of the upper key half wait for P V of the previous chunk,
//     and only before their first store;
//   * one MMA warp per group (one for both: +30 %) —
// and FOUR softmax warps per scheduler hide each other's hand-shake latency.  With two (thread = whole row, one warp of each group
// per scheduler) the groups lock IN PHASE — when one stalls the other speeds up, so any offset decays — and the hand-shake cycles are
// idle pipes: 565 us per 4AA launch; half-step phase barriers between the groups: 578 us; a baton that lets one group compute at a
// time: 650 - 690 us (one warp per scheduler reaches 3.7 exp/clk, two 5); three-slot rings of 64-key chunks: 584 - 762 us (twice the
// hand-shakes per exponential); all 64 logits of a thread loaded at once to release S earlier: 563 us (register spills at the
// 96-register limit of 20 warps); Q K^T and P V issued by separate warps (Q K^T issuer also loading its group's query tiles, one P V
// issuer for both groups): 539 us — the issue time of the MMA warp is not on the critical loop.  This version: 524 us against 661 us
// for the mma.sync kernel (the O epilogue of a tile deferred into the next tile's first chunk step, where nobody has to wait for the last
// P V: 535 us — it delays that step's hand-off; the first logit blocks of step n + 1 requested at the end of step n whenever S(n + 1) has
// already landed: 586 us); the exponent engine alone would need 390 us, and the pipeline WITHOUT any exponential (variant 4: logits
// passed through) takes 360 us: per chunk step and group ~1500 cycles of tcgen05.ld / wait / st / mbarrier latency in series in every
// softmax warp, of which the exponentials hide about half.
// After the last chunk of a tile the group reads O (tcgen05.ld), scales by 1 / row sum and stores bf16; tensor pipe ~25 % busy.
// The contraction of S runs over d padded to a multiple of 16: the Q image carries zero chunks there, so whatever (finite) bytes
// the K image has at those offsets do not matter; O is computed with N = 32, the columns >= hd are ignored.
#pragma once
#include "attn.cuh"

namespace lam {

constexpr int kAtcThreads = 640;   // 4 control warps + 2 groups x 8 softmax warps (<= 96 registers per thread)
constexpr int kAtcChunk = 128;    // keys per S tile
constexpr int kAtcQBytes = 8192;  // one 128-row query tile image: 16 row groups x 4 d-chunks x 128 B
// TMEM columns of group g (256 g + ...): S 0 .. 127 (fp32 logits); P (bf16 pairs) in two halves — keys 0 .. 63 double buffered at
// 128 + 32 (n & 1), keys 64 .. 127 at 192 — so that a step can start writing P while P V of the previous step is still reading; O 224 .. 255
constexpr int kAtcGroupCols = 256, kAtcPLoCol = 128, kAtcPHiCol = 192, kAtcOCol = 224;

__host__ __device__ inline int atc_spad(int S) { return (S + kAtcChunk - 1) / kAtcChunk * kAtcChunk; }  // keys held by the K / V images
template <int HD>
struct AtcCfg {
  static constexpr int KG = (HD / 8) * 128;      // bytes per 8-key group of the K / V images
  static constexpr int KSTEPS = (HD + 15) / 16;  // k16 steps of S = Q K^T
  static __host__ __device__ size_t kv_bytes(int S) { return (size_t)atc_spad(S) / 8 * KG; }
  // one item buffer: K image | V image | 128-byte zero tail (the N = 32 P V MMA and the padded Q K^T step read one d-chunk past the
  // last key group)
  static __host__ __device__ size_t item_bytes(int S) { return 2 * kv_bytes(S) + 128; }
  // 2 item buffers | 4 Q tiles (2 groups x 2) | partial row sums [2 groups][2 key halves][128 rows] | barriers
  static __host__ __device__ size_t smem_bytes(int S) { return 2 * item_bytes(S) + 4 * kAtcQBytes + 2048 + 256; }
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
// ---- packed f32x2 arithmetic (FFMA2 / FADD2: one issue slot for two lanes' worth of fp32 work)
__device__ __forceinline__ uint64_t f32x2_pack(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void f32x2_unpack(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t f32x2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t f32x2_add(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// 2^x for two values on the FMA pipe: round-to-nearest split x = n + f (magic-number add), degree-3 minimax polynomial for 2^f on
// [-0.5, 0.5] (max rel err 7.5e-5, far below the bf16 rounding P gets), exponent patched in with an integer shift-add.  |x| < 120.
__device__ __forceinline__ void poly_exp2_x2(float x0, float x1, float& e0, float& e1) {
  const uint64_t X = f32x2_pack(x0, x1);
  const uint64_t t = f32x2_add(X, f32x2_pack(12582912.f, 12582912.f));  // 1.5 * 2^23: the low mantissa bits of t hold round(x)
  const uint64_t u = f32x2_add(t, f32x2_pack(-12582912.f, -12582912.f));
  const uint64_t f = f32x2_fma(u, f32x2_pack(-1.f, -1.f), X);
  uint64_t p = f32x2_fma(f, f32x2_pack(0.05517146f, 0.05517146f), f32x2_pack(0.24261086f, 0.24261086f));
  p = f32x2_fma(p, f, f32x2_pack(0.69326099f, 0.69326099f));
  p = f32x2_fma(p, f, f32x2_pack(0.99992809f, 0.99992809f));
  float p0, p1, t0, t1;
  f32x2_unpack(p, p0, p1);
  f32x2_unpack(t, t0, t1);
  e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

// 16 logits of one query row -> 8 packed bf16x2 probabilities + packed partial row sums.  POLY of the 8 pairs go through the
// FMA-pipe polynomial, the rest through MUFU.EX2.  MASKED: keys >= nvalid get probability 0 (last chunk of a sequence only).
template <int POLY, bool MASKED>
__device__ __forceinline__ void atc_exp16(const uint32_t* sv, uint32_t* pk, uint64_t& lsum, int nvalid) {
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    const float x0 = __uint_as_float(sv[i]), x1 = __uint_as_float(sv[i + 1]);
    float e0, e1;
    if (POLY < 0) {  // profiling aid: no exponentials at all (what the rest of the pipeline costs)
      e0 = x0, e1 = x1;
    } else if (((i >> 1) & 7) < POLY) {
      poly_exp2_x2(x0, x1, e0, e1);
    } else {
      e0 = fast_exp2(x0), e1 = fast_exp2(x1);
    }
    if (MASKED) {
      if (i >= nvalid) e0 = 0.f;
      if (i + 1 >= nvalid) e1 = 0.f;
    }
    lsum = f32x2_add(lsum, f32x2_pack(e0, e1));
    pk[i >> 1] = pack_bf16x2(e0, e1);
  }
}

// Walks the (item, query tile, key chunk) sequence of one softmax group — the same order in the MMA warp, the loaders and the softmax
// warps.  The tiles of ALL items of the CTA form one stream u = k * ntiles + t (k = index in the CTA's item list); group g takes the
// tiles with u = g (mod 2), so the two groups never drift more than one tile apart whatever the tile count (with a per-item split
// an odd tile count lets one group run ahead until it needs a K / V buffer the other still uses).
struct AtcCursor {
  int u;               // position in the CTA's tile stream
  int k, t, c;         // item index in the CTA's list (K / V buffer = k & 1), query tile, key chunk
  int job, n;          // running counts of tiles and chunk steps of this group (barrier parities)
  int item;            // global item index = first + k * stride
  bool valid;
  __device__ __forceinline__ void place(int first, int stride, int n_items, int ntiles) {
    k = u / ntiles, t = u - k * ntiles;
    item = first + k * stride;
    valid = item < n_items;
  }
  __device__ __forceinline__ void init(int g, int first, int stride, int n_items, int ntiles) {
    u = g, c = 0, job = 0, n = 0;
    place(first, stride, n_items, ntiles);
  }
  __device__ __forceinline__ void advance(int first, int stride, int n_items, int ntiles, int nchunks) {
    ++n;
    if (++c < nchunks) return;
    c = 0, ++job, u += 2;
    place(first, stride, n_items, ntiles);
  }
  // first / last tile this group works on in the item (K / V buffer hand-over)
  __device__ __forceinline__ bool first_in_item() const { return t < 2; }
  __device__ __forceinline__ bool last_in_item(int ntiles) const { return t + 2 >= ntiles; }
};

// POLY: of every 8 exponential pairs, POLY are evaluated on the FMA pipe instead of MUFU.EX2.
// TRACE (profiling aid): MMA warp 0 and softmax warp 4 of every CTA accumulate the cycles they spend in each kind of wait / work and
// write them to trace[blockIdx.x * 32 + ...] (see scripts/gpu_time_kernels.py: attn_tc_trace).
template <int HD, int POLY, bool TRACE = false>
__global__ void __launch_bounds__(kAtcThreads, 1)
attn_tc_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int H, int ldo, SeqMap sm, int heads, int n_items,
               long long* trace) {
  long long tr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tr_t = 0;
#define ATC_TIC() do { if (TRACE) tr_t = clock64(); } while (0)
#define ATC_TOC(i) do { if (TRACE) { const long long _n = clock64(); tr[i] += _n - tr_t; tr_t = _n; } } while (0)
  using Cfg = AtcCfg<HD>;
  constexpr int CH = HD / 8;
  constexpr int KG = Cfg::KG;
  constexpr int KSTEPS = Cfg::KSTEPS;
  extern __shared__ __align__(1024) uint8_t atc_smem[];
  const int S = sm.S;
  const int Spad = atc_spad(S);
  const int ntiles = (S + 127) / 128, nchunks = Spad / kAtcChunk;
  const uint32_t kvb = (uint32_t)Cfg::kv_bytes(S), itemb = (uint32_t)Cfg::item_bytes(S);
  uint8_t* q_img = atc_smem + 2 * itemb;  // [group][parity] tiles of kAtcQBytes
  uint64_t* bars = reinterpret_cast<uint64_t*>(q_img + 4 * kAtcQBytes + 2048);  // (2 KB of partial row sums in between)
  uint64_t* kv_full = bars;        // [2] loader -> MMA
  uint64_t* kv_empty = bars + 2;   // [2] MMA (last P V of both groups) -> loader
  uint64_t* q_full = bars + 4;     // [2 groups][2] loader -> MMA
  uint64_t* q_empty = bars + 8;    // [2][2] MMA -> loader
  uint64_t* s_full = bars + 12;    // [2] MMA (Q K^T done) -> softmax group
  uint64_t* s_free = bars + 14;    // [2] softmax group (4 warps: every logit of S is in registers) -> MMA
  uint64_t* p_full = bars + 16;    // [2] softmax group (4 warps: P written) -> MMA
  uint64_t* p_free = bars + 18;    // [2] MMA (P V done) -> softmax group
  uint64_t* o_done = bars + 20;    // [2] MMA (last P V of a tile) -> softmax group
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int first = blockIdx.x, stride = gridDim.x;
  const size_t ldq = (size_t)3 * H;
  const int groups_per_item = ntiles >= 2 ? 2 : 1;  // with one tile per item the items alternate between the groups

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], groups_per_item);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 8);
      mbar_init(&p_full[i], 8);
      mbar_init(&p_free[i], 1);
      mbar_init(&o_done[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  // zero once what the loaders never write: the key rows >= S of both K / V images, the tails, the padding d-chunks of the Q tiles
  for (int idx = tid; idx < 2 * 2 * (Spad - S) * CH; idx += kAtcThreads) {
    const int img = idx / ((Spad - S) * CH), r = idx % ((Spad - S) * CH);
    const int key = S + r / CH, c = r % CH;
    *reinterpret_cast<uint4*>(atc_smem + (img >> 1) * itemb + (img & 1) * kvb + (key >> 3) * KG + c * 128 + (key & 7) * 16) = make_uint4(0, 0, 0, 0);
  }
  if (tid < 16) *reinterpret_cast<uint4*>(atc_smem + (tid >> 3) * itemb + 2 * kvb + (tid & 7) * 16) = make_uint4(0, 0, 0, 0);
  if constexpr (CH < 4) {
    constexpr int PADC = 4 - CH;
    for (int idx = tid; idx < 4 * 128 * PADC; idx += kAtcThreads) {
      const int b = idx / (128 * PADC), r = idx % (128 * PADC);
      const int row = r / PADC, c = CH + r % PADC;
      *reinterpret_cast<uint4*>(q_img + b * kAtcQBytes + (row >> 3) * 512 + c * 128 + (row & 7) * 16) = make_uint4(0, 0, 0, 0);
    }
  }
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 2) {
    // ===================================================== MMA issuer of group g = warp =====================================================
    // One issuing warp per group: a tcgen05.mma with a small N costs its issuing warp ~48 cycles whatever the tensor pipe needs for
    // it (N = 32: 16 cycles), so the 8 P V + KSTEPS Q K^T instructions of a chunk step plus the barrier wait and the commits keep it
    // busy for ~600 cycles.  Q K^T of the NEXT step is issued first: the softmax group waits for it, P V only has to finish before
    // the group writes P again.
    const int g = warp;
    constexpr uint32_t idesc_qk = umma_idesc_bf16(128, kAtcChunk);        // A, B K-major
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 32) | (1u << 16);  // B (= V) MN-major
    const uint32_t smem0 = smem_u32(atc_smem), q_addr = smem_u32(q_img);
    const uint32_t tg = tmem_base + g * kAtcGroupCols;
    auto issue_qk = [&](const AtcCursor& cu) {
      if (cu.c == 0) {
        if (cu.first_in_item()) {  // first tile of this group in the item: the item's K / V images must have landed
          mbar_wait(&kv_full[cu.k & 1], (cu.k >> 1) & 1);
        }
        mbar_wait(&q_full[g * 2 + (cu.job & 1)], (cu.job >> 1) & 1);
        tcgen05_fence_after();
      }
      if (elect_one()) {
        const uint32_t qa = q_addr + (g * 2 + (cu.job & 1)) * kAtcQBytes;
        const uint32_t ka = smem0 + (cu.k & 1) * itemb + cu.c * (kAtcChunk / 8) * KG;
#pragma unroll
        for (int j = 0; j < KSTEPS; ++j)
          umma_bf16_ss(tg, umma_desc_nosw(qa + j * 256, 128, 512), umma_desc_nosw(ka + j * 256, 128, KG), idesc_qk, j);
        umma_commit(&s_full[g]);
        if (cu.c == nchunks - 1) umma_commit(&q_empty[g * 2 + (cu.job & 1)]);  // the query tile is no longer read
      }
      __syncwarp();
    };
    AtcCursor cur, nxt;  // cur: next P V; nxt: next Q K^T (one chunk step ahead)
    cur.init(g, first, stride, n_items, ntiles);
    nxt = cur;
    if (nxt.valid) {
      issue_qk(nxt);
      nxt.advance(first, stride, n_items, ntiles, nchunks);
    }
    ATC_TIC();
    while (cur.valid) {
      if (nxt.valid) {
        mbar_wait(&s_free[g], cur.n & 1);  // every logit of S(n) is in registers (three quarters into the exponent phase)
        tcgen05_fence_after();
        ATC_TOC(0);
        issue_qk(nxt);
        nxt.advance(first, stride, n_items, ntiles, nchunks);
        ATC_TOC(2);
      }
      mbar_wait(&p_full[g], cur.n & 1);  // P(n) is written
      tcgen05_fence_after();
      ATC_TOC(1);
      if (elect_one()) {
        const uint32_t va = smem0 + (cur.k & 1) * itemb + kvb + cur.c * (kAtcChunk / 8) * KG;
#pragma unroll
        for (int s = 0; s < kAtcChunk / 16; ++s)
          umma_bf16_ts(tg + kAtcOCol, tg + (s < 4 ? kAtcPLoCol + (cur.n & 1) * 32 + s * 8 : kAtcPHiCol + (s - 4) * 8),
                       umma_desc_nosw(va + 2 * s * KG, KG, 128), idesc_pv, (cur.c | s) != 0);
        umma_commit(&p_free[g]);
        if (cur.c == nchunks - 1) {
          umma_commit(&o_done[g]);
          if (cur.last_in_item(ntiles)) umma_commit(&kv_empty[cur.k & 1]);  // this group's last use of the item's K / V images
        }
      }
      __syncwarp();
      cur.advance(first, stride, n_items, ntiles, nchunks);
      ATC_TOC(3);
    }
    if (TRACE && g == 0 && lane == 0 && trace)
      for (int i = 0; i < 8; ++i) trace[blockIdx.x * 32 + i] = tr[i];
  } else if (warp == 2) {
    // ===================================================== K / V loader =====================================================
    int k = 0;
    for (int item = first; item < n_items; item += stride, ++k) {
      const int z = item / heads, hh = item % heads;
      const long long base = sm.base(z);
      const __nv_bfloat16* kptr = qkv + H + hh * HD;
      const __nv_bfloat16* vptr = qkv + 2 * H + hh * HD;
      uint8_t* k_img = atc_smem + (k & 1) * itemb;
      uint8_t* v_img = k_img + kvb;
      mbar_wait(&kv_empty[k & 1], ((k >> 1) & 1) ^ 1);
      for (int idx = lane; idx < S * CH; idx += 32) {
        const int key = idx / CH, c = idx % CH;
        const uint32_t off = (key >> 3) * KG + c * 128 + (key & 7) * 16;
        const size_t tok = (size_t)(base + (long long)key * sm.seq_stride) * ldq + c * 8;
        cp_async16(k_img + off, kptr + tok, true);
        cp_async16(v_img + off, vptr + tok, true);
      }
      cp_async_commit();
      cp_async_wait<0>();
      fence_proxy_async();  // cp.async writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&kv_full[k & 1]);
    }
  } else if (warp == 3) {
    // ===================================================== Q loader (both groups, in tile-stream order) =====================================================
    AtcCursor cq[2];
    cq[0].init(0, first, stride, n_items, ntiles);
    cq[1].init(1, first, stride, n_items, ntiles);
    while (cq[0].valid || cq[1].valid) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        AtcCursor& cu = cq[g];
        if (!cu.valid) continue;
        const int z = cu.item / heads, hh = cu.item % heads;
        const long long base = sm.base(z);
        const __nv_bfloat16* qptr = qkv + hh * HD;
        const int b = g * 2 + (cu.job & 1);
        mbar_wait(&q_empty[b], ((cu.job >> 1) & 1) ^ 1);
        for (int idx = lane; idx < 128 * CH; idx += 32) {
          const int row = idx / CH, c = idx % CH;
          const int qrow = cu.t * 128 + row;
          const bool ok = qrow < S;
          uint8_t* dst = q_img + b * kAtcQBytes + (row >> 3) * 512 + c * 128 + (row & 7) * 16;
          cp_async16(dst, qptr + (size_t)(base + (long long)(ok ? qrow : 0) * sm.seq_stride) * ldq + c * 8, ok);
        }
        cp_async_commit();
        cp_async_wait<0>();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&q_full[b]);
        for (int c = 0; c < nchunks; ++c) cu.advance(first, stride, n_items, ntiles, nchunks);  // next tile of this group
      }
    }
  } else {
    // ===================================================== softmax group g: thread = (query row, half of the chunk's keys) =====================================================
    // 8 warps per group, 4 softmax warps per scheduler (two of each group): a warp spends ~500 cycles per chunk step in hand-shakes
    // (mbarrier polls, tcgen05.wait, fences, arrivals — each 50 - 100 cycles, in series with its own exponentials); with only two
    // warps per scheduler, one per group, the groups lock in phase and those cycles are idle pipes.  Four warps cover each other.
    const int g = (warp - 4) >> 3;
    const int half = ((warp - 4) >> 2) & 1;   // keys 64 half .. 64 half + 63 of every 128-key chunk
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_t = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + g * kAtcGroupCols;
    const uint32_t s_col = lane_t + half * 64, o_col = lane_t + kAtcOCol + half * 16;
    float* lsum_x = reinterpret_cast<float*>(q_img + 4 * kAtcQBytes) + (g * 2) * 128;  // [2 halves][128 rows] partial row sums of the group
    AtcCursor cu;
    cu.init(g, first, stride, n_items, ntiles);
    uint64_t lsum = f32x2_pack(0.f, 0.f);
    while (cu.valid) {
      uint32_t ra[16], rb[16], pk[8];
      ATC_TIC();
      mbar_wait(&s_full[g], cu.n & 1);
      tcgen05_fence_after();
      tmem_ld16(s_col, ra);
      tmem_ld16(s_col + 16, rb);
      // P: keys 0 .. 63 are double buffered, keys 64 .. 127 wait for P V of the previous step (polled now, consumed after the first block)
      const uint32_t p_col = half == 0 ? lane_t + kAtcPLoCol + (cu.n & 1) * 32 : lane_t + kAtcPHiCol;
      const bool ok_p = half == 0 ? true : mbar_try_wait(&p_free[g], (cu.n & 1) ^ 1);
      const int nvalid = S - cu.c * kAtcChunk - half * 64;  // keys >= S (zero K rows -> exp2(0) = 1) must not count: last chunk only
      tmem_ld_wait();
      ATC_TOC(0);
      if (nvalid >= 64) {
        atc_exp16<POLY, false>(ra, pk, lsum, 16);
        if (!ok_p) mbar_wait(&p_free[g], (cu.n & 1) ^ 1);
        tcgen05_fence_after();
        tmem_st8(p_col, pk);
        tmem_ld16(s_col + 32, ra);
        atc_exp16<POLY, false>(rb, pk, lsum, 16);
        tmem_st8(p_col + 8, pk);
        tmem_ld16(s_col + 48, rb);
        tmem_ld_wait();
      } else {
        atc_exp16<POLY, true>(ra, pk, lsum, nvalid);
        if (!ok_p) mbar_wait(&p_free[g], (cu.n & 1) ^ 1);
        tcgen05_fence_after();
        tmem_st8(p_col, pk);
        tmem_ld16(s_col + 32, ra);
        atc_exp16<POLY, true>(rb, pk, lsum, nvalid - 16);
        tmem_st8(p_col + 8, pk);
        tmem_ld16(s_col + 48, rb);
        tmem_ld_wait();
      }
      // every logit of this warp's half is in registers: the S buffer can take the next Q K^T while the second half is computed
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[g]);
      if (nvalid >= 64) {
        atc_exp16<POLY, false>(ra, pk, lsum, 16);
        tmem_st8(p_col + 16, pk);
        atc_exp16<POLY, false>(rb, pk, lsum, 16);
        tmem_st8(p_col + 24, pk);
      } else {
        atc_exp16<POLY, true>(ra, pk, lsum, nvalid - 32);
        tmem_st8(p_col + 16, pk);
        atc_exp16<POLY, true>(rb, pk, lsum, nvalid - 48);
        tmem_st8(p_col + 24, pk);
      }
      ATC_TOC(2);
      // ---- hand P to the MMA warp
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[g]);
      ATC_TOC(4);
      if (cu.c == nchunks - 1) {
        // ---- O of this tile: the two key halves of a row exchange their partial sums through shared memory, normalise, and each
        // stores 16 of the 32 accumulator columns as bf16.  (The next tile's first P V — the only MMA that overwrites O — is issued
        // after this group's next p_full arrival, i.e. after these loads have completed.)
        const int z = cu.item / heads, hh = cu.item % heads;
        const int qrow = cu.t * 128 + row;
        float l0, l1;
        f32x2_unpack(lsum, l0, l1);
        lsum = f32x2_pack(0.f, 0.f);
        lsum_x[half * 128 + row] = l0 + l1;
        if (g == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
        else asm volatile("bar.sync 2, 256;" ::: "memory");
        const float inv = 1.f / ((l0 + l1) + lsum_x[(half ^ 1) * 128 + row]);
        mbar_wait(&o_done[g], cu.job & 1);
        tcgen05_fence_after();
        uint32_t ov[16];
        tmem_ld16(o_col, ov);
        tmem_ld_wait();
        tcgen05_fence_before();
        if (qrow < S) {
          __nv_bfloat16* op = out + (size_t)(sm.base(z) + (long long)qrow * sm.seq_stride) * ldo + hh * HD + half * 16;
#pragma unroll
          for (int d = 0; d < 16; d += 8) {
            if (half * 16 + d < HD) {
              uint4 o4;
              o4.x = pack_bf16x2(__uint_as_float(ov[d + 0]) * inv, __uint_as_float(ov[d + 1]) * inv);
              o4.y = pack_bf16x2(__uint_as_float(ov[d + 2]) * inv, __uint_as_float(ov[d + 3]) * inv);
              o4.z = pack_bf16x2(__uint_as_float(ov[d + 4]) * inv, __uint_as_float(ov[d + 5]) * inv);
              o4.w = pack_bf16x2(__uint_as_float(ov[d + 6]) * inv, __uint_as_float(ov[d + 7]) * inv);
              *reinterpret_cast<uint4*>(op + d) = o4;
            }
          }
        }
        // both halves have read the exchanged sums before the next tile's sums are written
        if (g == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
        else asm volatile("bar.sync 2, 256;" ::: "memory");
        ATC_TOC(5);
      }
      cu.advance(first, stride, n_items, ntiles, nchunks);
    }
    if (TRACE && warp == 4 && lane == 0 && trace)  // (group A, first key half, lane quarter 0)
      for (int i = 0; i < 8; ++i) trace[blockIdx.x * 32 + 8 + i] = tr[i];
  }
#undef ATC_TIC
#undef ATC_TOC

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace lam

// tcgen05 / TMEM temporal attention, THREE query tiles in flight per CTA with the probabilities written IN PLACE over the logits.
//
// attn_tc.cuh (two tile groups, thread = row x key half, separate P columns) spends per chunk step and group ~1500 cycles of
// tcgen05.ld / wait / st / mbarrier latency in series in every softmax warp: without a single exponential its pipeline needs 360 us
// per 4AA launch, with them 524 us, against 390 us for the exponent engine alone.  The latency is hidden only by the other warps of
// the same scheduler, and there the two groups give two independent phase streams.  This kernel has three:
//   * thread = whole query row (128 keys per chunk step), 4 softmax warps per group, 3 groups = 12 softmax warps: every scheduler holds
//     one warp of each group, i.e. three warps in unrelated phases;
//   * P(n) (bf16 pairs, 64 columns) is written over the columns of S(n) itself — block j of the probabilities lands on logits the
//     thread already holds in registers — so a group needs 128 (S / P) + 32 (O) = 160 TMEM columns and three groups fit (480);
//   * one MMA-issuing warp polls the three groups: after P(n) of a group it issues O += P(n) V and, right behind it, S(n + 1) = Q K^T (the tensor pipe executes
//     in issue order, so S(n + 1) does not overwrite P(n) before it is read): ONE hand-shake per group and chunk step in each
//     direction (s_full, p_full), no s_free / p_free, no row-sum exchange between threads, no named barriers.
// The MMA round trip (~700 cycles from p_full to s_full) is dead time for the group, but not for the scheduler: the other two groups'
// warps have the pipes.
// MEASURED (B200, 4AA launch): 554 us against 524 us for attn_tc.cuh, 383 us against 360 us without exponentials — the in-place P puts
// S(n + 1) behind the whole of softmax(n) and P V(n) (attn_tc.cuh overlaps Q K^T with the second half of the exponentials) and a
// thread walks 128 keys in series, which lengthens the per-step chain by more than the third phase stream hides.  Validated for
// sequences of >= 3 query tiles (S > 256; with two tiles and several items per CTA it faults, and the launcher refuses).  Kept as a debug
// variant (lamslide_debug_attention mode 3 + 4 * 7 .. 10) and as the record of the experiment; the product path uses attn_tc.cuh.
// Query tiles: a ring of 4 images in tile-stream order (tile u in image u % 4; group g owns the tiles u = g mod 3), filled by a loader
// warp one tile ahead.  K / V images, item walk and the no-running-maximum softmax are those of attn_tc.cuh.
#pragma once
#include "attn_tc.cuh"

namespace lam {

constexpr int kAtc3Threads = 32 * 16;  // MMA issuer, K / V loader, Q loader, (one idle), 3 groups x 4 softmax warps: 128 registers per thread
constexpr int kAtc3GroupCols = 160, kAtc3OCol = 128;

struct Atc3Cursor {
  int u;            // position in the CTA's tile stream (= k * ntiles + t)
  int k, t, c;      // item index in the CTA's list (K / V buffer = k & 1), query tile, key chunk
  int job, n;       // running counts of tiles and chunk steps of this group (barrier parities)
  int item;
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
    c = 0, ++job, u += 3;
    place(first, stride, n_items, ntiles);
  }
  __device__ __forceinline__ bool first_in_item() const { return t < 3; }
  __device__ __forceinline__ bool last_in_item(int ntiles) const { return t + 3 >= ntiles; }
};

template <int HD>
struct Atc3Cfg {
  // 2 item buffers | 4 Q tiles | barriers
  static __host__ __device__ size_t smem_bytes(int S) { return 2 * AtcCfg<HD>::item_bytes(S) + 4 * kAtcQBytes + 256; }
};

template <int HD, int POLY>
__global__ void __launch_bounds__(kAtc3Threads, 1)
attn_tc3_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int H, int ldo, SeqMap sm, int heads, int n_items) {
  using Cfg = AtcCfg<HD>;
  constexpr int CH = HD / 8;
  constexpr int KG = Cfg::KG;
  constexpr int KSTEPS = Cfg::KSTEPS;
  extern __shared__ __align__(1024) uint8_t atc_smem[];
  const int S = sm.S;
  const int Spad = atc_spad(S);
  const int ntiles = (S + 127) / 128, nchunks = Spad / kAtcChunk;
  const uint32_t kvb = (uint32_t)Cfg::kv_bytes(S), itemb = (uint32_t)Cfg::item_bytes(S);
  uint8_t* q_img = atc_smem + 2 * itemb;  // ring of 4 tile images
  uint64_t* bars = reinterpret_cast<uint64_t*>(q_img + 4 * kAtcQBytes);
  uint64_t* kv_full = bars;        // [2] loader -> MMA
  uint64_t* kv_empty = bars + 2;   // [2] MMA (last P V of every group that works on the item) -> loader
  uint64_t* q_full = bars + 4;     // [4] loader -> MMA
  uint64_t* q_empty = bars + 8;    // [4] MMA (last Q K^T of the tile) -> loader
  uint64_t* s_full = bars + 12;    // [3] MMA (Q K^T done) -> softmax group
  uint64_t* p_full = bars + 15;    // [3] softmax group (4 warps: P written) -> MMA
  uint64_t* o_done = bars + 18;    // [3] MMA (last P V of a tile) -> softmax group
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int first = blockIdx.x, stride = gridDim.x;
  const size_t ldq = (size_t)3 * H;
  const int groups_per_item = ntiles >= 3 ? 3 : ntiles;  // every group with a tile in the item arrives once on kv_empty

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], groups_per_item);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
    }
    for (int i = 0; i < 3; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  // zero once what the loaders never write: the key rows >= S of both K / V images, the tails, the padding d-chunks of the Q tiles
  for (int idx = tid; idx < 2 * 2 * (Spad - S) * CH; idx += kAtc3Threads) {
    const int img = idx / ((Spad - S) * CH), r = idx % ((Spad - S) * CH);
    const int key = S + r / CH, c = r % CH;
    *reinterpret_cast<uint4*>(atc_smem + (img >> 1) * itemb + (img & 1) * kvb + (key >> 3) * KG + c * 128 + (key & 7) * 16) = make_uint4(0, 0, 0, 0);
  }
  if (tid < 16) *reinterpret_cast<uint4*>(atc_smem + (tid >> 3) * itemb + 2 * kvb + (tid & 7) * 16) = make_uint4(0, 0, 0, 0);
  if constexpr (CH < 4) {
    constexpr int PADC = 4 - CH;
    for (int idx = tid; idx < 4 * 128 * PADC; idx += kAtc3Threads) {
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

  if (warp == 0) {
    // ===================================================== MMA issuer of all three groups =====================================================
    // Takes whichever group has its P(n) written (and, at a tile boundary, the operands of the next Q K^T in place): ~500 cycles of
    // issue per group and chunk step (8 P V + KSTEPS Q K^T instructions and the commits) against ~2300 cycles per step.
    constexpr uint32_t idesc_qk = umma_idesc_bf16(128, kAtcChunk);        // A, B K-major
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 32) | (1u << 16);  // B (= V) MN-major
    const uint32_t smem0 = smem_u32(atc_smem), q_addr = smem_u32(q_img);
    // operands of a tile's first Q K^T: the item's K / V images (first tile of this group in the item) and the tile's Q image
    auto tile_ready = [&](const Atc3Cursor& cu) {
      if (cu.first_in_item() && !mbar_test_wait(&kv_full[cu.k & 1], (cu.k >> 1) & 1)) return false;
      return mbar_test_wait(&q_full[cu.u & 3], (cu.u >> 2) & 1);
    };
    auto issue_qk = [&](const Atc3Cursor& cu, int g) {  // one elected lane
      const uint32_t tg = tmem_base + g * kAtc3GroupCols;
      const uint32_t qa = q_addr + (cu.u & 3) * kAtcQBytes;
      const uint32_t ka = smem0 + (cu.k & 1) * itemb + cu.c * (kAtcChunk / 8) * KG;
#pragma unroll
      for (int j = 0; j < KSTEPS; ++j)
        umma_bf16_ss(tg, umma_desc_nosw(qa + j * 256, 128, 512), umma_desc_nosw(ka + j * 256, 128, KG), idesc_qk, j);
      umma_commit(&s_full[g]);
      if (cu.c == nchunks - 1) umma_commit(&q_empty[cu.u & 3]);  // the query tile is no longer read
    };
    Atc3Cursor cur[3], nxt[3];  // cur: next P V; nxt: next Q K^T (one chunk step ahead)
    bool started[3];
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      cur[g].init(g, first, stride, n_items, ntiles);
      nxt[g] = cur[g];
      started[g] = false;
    }
    uint32_t idle_polls = 0;  // a protocol bug becomes a trapped launch, never a hung GPU
    uint64_t idle_t0 = 0;
    while (cur[0].valid || cur[1].valid || cur[2].valid) {
      if ((++idle_polls & 0xffff) == 0) {
        uint64_t t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (idle_t0 == 0) idle_t0 = t1;
        if (t1 - idle_t0 > 4000000000ull) __trap();  // 4 s without the kernel finishing
      }
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        if (!cur[g].valid) continue;
        if (!started[g]) {  // the group's very first Q K^T
          if (!tile_ready(nxt[g])) continue;
          tcgen05_fence_after();
          if (elect_one()) issue_qk(nxt[g], g);
          __syncwarp();
          nxt[g].advance(first, stride, n_items, ntiles, nchunks);
          started[g] = true;
          continue;
        }
        if (nxt[g].valid && nxt[g].c == 0 && !tile_ready(nxt[g])) continue;  // (prefetched a tile ahead: normally ready)
        if (!mbar_test_wait(&p_full[g], cur[g].n & 1)) continue;            // P(n) is written (over S(n))
        tcgen05_fence_after();
        if (elect_one()) {
          const Atc3Cursor& c = cur[g];
          const uint32_t tg = tmem_base + g * kAtc3GroupCols;
          const uint32_t va = smem0 + (c.k & 1) * itemb + kvb + c.c * (kAtcChunk / 8) * KG;
#pragma unroll
          for (int s = 0; s < kAtcChunk / 16; ++s)
            umma_bf16_ts(tg + kAtc3OCol, tg + s * 8, umma_desc_nosw(va + 2 * s * KG, KG, 128), idesc_pv, (c.c | s) != 0);
          if (c.c == nchunks - 1) {
            umma_commit(&o_done[g]);
            if (c.last_in_item(ntiles)) umma_commit(&kv_empty[c.k & 1]);  // this group's last use of the item's K / V images
          }
          if (nxt[g].valid) issue_qk(nxt[g], g);  // executes behind P V: S(n + 1) may overwrite P(n)
        }
        __syncwarp();
        cur[g].advance(first, stride, n_items, ntiles, nchunks);
        if (nxt[g].valid) nxt[g].advance(first, stride, n_items, ntiles, nchunks);
      }
    }
  } else if (warp == 1) {
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
  } else if (warp == 2) {
    // ===================================================== Q loader: every tile of the stream, in order =====================================================
    for (int u = 0;; ++u) {
      const int k = u / ntiles, t = u - k * ntiles;
      const int item = first + k * stride;
      if (item >= n_items) break;
      const int z = item / heads, hh = item % heads;
      const long long base = sm.base(z);
      const __nv_bfloat16* qptr = qkv + hh * HD;
      const int b = u & 3;
      mbar_wait(&q_empty[b], ((u >> 2) & 1) ^ 1);
      for (int idx = lane; idx < 128 * CH; idx += 32) {
        const int row = idx / CH, c = idx % CH;
        const int qrow = t * 128 + row;
        const bool ok = qrow < S;
        uint8_t* dst = q_img + b * kAtcQBytes + (row >> 3) * 512 + c * 128 + (row & 7) * 16;
        cp_async16(dst, qptr + (size_t)(base + (long long)(ok ? qrow : 0) * sm.seq_stride) * ldq + c * 8, ok);
      }
      cp_async_commit();
      cp_async_wait<0>();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&q_full[b]);
    }
  } else if (warp >= 4) {
    // ===================================================== softmax group g: thread = query row =====================================================
    const int g = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t s_col = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + g * kAtc3GroupCols;
    const uint32_t o_col = s_col + kAtc3OCol;
    Atc3Cursor cu;
    cu.init(g, first, stride, n_items, ntiles);
    uint64_t lsum = f32x2_pack(0.f, 0.f);
    while (cu.valid) {
      uint32_t ra[16], rb[16], pk[8];
      mbar_wait(&s_full[g], cu.n & 1);
      tcgen05_fence_after();
      tmem_ld16(s_col, ra);
      tmem_ld16(s_col + 16, rb);
      const int nvalid = S - cu.c * kAtcChunk;  // keys >= S (zero K rows -> exp2(0) = 1) must not count: last chunk only
      tmem_ld_wait();
      // 8 blocks of 16 logits; the probabilities of block j go to columns 8 j .. 8 j + 7, i.e. over logits of block j / 2 <= j that are
      // already in registers
      if (nvalid >= kAtcChunk) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          atc_exp16<POLY, false>(ra, pk, lsum, 16);
          tmem_st8(s_col + 8 * j, pk);
          if (j + 2 < 8) tmem_ld16(s_col + 16 * (j + 2), ra);
          atc_exp16<POLY, false>(rb, pk, lsum, 16);
          tmem_st8(s_col + 8 * (j + 1), pk);
          if (j + 2 < 8) {
            tmem_ld16(s_col + 16 * (j + 3), rb);
            tmem_ld_wait();
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          atc_exp16<POLY, true>(ra, pk, lsum, nvalid - 16 * j);
          tmem_st8(s_col + 8 * j, pk);
          if (j + 2 < 8) tmem_ld16(s_col + 16 * (j + 2), ra);
          atc_exp16<POLY, true>(rb, pk, lsum, nvalid - 16 * (j + 1));
          tmem_st8(s_col + 8 * (j + 1), pk);
          if (j + 2 < 8) {
            tmem_ld16(s_col + 16 * (j + 3), rb);
            tmem_ld_wait();
          }
        }
      }
      // ---- hand P to the MMA warp
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[g]);
      if (cu.c == nchunks - 1) {
        // ---- O of this tile: scale by 1 / row sum and store bf16.  (The next tile's first P V — the only MMA that overwrites O — is
        // issued after this group's next p_full arrival, i.e. after these loads have completed.)
        const int z = cu.item / heads, hh = cu.item % heads;
        const int qrow = cu.t * 128 + row;
        float l0, l1;
        f32x2_unpack(lsum, l0, l1);
        lsum = f32x2_pack(0.f, 0.f);
        const float inv = 1.f / (l0 + l1);
        mbar_wait(&o_done[g], cu.job & 1);
        tcgen05_fence_after();
        uint32_t ov[32];
        tmem_ld16(o_col, ov);
        tmem_ld16(o_col + 16, ov + 16);
        tmem_ld_wait();
        tcgen05_fence_before();
        if (qrow < S) {
          __nv_bfloat16* op = out + (size_t)(sm.base(z) + (long long)qrow * sm.seq_stride) * ldo + hh * HD;
#pragma unroll
          for (int d = 0; d < HD; d += 8) {
            uint4 o4;
            o4.x = pack_bf16x2(__uint_as_float(ov[d + 0]) * inv, __uint_as_float(ov[d + 1]) * inv);
            o4.y = pack_bf16x2(__uint_as_float(ov[d + 2]) * inv, __uint_as_float(ov[d + 3]) * inv);
            o4.z = pack_bf16x2(__uint_as_float(ov[d + 4]) * inv, __uint_as_float(ov[d + 5]) * inv);
            o4.w = pack_bf16x2(__uint_as_float(ov[d + 6]) * inv, __uint_as_float(ov[d + 7]) * inv);
            *reinterpret_cast<uint4*>(op + d) = o4;
          }
        }
      }
      cu.advance(first, stride, n_items, ntiles, nchunks);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace lam

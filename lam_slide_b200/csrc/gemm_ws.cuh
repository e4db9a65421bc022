// Persistent, warp-specialised tcgen05 GEMM for the big projections of the second-stage transformer
// (mmdit.py:240-249, latent_si_v31.py:172):  D[rows, N] = A[rows, K] · B[N, K]^T, bf16 operands, fp32 accumulation in TMEM.
//
// One CTA per SM (grid = min(#SM, #m-blocks)); a CTA walks m-blocks  mb = blockIdx.x, + gridDim.x, ...  and, inside an
// m-block, all n-tiles, so the 128 x K activation tile is fetched ONCE per m-block when it fits (A-resident mode,
// K <= 512: linear1) and only the weight tiles stream through the TMA ring from L2.  Warp roles (576 threads):
//   warp 0      : TMA producer (one lane): A k-blocks (resident or ring) + B ring, mbarrier complete_tx
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer, TWO accumulator stages (2 x BN TMEM columns)
//   warps 2..17 : epilogue.  All 16 warps work on the same tile: warp w owns TMEM lanes 32*(w%4).. (hardware rule) and
//                 the column quarter (w-2)/4, which it walks in chunks of Epi::CW columns: tcgen05.ld -> registers ->
//                 fused math -> padded shared-memory box (thread-per-row writes, conflict-free) -> row-contiguous
//                 16-byte global stores by the whole warp (linear1, embedding) or a TMA f32 reduce-add (linear2:
//                 h += gate * (acc + bias), performed at L2).  The TMEM stage is released right after the LAST chunk has
//                 been pulled into registers, i.e. before the tile's epilogue is finished, so the MMA warp (already
//                 busy with the other stage) never waits for it.
// Every mbarrier wait is bounded (ptx.cuh: mbar_wait) so a protocol error is a failed launch, never a hung GPU.
#pragma once
#include "gemm_tc.cuh"

namespace lam {

constexpr int kWsEpiWarps = 16;                      // 4 TMEM lane quarters x 4 column quarters
constexpr int kWsThreads = 64 + 32 * kWsEpiWarps;   // + TMA warp + MMA warp
constexpr int kWsStageBytesPerWarp = 3072;           // one staged box (<= 32 rows x 96 B) per warp
constexpr int kWsMaxKBlocksResident = 8;

struct WsCtx {
  const CUtensorMap* o0;
  const CUtensorMap* o1;
  uint32_t stage_s;  // this warp's staging box (32-bit shared address)
  uint32_t smf_s;    // per-kernel fp32 constants in shared memory (bias, ...), 32-bit shared address
  int lane;
  int row0;          // first global row of this warp's 32-row slice
};

__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
  const uint4 u = ld_shared_v4(addr);
  return make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
}

// byte offset of (row r, 16-byte chunk c) inside a TMA-staged box with 64-byte rows (SWIZZLE_64B on the host side):
// the 8 threads of a quarter warp hit 8 distinct 16-byte bank groups.
__device__ __forceinline__ uint32_t stage_off64(int r, int c) { return r * 64 + ((c ^ ((r >> 1) & 3)) << 4); }

// ------------------------------------------------------------------------------------------------ linear1 epilogue
// Columns [0,3H) are (K=3, heads, hd) q|k|v, columns [3H,3H+M) the MLP (mmdit.py:241-247):
//   q,k : + bias -> RMSNorm over the head (fp32, eps 1e-6, * scale; mmdit.py:132-136) -> RoPE on interleaved pairs
//         (mmdit.py:85-90) -> q additionally * (hd^-0.5 * log2 e) -> bf16 -> qkv[rows, 3H]
//   v   : + bias -> bf16 -> qkv
//   mlp : + bias -> erf-GELU (mmdit.py:11-18; ptx.cuh gelu_fast) -> bf16 -> act[:, H + j]   (A operand of linear2)
// A chunk is one head (HD columns).  Output: thread-per-row packs the chunk into a padded shared-memory box (pitch = odd
// number of 16-byte units), then the warp stores the box with row-contiguous 16-byte st.global (HD * 2 bytes per row).
// AL > 0 (spatial blocks, sequences of AL CONSECUTIVE rows, AL a power of two <= 8, M = 0): the attention itself runs in this epilogue
// (latent_si_v31.py:51-54, mmdit.py:42-55 for S = L).  The n-tiles are walked as (q, k, v) triples over the same heads; thread = row
// keeps its normalised / rotated q (bf16) in the warp's shared-memory scratch, gets the k and v rows of its sequence from the
// neighbouring lanes with shuffles, and writes softmax(q k^T) v straight into act[:, :H] — q, k and v of a spatial block never go to
// HBM and the separate attention launch disappears (4AA: 87 us and ~670 MB per block).
template <int HD, int AL = 0>
struct EpiLinear1Ws {
  struct Params {
    const float* bias;      // [3H + M]
    float gam[2][HD];       // [0]: query_norm.scale * hd^-0.5 * log2(e), [1]: key_norm.scale  (by value: constant-bank operands)
    const float* rope_cos;  // [S, HD/2]
    const float* rope_sin;
    __nv_bfloat16* qkv;     // [rows, 3H]
    __nv_bfloat16* act;     // [rows, H + M]
    int H, M, rows;
    int pos_div, pos_mod;   // rope position of a row = (row / pos_div) % pos_mod
    int debug;              // profiling aid (lamslide_debug_linear1): bit 0 = skip the global stores, bit 1 = skip the epilogue math
    int act_ld;             // row pitch of act in elements (H + mlp hidden; M above is 0 when the MLP half runs in the fused kernel)
  };
  static constexpr int CW = HD;                                  // chunk width (columns)
  static constexpr int CH = HD / 8;                              // 16-byte units per chunk row
  // Global stores must cover whole 32-byte sectors (partial-sector writes ran at ~1.8 TB/s into L2 on B200 and bounded the
  // kernel): when a chunk row (HD * 2 bytes) is not a multiple of 32 bytes, two consecutive chunks are staged side by side
  // and written out together.
  static constexpr bool kPair = (HD * 2) % 32 != 0;
  // AL mode scratch: 32 rows x (heads of a warp's column quarter) x HD bf16 — q, later the probabilities.  Paired boxes (hd = 24)
  // have exactly that shape and are reused; the other layouts get 32 x 64 bytes behind the staging box.
  static constexpr int kScratchOff = (AL > 0 && !kPair) ? kWsStageBytesPerWarp : 0;
  static constexpr int kStageBytes = kWsStageBytesPerWarp + ((AL > 0 && !kPair) ? 2048 : 0);
  static_assert(AL == 0 || (AL & (AL - 1)) == 0, "sequence length must be a power of two");
  static_assert(AL * 4 <= HD * 2, "the probabilities of a head reuse its q slot");
  static constexpr int UNITS = kPair ? 2 * CH : CH;             // 16-byte units per staged row
  static constexpr int PITCH = (kPair ? UNITS : (UNITS % 2 == 0 ? UNITS + 1 : UNITS)) * 16;  // bytes (pairs: dense, 2-way conflicts on the writes only)
  static_assert(32 * PITCH <= kWsStageBytesPerWarp, "staging box too large");
  // Paired boxes are dense 32 x (2 HD) bf16 tiles: they go out as ONE TMA tensor store per warp and tile half, so the epilogue
  // warps never wait on the LSU store path (with st.global the stores and the math did not overlap: +53 us and +56 us on a
  // 187 us floor gave 295 us).  o0 / o1 = store maps of qkv / act with a {2 HD, 32} box, no swizzle.
  static constexpr bool kTmaStore = kPair;
  struct Tile {
    int kind;  // 0 q, 1 k, 2 v, 3 mlp
    const float4 *cs, *sn;  // this row's RoPE table entries (L1-resident: re-read per head rather than held in 24 registers)
    int qrow_bytes;         // AL mode: bytes per scratch row = heads per column quarter * HD * 2
  };
  static __host__ __device__ int smem_floats(const Params& p) { return 3 * p.H + p.M; }
  static __device__ void load_consts(const Params& p, float* smf, int tid, int nthreads) {
    const int N = 3 * p.H + p.M;
    for (int i = tid; i < N; i += nthreads) smf[i] = p.bias[i];
  }
  // n-tile order: MLP (MUFU-heavy epilogue) and q/k/v tiles alternate so the epilogue load is even over time
  template <int BN>
  static __device__ __forceinline__ int tile_n0(const Params& p, int nt) {
    if constexpr (AL > 0) return (nt % 3) * p.H + (nt / 3) * BN;  // (q, k, v) triples over the same BN / HD heads
    const int nq = 3 * p.H / BN, nm = p.M / BN;
    const int pairs = nq < nm ? nq : nm;
    if (nt < 2 * pairs) return (nt & 1) ? (nt >> 1) * BN : 3 * p.H + (nt >> 1) * BN;
    const int r = nt - 2 * pairs + pairs;
    return nq > nm ? r * BN : 3 * p.H + r * BN;
  }
  static __device__ __forceinline__ void tile_begin(const Params& p, const WsCtx&, Tile& t, int row, int n0w, int qw) {
    t.qrow_bytes = qw * 2;
    t.kind = n0w >= 3 * p.H ? 3 : n0w / p.H;  // BN divides H, so a tile — and a slice of it — is one kind
    const int pos = row < p.rows ? (row / p.pos_div) % p.pos_mod : 0;
    t.cs = reinterpret_cast<const float4*>(p.rope_cos + (size_t)pos * (HD / 2));
    t.sn = reinterpret_cast<const float4*>(p.rope_sin + (size_t)pos * (HD / 2));
  }
  // packed chunk (HD bf16 of this thread's row, chunk index ck inside the warp's column quarter) -> staging box ->
  // row-contiguous, sector-aligned 16-byte global stores by the whole warp.  (Direct per-thread row stores were measured at
  // 474 us per 4AA launch against 391 us staged; small-box TMA stores were no better.)
  static __device__ __forceinline__ void emit(const Params& p, const WsCtx& c, const uint32_t* w, __nv_bfloat16* out, int ld, int col, int ck) {
    const int half = kPair ? (ck & 1) : 0;
    if (half == 0) {
      if (kTmaStore && c.lane == 0) bulk_wait_read<0>();  // the TMA unit has read the previous box out of shared memory
      __syncwarp();                                       // ... / the previous box has been read out by every lane
    }
#pragma unroll
    for (int ch = 0; ch < CH; ++ch)
      st_shared_v4(c.stage_s + c.lane * PITCH + (half * CH + ch) * 16, w[4 * ch], w[4 * ch + 1], w[4 * ch + 2], w[4 * ch + 3]);
    if (kPair && half == 0) return;
    const int col0 = col - half * HD;  // first column of the staged row
    if constexpr (kTmaStore) {
      fence_proxy_async();  // generic-proxy writes of every lane -> visible to the async proxy (TMA)
      __syncwarp();
      if (c.lane == 0 && !(p.debug & 1)) {
        if (p.debug & 4) tma_store_2d_s_hint(out == p.qkv ? c.o0 : c.o1, c.stage_s, col0, c.row0, l2_policy_evict_first());
        else tma_store_2d_s(out == p.qkv ? c.o0 : c.o1, c.stage_s, col0, c.row0);
        bulk_commit();
      }
      return;
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < UNITS; ++k) {
      const int id = c.lane + 32 * k;
      const int r = id / UNITS, ch = id % UNITS;
      const uint4 val = ld_shared_v4(c.stage_s + r * PITCH + ch * 16);
      if (c.row0 + r < p.rows && !(p.debug & 1)) *reinterpret_cast<uint4*>(out + (size_t)(c.row0 + r) * ld + col0 + ch * 8) = val;
    }
  }
  // v: HD accumulators of this thread's row, columns col .. col + HD
  static __device__ __forceinline__ void chunk(const Params& p, const WsCtx& c, const Tile& t, const uint32_t* v, int col, int ck) {
    const int H3 = 3 * p.H;
    const uint32_t bias_s = c.smf_s + col * 4;
    uint32_t w[HD / 2];
    if (p.debug & 2) {  // raw accumulators, no math
#pragma unroll
      for (int j = 0; j < HD / 2; ++j) w[j] = pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
      if (t.kind == 3) emit(p, c, w, p.act, p.H + p.M, p.H + (col - H3), ck);
      else emit(p, c, w, p.qkv, H3, col, ck);
      return;
    }
    if (t.kind == 3) {  // ---- MLP: GELU
#pragma unroll
      for (int j = 0; j < HD / 4; ++j) {
        const float4 bv = ld_shared_f4(bias_s + j * 16);
        const float y0 = gelu_fast(__uint_as_float(v[4 * j + 0]) + bv.x);
        const float y1 = gelu_fast(__uint_as_float(v[4 * j + 1]) + bv.y);
        const float y2 = gelu_fast(__uint_as_float(v[4 * j + 2]) + bv.z);
        const float y3 = gelu_fast(__uint_as_float(v[4 * j + 3]) + bv.w);
        w[2 * j] = pack_bf16x2(y0, y1);
        w[2 * j + 1] = pack_bf16x2(y2, y3);
      }
      emit(p, c, w, p.act, p.H + p.M, p.H + (col - H3), ck);
      return;
    }
    if (t.kind == 2) {  // ---- v: bias only
#pragma unroll
      for (int j = 0; j < HD / 4; ++j) {
        const float4 bv = ld_shared_f4(bias_s + j * 16);
        w[2 * j] = pack_bf16x2(__uint_as_float(v[4 * j + 0]) + bv.x, __uint_as_float(v[4 * j + 1]) + bv.y);
        w[2 * j + 1] = pack_bf16x2(__uint_as_float(v[4 * j + 2]) + bv.z, __uint_as_float(v[4 * j + 3]) + bv.w);
      }
      if constexpr (AL > 0) {
        // ---- out = sum_j p_j v_j over the AL rows of this row's sequence (v_j from the neighbouring lanes), written as the
        // attention half of act (mmdit.py:248: cat(attn, gelu(mlp)))
        const uint32_t slot = c.stage_s + kScratchOff + c.lane * t.qrow_bytes + ck * (HD * 2);
        float pr[AL];
#pragma unroll
        for (int j = 0; j < AL; j += 4) {
          const float4 pv = ld_shared_f4(slot + j * 4);
          pr[j] = pv.x;
          if (j + 1 < AL) pr[j + 1] = pv.y;
          if (j + 2 < AL) pr[j + 2] = pv.z;
          if (j + 3 < AL) pr[j + 3] = pv.w;
        }
        float acc[HD];
#pragma unroll
        for (int d = 0; d < HD; ++d) acc[d] = 0.f;
        const int seq0 = c.lane & ~(AL - 1);
#pragma unroll
        for (int j = 0; j < AL; ++j) {
#pragma unroll
          for (int i = 0; i < HD / 2; ++i) {
            const uint32_t vw = __shfl_sync(0xffffffffu, w[i], seq0 + j);
            acc[2 * i] = fmaf(pr[j], __uint_as_float(vw << 16), acc[2 * i]);
            acc[2 * i + 1] = fmaf(pr[j], __uint_as_float(vw & 0xffff0000u), acc[2 * i + 1]);
          }
        }
#pragma unroll
        for (int i = 0; i < HD / 2; ++i) w[i] = pack_bf16x2(acc[2 * i], acc[2 * i + 1]);
        emit(p, c, w, p.act, p.act_ld, col - 2 * p.H, ck);
        return;
      }
      emit(p, c, w, p.qkv, H3, col, ck);
      return;
    }
    // ---- q or k: RMSNorm + RoPE over the head.  Two instantiations so that the scales are compile-time-indexed constant-bank
    // operands of the FMULs (a run-time q / k select made them 24 indexed LDC per chunk: 18 % of the kernel's stall samples).
    if (t.kind == 0) rms_rope<0>(p, t, v, bias_s, w);
    else rms_rope<1>(p, t, v, bias_s, w);
    if constexpr (AL > 0) {
      const uint32_t slot = c.stage_s + kScratchOff + c.lane * t.qrow_bytes + ck * (HD * 2);
      if (t.kind == 0) {  // ---- q (already * hd^-0.5 * log2 e): park it until the k tile of the same heads arrives
        if (ck == 0) {
          if (kTmaStore && c.lane == 0) bulk_wait_read<0>();  // the previous triple's output box has been read by the TMA unit
          __syncwarp();
        }
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) st_shared_v4(slot + ch * 16, w[4 * ch], w[4 * ch + 1], w[4 * ch + 2], w[4 * ch + 3]);
        return;
      }
      // ---- k: logits of this row against the AL keys of its sequence, softmax in the exp2 domain, probabilities over the q slot
      float q[HD];
#pragma unroll
      for (int ch = 0; ch < CH; ++ch) {
        const uint4 qv = ld_shared_v4(slot + ch * 16);
        const uint32_t qq[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          q[ch * 8 + 2 * i] = __uint_as_float(qq[i] << 16);
          q[ch * 8 + 2 * i + 1] = __uint_as_float(qq[i] & 0xffff0000u);
        }
      }
      float sc[AL];
      const int seq0 = c.lane & ~(AL - 1);
#pragma unroll
      for (int j = 0; j < AL; ++j) {
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < HD / 2; ++i) {
          const uint32_t kw = __shfl_sync(0xffffffffu, w[i], seq0 + j);
          dot = fmaf(q[2 * i], __uint_as_float(kw << 16), dot);
          dot = fmaf(q[2 * i + 1], __uint_as_float(kw & 0xffff0000u), dot);
        }
        sc[j] = dot;
      }
      float mx = sc[0];
#pragma unroll
      for (int j = 1; j < AL; ++j) mx = fmaxf(mx, sc[j]);
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < AL; ++j) {
        sc[j] = fast_exp2(sc[j] - mx);
        sum += sc[j];
      }
      const float inv = 1.f / sum;
#pragma unroll
      for (int j = 0; j < AL; j += 4)
        st_shared_v4(slot + j * 4, __float_as_uint(sc[j] * inv), __float_as_uint(j + 1 < AL ? sc[j + 1] * inv : 0.f),
                     __float_as_uint(j + 2 < AL ? sc[j + 2] * inv : 0.f), __float_as_uint(j + 3 < AL ? sc[j + 3] * inv : 0.f));
      return;
    }
    emit(p, c, w, p.qkv, H3, col, ck);
  }
  template <int IS_K>
  static __device__ __forceinline__ void rms_rope(const Params& p, const Tile& t, const uint32_t* v, uint32_t bias_s, uint32_t* w) {
    float x[HD];
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < HD / 4; ++j) {
      const float4 bv = ld_shared_f4(bias_s + j * 16);
      x[4 * j + 0] = __uint_as_float(v[4 * j + 0]) + bv.x;
      x[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + bv.y;
      x[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + bv.z;
      x[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + bv.w;
      ss = fmaf(x[4 * j + 0], x[4 * j + 0], ss);
      ss = fmaf(x[4 * j + 1], x[4 * j + 1], ss);
      ss = fmaf(x[4 * j + 2], x[4 * j + 2], ss);
      ss = fmaf(x[4 * j + 3], x[4 * j + 3], ss);
    }
    const float rstd = rsqrtf(ss * (1.0f / HD) + 1e-6f);
#pragma unroll
    for (int j2 = 0; j2 < HD / 8; ++j2) {  // 8 columns = 4 rotation pairs per step
      const float4 cv = __ldg(t.cs + j2), sv = __ldg(t.sn + j2);
      const float cc[4] = {cv.x * rstd, cv.y * rstd, cv.z * rstd, cv.w * rstd};
      const float sc[4] = {sv.x * rstd, sv.y * rstd, sv.z * rstd, sv.w * rstd};
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const int j = 2 * j2 + h2;
        const float e0 = x[4 * j + 0] * p.gam[IS_K][4 * j], d0 = x[4 * j + 1] * p.gam[IS_K][4 * j + 1];
        const float e1 = x[4 * j + 2] * p.gam[IS_K][4 * j + 2], d1 = x[4 * j + 3] * p.gam[IS_K][4 * j + 3];
        const float c0 = cc[2 * h2], s0 = sc[2 * h2], c1 = cc[2 * h2 + 1], s1 = sc[2 * h2 + 1];
        w[2 * j] = pack_bf16x2(fmaf(c0, e0, -s0 * d0), fmaf(s0, e0, c0 * d0));
        w[2 * j + 1] = pack_bf16x2(fmaf(c1, e1, -s1 * d1), fmaf(s1, e1, c1 * d1));
      }
    }
  }
  static __device__ __forceinline__ void finish(const WsCtx& c) {
    if (kTmaStore && c.lane == 0) bulk_wait_read<0>();  // staged boxes must be read out before the CTA's shared memory goes away
  }
};

// ------------------------------------------------------------------------------------------------ linear2 epilogue
// linear2 + gated residual (mmdit.py:248, latent_si_v31.py:54,61):  h[row, n] += gate[b(row), n] * (acc + bias[n]),
// issued as a TMA f32 reduce-add (16-column boxes, 64-byte rows, SWIZZLE_64B) so the SM never reads h.
struct EpiLinear2Ws {
  struct Params {
    const float* bias;  // [H]
    const float* gate;  // gate of sample b at gate + b * gate_stride, [H]
    int gate_stride;
    int rows_per_sample;  // T * L
    int H, rows;
  };
  static constexpr int CW = 16;
  static constexpr int kStageBytes = 2048;  // one 32-row x 64-byte box
  struct Tile {
    const float* gate;
  };
  static __host__ __device__ int smem_floats(const Params& p) { return p.H; }
  static __device__ void load_consts(const Params& p, float* smf, int tid, int nthreads) {
    for (int i = tid; i < p.H; i += nthreads) smf[i] = p.bias[i];
  }
  template <int BN>
  static __device__ __forceinline__ int tile_n0(const Params&, int nt) { return nt * BN; }
  static __device__ __forceinline__ void tile_begin(const Params& p, const WsCtx&, Tile& t, int row, int, int) {
    const int b = (row < p.rows ? row : p.rows - 1) / p.rows_per_sample;
    t.gate = p.gate + (size_t)b * p.gate_stride;
  }
  static __device__ __forceinline__ void chunk(const Params&, const WsCtx& c, const Tile& t, const uint32_t* v, int col, int) {
    float4 o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 gv = __ldg(reinterpret_cast<const float4*>(t.gate + col) + j);
      const float4 bv = ld_shared_f4(c.smf_s + col * 4 + j * 16);
      o[j].x = gv.x * (__uint_as_float(v[4 * j + 0]) + bv.x);
      o[j].y = gv.y * (__uint_as_float(v[4 * j + 1]) + bv.y);
      o[j].z = gv.z * (__uint_as_float(v[4 * j + 2]) + bv.z);
      o[j].w = gv.w * (__uint_as_float(v[4 * j + 3]) + bv.w);
    }
    if (c.lane == 0) bulk_wait_read<0>();  // the previous box of this warp has been read out of shared memory
    __syncwarp();
#pragma unroll
    for (int ch = 0; ch < 4; ++ch)
      st_shared_v4(c.stage_s + stage_off64(c.lane, ch), __float_as_uint(o[ch].x), __float_as_uint(o[ch].y), __float_as_uint(o[ch].z),
                   __float_as_uint(o[ch].w));
    fence_proxy_async();  // generic-proxy writes of every lane -> visible to the async proxy (TMA)
    __syncwarp();
    if (c.lane == 0) {
      tma_reduce_add_2d_s(c.o0, c.stage_s, col, c.row0);
      bulk_commit();
    }
  }
  static __device__ __forceinline__ void finish(const WsCtx& c) {
    if (c.lane == 0) bulk_wait_read<0>();  // staged boxes must be read out before the CTA's shared memory goes away
  }
};

// ------------------------------------------------------------------------------------------------ input embedding epilogue
// h[row, n] = acc + (bx + bc)[n] + E_mask[mask[row]][n]   (latent_si_v31.py:172), fp32, staged row-contiguous stores.
// The GEMM operands are the bf16 hi/lo split of [x | x_cond] and of [Wx | Wc] (elementwise.cuh: split3_embed_kernel),
// so the product carries ~16 mantissa bits.
struct EpiEmbedWs {
  struct Params {
    const float* bias;        // [H] = x_in.bias + cond_to_emb.bias
    const float* emask;       // [2, H] mask_to_emb.weight
    const long long* mask;    // [rows] int64 in {0, 1}
    float* h;                 // [rows, H]
    int H, rows;
  };
  static constexpr int CW = 16;
  static constexpr int kStageBytes = 2560;
  static constexpr int PITCH = 80;  // 16 fp32 = 64 B per staged row, padded to an odd number of 16-byte units
  struct Tile {
    uint32_t em_s;  // shared address of this row's mask embedding
  };
  static __host__ __device__ int smem_floats(const Params& p) { return 3 * p.H; }
  static __device__ void load_consts(const Params& p, float* smf, int tid, int nthreads) {
    for (int i = tid; i < p.H; i += nthreads) smf[i] = p.bias[i];
    for (int i = tid; i < 2 * p.H; i += nthreads) smf[p.H + i] = p.emask[i];
  }
  template <int BN>
  static __device__ __forceinline__ int tile_n0(const Params&, int nt) { return nt * BN; }
  static __device__ __forceinline__ void tile_begin(const Params& p, const WsCtx& c, Tile& t, int row, int, int) {
    const int m = row < p.rows ? (p.mask[row] != 0 ? 1 : 0) : 0;
    t.em_s = c.smf_s + (p.H + m * p.H) * 4;
  }
  static __device__ __forceinline__ void chunk(const Params& p, const WsCtx& c, const Tile& t, const uint32_t* v, int col, int) {
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 bv = ld_shared_f4(c.smf_s + col * 4 + j * 16), ev = ld_shared_f4(t.em_s + col * 4 + j * 16);
      st_shared_v4(c.stage_s + c.lane * PITCH + j * 16, __float_as_uint(__uint_as_float(v[4 * j + 0]) + bv.x + ev.x),
                   __float_as_uint(__uint_as_float(v[4 * j + 1]) + bv.y + ev.y), __float_as_uint(__uint_as_float(v[4 * j + 2]) + bv.z + ev.z),
                   __float_as_uint(__uint_as_float(v[4 * j + 3]) + bv.w + ev.w));
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int id = c.lane + 32 * k;
      const int r = id >> 2, ch = id & 3;
      const uint4 val = ld_shared_v4(c.stage_s + r * PITCH + ch * 16);
      if (c.row0 + r < p.rows) *reinterpret_cast<uint4*>(p.h + (size_t)(c.row0 + r) * p.H + col + ch * 4) = val;
    }
  }
  static __device__ __forceinline__ void finish(const WsCtx&) {}
};

// drains the accumulator and stores nothing: isolates the TMA + MMA main loop (lamslide_debug_gemm_mainloop)
struct EpiNullWs {
  struct Params {
    int dummy;
  };
  static constexpr int CW = 16;
  static constexpr int kStageBytes = 0;
  struct Tile {};
  static __host__ __device__ int smem_floats(const Params&) { return 0; }
  static __device__ void load_consts(const Params&, float*, int, int) {}
  template <int BN>
  static __device__ __forceinline__ int tile_n0(const Params&, int nt) { return nt * BN; }
  static __device__ __forceinline__ void tile_begin(const Params&, const WsCtx&, Tile&, int, int, int) {}
  static __device__ __forceinline__ void chunk(const Params&, const WsCtx&, const Tile&, const uint32_t*, int, int) {}
  static __device__ __forceinline__ void finish(const WsCtx&) {}
};

// ------------------------------------------------------------------------------------------------ kernel
struct WsSmemPlan {
  int a_res_bytes, stage_bytes, ring_bytes, staging_bytes, const_bytes, bar_bytes, total;
};
// cl = 1: one CTA per tile row block; cl = 2: CTA pair (each CTA stages its own A rows and HALF of every weight tile)
static inline __host__ __device__ WsSmemPlan ws_smem_plan(int BN, int num_k_blocks, int stages, int a_resident, int const_floats,
                                                          int stage_bytes_per_warp, int cl) {
  WsSmemPlan s;
  constexpr int kABytes = kBlockM * kBlockK * 2;
  s.a_res_bytes = a_resident ? num_k_blocks * kABytes : 0;
  s.stage_bytes = (a_resident ? 0 : kABytes) + (BN / cl) * kBlockK * 2;
  s.ring_bytes = stages * s.stage_bytes;
  s.staging_bytes = kWsEpiWarps * stage_bytes_per_warp;
  s.const_bytes = (const_floats * 4 + 15) / 16 * 16;
  s.bar_bytes = 320;  // 37 mbarriers + the TMEM base slot
  s.total = s.a_res_bytes + s.ring_bytes + s.staging_bytes + s.const_bytes + s.bar_bytes;  // dynamic smem base is 1024-aligned (checked)
  return s;
}

// CL = 1: every CTA is on its own (tcgen05 cta_group::1, M = 128).
// CL = 2: CTA pairs (2-CTA clusters, tcgen05 cta_group::2, M = 256): the two CTAs own neighbouring m-blocks and walk the n-tiles
// together; each stages its own 128 A rows and HALF of every weight tile, the leader's MMA thread issues one 256-row MMA per
// k-step that reads both halves.  Per SM this halves the weight bytes that cross the L2 -> SM port, land in shared memory and are
// fetched by the tensor core — the one-CTA version tops out at ~1.1 PFLOP/s (operand-feed bound, not L2- or latency-bound:
// deeper rings and weight-tile multicast left it unchanged).  Protocol (same barrier offsets in both CTAs):
//   full[s], a_full[kb]  : on the leader only; both producers' TMA loads complete_tx there (leader arms 2x the bytes)
//   empty[s], a_empty[kb]: per CTA; the leader's MMA thread commits to both CTAs
//   tmem_full[acc]       : per CTA; the leader's MMA thread commits to both CTAs
//   tmem_empty[acc]      : on the leader; the epilogue warps of BOTH CTAs arrive (the peer's through its cluster address)
template <int BN, int CL, class Epi>
__global__ void __launch_bounds__(kWsThreads, 1)  // 18 warps are allocated as 20 (granularity 4): <= 96 registers per thread
gemm_ws_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_o0, const __grid_constant__ CUtensorMap tmap_o1, int num_m_blocks,
               int num_n_tiles, int num_k_blocks, int stages, int a_resident, typename Epi::Params ep) {
  static_assert(CL == 1 || CL == 2, "cluster size 1 or 2");
  static_assert(BN % 64 == 0 && BN >= 64 && BN <= 256, "two accumulator stages of BN columns must fit 512 TMEM columns");
  constexpr int kABytes = kBlockM * kBlockK * 2;
  constexpr uint32_t kTmemCols = tmem_cols_for(2 * BN);
  constexpr int QW = BN / 4;
  constexpr int CW = Epi::CW;
  static_assert(QW % CW == 0, "a warp's column quarter must hold whole epilogue chunks");
  constexpr int kBRows = BN / CL;  // weight-tile rows staged by this CTA
  static_assert(kBRows % 8 == 0, "whole 8-row swizzle atoms");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // the 128-byte-swizzled operand tiles need a 1024-byte aligned base
  const WsSmemPlan plan = ws_smem_plan(BN, num_k_blocks, stages, a_resident, Epi::smem_floats(ep), Epi::kStageBytes, CL);
  uint8_t* a_res = smem;
  uint8_t* ring = a_res + plan.a_res_bytes;
  uint8_t* staging = ring + plan.ring_bytes;
  float* smf = reinterpret_cast<float*>(staging + plan.staging_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(smf) + plan.const_bytes);
  uint64_t* full_bar = bars;                                   // [stages]  (stages <= 8)
  uint64_t* empty_bar = bars + 8;                              // [stages]
  uint64_t* a_full = bars + 16;                                // [num_k_blocks] (<= 8)
  uint64_t* a_empty = bars + 24;
  uint64_t* tmem_full = bars + 32;                             // [2]
  uint64_t* tmem_empty = bars + 34;                            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 36);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform: TMA / MMA operands stay in uniform registers
  const int lane = threadIdx.x & 31;
  const int cta_rank = CL > 1 ? (int)cluster_ctarank() : 0;
  const bool leader = cta_rank == 0;
  const int m_first = (blockIdx.x / CL) * CL;          // first m-block of this CTA's cluster
  const int m_step = (gridDim.x / CL) * CL;            // m-blocks per sweep of the whole grid

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int k = 0; k < kWsMaxKBlocksResident; ++k) {
      mbar_init(&a_full[k], 1);
      mbar_init(&a_empty[k], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], CL * kWsEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (CL == 1) tmem_alloc<kTmemCols>(tmem_slot);
    else tmem_alloc_pair<kTmemCols>(tmem_slot);
  }
  if (warp >= 2) Epi::load_consts(ep, smf, threadIdx.x - 64, kWsThreads - 64);
  tcgen05_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();  // the peer's barriers are initialised before any TMA completion / remote arrive
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // The two control warps run their loops with ALL 32 lanes (warp-uniform control flow and addresses, so descriptors and
  // barrier addresses live in uniform registers) and elect one lane only around the TMA / tcgen05 instructions themselves.
  // (Putting the whole loop under `if (lane == 0)` made the compiler wrap every UTCHMMA / UTMALDG in a lane-serialising
  // R2UR loop: ~130 instructions per k-block on the MMA thread, which capped the tensor pipe at ~50 %.)
  if (warp == 0) {
    // ===== TMA producer =====
    int s = 0;
    uint32_t ph = 0, mb_iter = 0;
    for (int mbase = m_first; mbase < num_m_blocks; mbase += m_step, ++mb_iter) {
      const int mb = mbase + cta_rank;  // may be past the end in the last sweep: the TMA zero-fills, nothing is stored
      for (int nt = 0; nt < num_n_tiles; ++nt) {
        const int n0 = Epi::template tile_n0<BN>(ep, nt) + cta_rank * kBRows;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          if (a_resident && nt == 0) {  // refill A k-block kb as soon as the previous m-block's last tile has consumed it
            mbar_wait(&a_empty[kb], (mb_iter & 1) ^ 1);
            if (elect_one()) {
              if constexpr (CL == 1) {
                mbar_arrive_expect_tx(&a_full[kb], kABytes);
                tma_load_2d(&tmap_a, &a_full[kb], a_res + kb * kABytes, kb * kBlockK, mb * kBlockM);
              } else {
                if (leader) mbar_arrive_expect_tx(&a_full[kb], 2 * kABytes);
                tma_load_2d_pair(&tmap_a, mapa_u32(smem_u32(&a_full[kb]), 0), a_res + kb * kABytes, kb * kBlockK, mb * kBlockM);
              }
            }
            __syncwarp();
          }
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (elect_one()) {
            uint8_t* dst = ring + s * plan.stage_bytes;
            if constexpr (CL == 1) {
              mbar_arrive_expect_tx(&full_bar[s], plan.stage_bytes);
              if (!a_resident) {
                tma_load_2d(&tmap_a, &full_bar[s], dst, kb * kBlockK, mb * kBlockM);
                dst += kABytes;
              }
              tma_load_2d(&tmap_b, &full_bar[s], dst, kb * kBlockK, n0);
            } else {
              if (leader) mbar_arrive_expect_tx(&full_bar[s], 2 * plan.stage_bytes);
              const uint32_t bar = mapa_u32(smem_u32(&full_bar[s]), 0);
              if (!a_resident) {
                tma_load_2d_pair(&tmap_a, bar, dst, kb * kBlockK, mb * kBlockM);
                dst += kABytes;
              }
              tma_load_2d_pair(&tmap_b, bar, dst, kb * kBlockK, n0);
            }
          }
          __syncwarp();
          if (++s == stages) s = 0, ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (in a CTA pair only the leader's warp) =====
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBlockM * CL, BN);
      int s = 0;
      uint32_t ph = 0, mb_iter = 0, tile = 0;
      for (int mbase = m_first; mbase < num_m_blocks; mbase += m_step, ++mb_iter) {
        for (int nt = 0; nt < num_n_tiles; ++nt, ++tile) {
          const uint32_t acc = tile & 1, use = tile >> 1;
          mbar_wait(&tmem_empty[acc], (use & 1) ^ 1);  // the epilogue (of both CTAs) has drained this accumulator stage
          tcgen05_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BN;
          for (int kb = 0; kb < num_k_blocks; ++kb) {
            if (a_resident && nt == 0) mbar_wait(&a_full[kb], mb_iter & 1);
            mbar_wait(&full_bar[s], ph);
            tcgen05_fence_after();
            const uint32_t st_addr = smem_u32(ring + s * plan.stage_bytes);
            const uint32_t a_addr = a_resident ? smem_u32(a_res + kb * kABytes) : st_addr;
            const uint32_t b_addr = a_resident ? st_addr : st_addr + kABytes;
            const uint64_t a_desc = umma_desc_sw128(a_addr);
            const uint64_t b_desc = umma_desc_sw128(b_addr);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                if constexpr (CL == 1) umma_bf16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
                else umma_bf16_ss_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
              }
              if constexpr (CL == 1) {
                umma_commit(&empty_bar[s]);
                if (a_resident && nt == num_n_tiles - 1) umma_commit(&a_empty[kb]);
              } else {
                umma_commit_pair(&empty_bar[s]);
                if (a_resident && nt == num_n_tiles - 1) umma_commit_pair(&a_empty[kb]);
              }
              if (kb == num_k_blocks - 1) {
                if constexpr (CL == 1) umma_commit(&tmem_full[acc]);
                else umma_commit_pair(&tmem_full[acc]);
              }
            }
            __syncwarp();
            if (++s == stages) s = 0, ph ^= 1;
          }
        }
      }
    }
  } else {
    // ===== epilogue: 16 warps, warp = (TMEM lane quarter, column quarter) =====
    const int q = warp & 3;
    const int cq = (warp - 2) >> 2;
    WsCtx c;
    c.o0 = &tmap_o0, c.o1 = &tmap_o1;
    c.stage_s = smem_u32(staging + (warp - 2) * Epi::kStageBytes);
    c.smf_s = smem_u32(smf);
    c.lane = lane;
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + cq * QW;
    // where this warp reports "accumulator drained": the leader's barriers (stage acc at + 8 * acc)
    const uint32_t tmem_empty_addr = CL > 1 ? mapa_u32(smem_u32(&tmem_empty[0]), 0) : smem_u32(&tmem_empty[0]);
    uint32_t tile = 0;
    for (int mbase = m_first; mbase < num_m_blocks; mbase += m_step) {
      const int mb = mbase + cta_rank;
      c.row0 = mb * kBlockM + q * 32;
      const int row = c.row0 + lane;
      for (int nt = 0; nt < num_n_tiles; ++nt, ++tile) {
        const uint32_t acc = tile & 1, use = tile >> 1;
        const int n0w = Epi::template tile_n0<BN>(ep, nt) + cq * QW;
        typename Epi::Tile ts;
        Epi::tile_begin(ep, c, ts, row, n0w, QW);
        mbar_wait(&tmem_full[acc], use & 1);
        tcgen05_fence_after();
#pragma unroll 1
        for (int ck = 0; ck < QW / CW; ++ck) {
          uint32_t v[CW];
          tmem_ld<CW>(lane_taddr + acc * BN + ck * CW, v);
          tmem_ld_wait();
          if (ck == QW / CW - 1) {  // last chunk is in registers: hand the accumulator stage back to the MMA warp
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (CL == 1) mbar_arrive(&tmem_empty[acc]);
              else mbar_arrive_cluster(tmem_empty_addr + 8 * acc);
            }
          }
          Epi::chunk(ep, c, ts, v, n0w + ck * CW, ck);
        }
      }
    }
    Epi::finish(c);
  }

  tcgen05_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();  // no CTA leaves (or frees TMEM) while the pair still works
  if (warp == 1) {
    tcgen05_fence_after();
    if constexpr (CL == 1) tmem_dealloc<kTmemCols>(tmem_base);
    else tmem_dealloc_pair<kTmemCols>(tmem_base);
  }
}

}  // namespace lam

// tcgen05 / TMEM / TMA GEMM for the second-stage transformer:  D[M,N] = A[M,K] · B[N,K]^T  (bf16 x bf16 -> fp32)
//
//   A : activations, row-major [rows, K] bf16 (K contiguous)  -> TMA tile 128 x 64, 128B swizzle
//   B : nn.Linear weight [N, K] bf16 (K contiguous)            -> TMA tile BN  x 64, 128B swizzle
//   D : 128 x BN fp32 accumulator in TMEM (lane = row, column = n), read back with tcgen05.ld by 4 epilogue warps,
//       one thread per output row, so every per-row epilogue of the reference block (bias, QK-RMSNorm over a head,
//       RoPE pairs, exact-erf GELU, gate * x + residual) is thread-local: no shuffles, no shared memory.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue (TMEM lane quarter = warp_id % 4).  One output tile per CTA; 2 CTAs co-reside per SM so one
// CTA's epilogue overlaps the other's main loop.  grid = (N / BN, ceil(rows / 128)), N fastest so the CTAs that
// share an A tile run together and A is fetched from HBM once.
#pragma once
#include <cuda.h>

#include "ptx.cuh"

namespace lam {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 bf16 = 128 B = one swizzle atom row
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 192;

__host__ __device__ constexpr uint32_t tmem_cols_for(int bn) { return bn <= 32 ? 32 : bn <= 64 ? 64 : bn <= 128 ? 128 : bn <= 256 ? 256 : 512; }

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = BN * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarrierBytes = 256;
  static constexpr int kTotal = STAGES * kStageBytes + kBarrierBytes + 1024;  // +1024: manual 1024B alignment
};

// ------------------------------------------------------------------------------------------------ epilogues
// Every epilogue exposes:  struct Params;  static void run(const Params&, tmem_row_addr, row, n0, smem_f32)
// `row` is the global output row of this thread (may be >= rows: then nothing is stored),
// `taddr` the TMEM address of (this thread's lane, column 0 of the tile).

// plain: out[row, n] = acc + bias[n]   (fp32; final `linear` of LatentSIV3 — latent_si_v31.py:187 — and the GEMM unit test)
struct EpiPlain {
  struct Params {
    float* out;         // [rows, ldo]
    const float* bias;  // [N] or nullptr
    int ldo;
    int rows;
  };
  template <int BN>
  static __device__ __forceinline__ void run(const Params& p, uint32_t taddr, int row, int n0) {
#pragma unroll 1
    for (int c = 0; c < BN; c += 16) {
      uint32_t r[16];
      tmem_ld<16>(taddr + c, r);
      tmem_ld_wait();
      if (row < p.rows) {
        float* o = p.out + (size_t)row * p.ldo + n0 + c;
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 v;
          v.x = __uint_as_float(r[j + 0]) + (p.bias ? __ldg(p.bias + n0 + c + j + 0) : 0.f);
          v.y = __uint_as_float(r[j + 1]) + (p.bias ? __ldg(p.bias + n0 + c + j + 1) : 0.f);
          v.z = __uint_as_float(r[j + 2]) + (p.bias ? __ldg(p.bias + n0 + c + j + 2) : 0.f);
          v.w = __uint_as_float(r[j + 3]) + (p.bias ? __ldg(p.bias + n0 + c + j + 3) : 0.f);
          *reinterpret_cast<float4*>(o + j) = v;
        }
      }
    }
  }
};

// linear1 of ParallelMLPAttentionV2 (mmdit.py:241-247): columns [0,3H) are (K=3, heads, hd) q|k|v, columns [3H,3H+M) the MLP.
//   q,k : + bias -> RMSNorm over the head (fp32, eps 1e-6, * scale; mmdit.py:132-136) -> RoPE on interleaved pairs
//         (mmdit.py:85-90) -> q additionally * (hd^-0.5 * log2 e) so the attention kernel can use exp2 -> bf16 -> qkv buffer
//   v   : + bias -> bf16 -> qkv buffer
//   mlp : + bias -> exact-erf GELU (mmdit.py:11-18) -> bf16 -> act[:, H + j]   (the A operand of linear2)
// BN divides H, so a tile is entirely q, k, v or mlp (block-uniform branch).
template <int HD>
struct EpiLinear1 {
  struct Params {
    const float* bias;      // [3H + M]
    const float* q_scale;   // [HD] query_norm.scale
    const float* k_scale;   // [HD] key_norm.scale
    const float* rope_cos;  // [S, HD/2]
    const float* rope_sin;  // [S, HD/2]
    __nv_bfloat16* qkv;     // [rows, 3H]
    __nv_bfloat16* act;     // [rows, H + M]
    int H, M, rows;
    int pos_div, pos_mod;   // rope position of a row = (row / pos_div) % pos_mod   (spatial: 1, L; temporal: L, T)
    float q_premul;         // hd^-0.5 * log2(e)
  };
  template <int BN>
  static __device__ __forceinline__ void run(const Params& p, uint32_t taddr, int row, int n0) {
    static_assert(BN % HD == 0, "a tile must hold whole heads");
    const bool live = row < p.rows;
    const int H3 = 3 * p.H;
    if (n0 >= H3) {  // ---- MLP half: GELU
      __nv_bfloat16* o = p.act + (size_t)row * (p.H + p.M) + p.H + (n0 - H3);
#pragma unroll 1
      for (int c = 0; c < BN; c += 16) {
        uint32_t r[16];
        tmem_ld<16>(taddr + c, r);
        tmem_ld_wait();
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          float a = gelu_erf(__uint_as_float(r[j]) + __ldg(p.bias + n0 + c + j));
          float b = gelu_erf(__uint_as_float(r[j + 1]) + __ldg(p.bias + n0 + c + j + 1));
          w[j >> 1] = pack_bf16x2(a, b);
        }
        if (live) {
          *reinterpret_cast<uint4*>(o + c) = make_uint4(w[0], w[1], w[2], w[3]);
          *reinterpret_cast<uint4*>(o + c + 8) = make_uint4(w[4], w[5], w[6], w[7]);
        }
      }
      return;
    }
    const int which = n0 / p.H;  // 0 q, 1 k, 2 v
    __nv_bfloat16* o = p.qkv + (size_t)row * H3 + n0;
    if (which == 2) {  // ---- V: bias only
#pragma unroll 1
      for (int c = 0; c < BN; c += 16) {
        uint32_t r[16];
        tmem_ld<16>(taddr + c, r);
        tmem_ld_wait();
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 16; j += 2)
          w[j >> 1] = pack_bf16x2(__uint_as_float(r[j]) + __ldg(p.bias + n0 + c + j),
                                  __uint_as_float(r[j + 1]) + __ldg(p.bias + n0 + c + j + 1));
        if (live) {
          *reinterpret_cast<uint4*>(o + c) = make_uint4(w[0], w[1], w[2], w[3]);
          *reinterpret_cast<uint4*>(o + c + 8) = make_uint4(w[4], w[5], w[6], w[7]);
        }
      }
      return;
    }
    // ---- Q or K: RMSNorm + RoPE per head
    const float* gamma = which == 0 ? p.q_scale : p.k_scale;
    const float post = which == 0 ? p.q_premul : 1.0f;
    const int pos = live ? (row / p.pos_div) % p.pos_mod : 0;
    float cs[HD / 2], sn[HD / 2];
#pragma unroll
    for (int i = 0; i < HD / 2; i += 4) {
      float4 c4 = __ldg(reinterpret_cast<const float4*>(p.rope_cos + (size_t)pos * (HD / 2) + i));
      float4 s4 = __ldg(reinterpret_cast<const float4*>(p.rope_sin + (size_t)pos * (HD / 2) + i));
      cs[i] = c4.x, cs[i + 1] = c4.y, cs[i + 2] = c4.z, cs[i + 3] = c4.w;
      sn[i] = s4.x, sn[i + 1] = s4.y, sn[i + 2] = s4.z, sn[i + 3] = s4.w;
    }
#pragma unroll 1
    for (int c = 0; c < BN; c += HD) {
      uint32_t r[HD];
      tmem_ld<HD>(taddr + c, r);
      tmem_ld_wait();
      float x[HD];
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < HD; ++j) {
        x[j] = __uint_as_float(r[j]) + __ldg(p.bias + n0 + c + j);
        ss = fmaf(x[j], x[j], ss);
      }
      const float rstd = rsqrtf(ss * (1.0f / HD) + 1e-6f);
      uint32_t w[HD / 2];
#pragma unroll
      for (int j = 0; j < HD; j += 2) {
        float e = x[j] * rstd * __ldg(gamma + j);
        float d = x[j + 1] * rstd * __ldg(gamma + j + 1);
        float oe = (cs[j >> 1] * e - sn[j >> 1] * d) * post;
        float od = (sn[j >> 1] * e + cs[j >> 1] * d) * post;
        w[j >> 1] = pack_bf16x2(oe, od);
      }
      if (live) {
#pragma unroll
        for (int j = 0; j < HD / 2; j += 4) *reinterpret_cast<uint4*>(o + c + 2 * j) = make_uint4(w[j], w[j + 1], w[j + 2], w[j + 3]);
      }
    }
  }
};

// linear2 + gated residual (mmdit.py:248, latent_si_v31.py:54,61):  h[row, n] += gate[b(row), n] * (acc + bias[n])
struct EpiLinear2 {
  struct Params {
    float* h;           // [rows, H] fp32 residual stream (updated in place)
    const float* bias;  // [H]
    const float* gate;  // gate of sample b at gate + b * gate_stride, [H]
    int gate_stride;
    int rows_per_sample;  // T * L
    int H, rows;
  };
  template <int BN>
  static __device__ __forceinline__ void run(const Params& p, uint32_t taddr, int row, int n0) {
    const bool live = row < p.rows;
    const int b = live ? row / p.rows_per_sample : 0;
    const float* g = p.gate + (size_t)b * p.gate_stride + n0;
    float* hp = p.h + (size_t)row * p.H + n0;
#pragma unroll 1
    for (int c = 0; c < BN; c += 16) {
      uint32_t r[16];
      tmem_ld<16>(taddr + c, r);
      tmem_ld_wait();
      if (live) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 hv = *reinterpret_cast<float4*>(hp + c + j);
          float4 gv = __ldg(reinterpret_cast<const float4*>(g + c + j));
          float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c + j));
          hv.x = fmaf(gv.x, __uint_as_float(r[j + 0]) + bv.x, hv.x);
          hv.y = fmaf(gv.y, __uint_as_float(r[j + 1]) + bv.y, hv.y);
          hv.z = fmaf(gv.z, __uint_as_float(r[j + 2]) + bv.z, hv.z);
          hv.w = fmaf(gv.w, __uint_as_float(r[j + 3]) + bv.w, hv.w);
          *reinterpret_cast<float4*>(hp + c + j) = hv;
        }
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------ kernel
template <int BN, int STAGES, class Epi>
__global__ void __launch_bounds__(kGemmThreads, 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, int num_k_blocks,
               typename Epi::Params ep) {
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA M=128 needs N % 16 == 0, 16 <= N <= 256");
  using SM = GemmSmem<BN, STAGES>;
  constexpr uint32_t kTmemCols = tmem_cols_for(BN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * SM::kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * kBlockM;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* a_dst = smem + s * SM::kStageBytes;
        uint8_t* b_dst = a_dst + SM::kABytes;
        mbar_arrive_expect_tx(&full_bar[s], SM::kStageBytes);
        tma_load_2d(&tmap_a, &full_bar[s], a_dst, kb * kBlockK, m0);
        tma_load_2d(&tmap_b, &full_bar[s], b_dst, kb * kBlockK, n0);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BN);
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tcgen05_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * SM::kStageBytes);
        const uint32_t b_addr = a_addr + SM::kABytes;
        const uint64_t a_desc = umma_desc_sw128(a_addr);
        const uint64_t b_desc = umma_desc_sw128(b_addr);
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
          // advance 16 bf16 = 32 B inside the 128 B swizzle atom: +2 in the (addr >> 4) field
          umma_bf16_ss(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
        }
        umma_commit(&empty_bar[s]);  // frees the smem slot when these MMAs have read it
      }
      umma_commit(tmem_full_bar);  // accumulator complete
    }
  } else {
    // ===== epilogue: 4 warps, one thread per accumulator row =====
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = m0 + q * 32 + lane;
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    Epi::template run<BN>(ep, taddr, row, n0);
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

}  // namespace lam

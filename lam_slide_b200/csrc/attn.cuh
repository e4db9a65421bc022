// Attention kernels of the second-stage transformer (mmdit.py:42-55: softmax(q k^T / sqrt(hd)) v, no mask).
// q and k arrive already RMS-normalised + rotated, and q pre-multiplied by hd^-0.5 * log2(e), from the linear1 GEMM
// epilogue, so both kernels work in the exp2 domain.
//
//   attn_flash_kernel : flash-style streaming softmax over long sequences (temporal axis, S = T up to 1000; spatial axis
//                       when L is large, e.g. MD17 L = 192).  128 queries per CTA (8 warps x 16 rows), 64-key K/V tiles
//                       double-buffered in shared memory with cp.async, QK^T and PV on warp-level bf16 MMA
//                       (m16n8k16, fp32 accumulate), softmax statistics in fp32 registers.  The kernel is bound by
//                       MUFU.EX2 (S*S exps per head), not by the tensor pipe: hd = 16/24/32 gives only ~100 tensor
//                       FLOPs per exp.
//   attn_small_kernel : S <= 32 (4AA / pedestrian L = 2, NBA L = 8): one thread per (token, head).
//
// Sequence addressing (token-major qkv [tokens, 3H], (K=3, heads, hd) feature order):
//   sequence z -> base token = (z / inner) * outer_stride + (z % inner) * inner_stride,  token(s) = base + s * seq_stride
//   temporal: inner = L, outer_stride = T*L, inner_stride = 1, seq_stride = L;  spatial: inner = 1, outer_stride = L, seq_stride = 1
#pragma once
#include "ptx.cuh"

namespace lam {

struct SeqMap {
  int S, inner, outer_stride, inner_stride, seq_stride;
  __device__ __forceinline__ long long base(int z) const {
    return (long long)(z / inner) * outer_stride + (long long)(z % inner) * inner_stride;
  }
};

template <int HD>
__global__ void __launch_bounds__(256)
attn_flash_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int H, int ldo, SeqMap sm, int n_qtiles) {
  constexpr int HDP = (HD == 24) ? 32 : HD;  // contraction dim of QK^T padded to a multiple of 16
  constexpr int KSTEPS = HDP / 16;
  constexpr int PITCH = HDP + 8;  // bf16 elements; 80 B / 48 B row pitch: conflict-free fragment loads and ldmatrix
  constexpr int KT = 64;          // keys per tile
  constexpr int QT = 128;         // queries per CTA
  constexpr int CH = HD / 8;      // 16-byte chunks per row that carry data
  constexpr int CHP = HDP / 8;    // chunks per K row incl. zero padding

  __shared__ __align__(16) __nv_bfloat16 Ks[2][KT][PITCH];
  __shared__ __align__(16) __nv_bfloat16 Vs[2][KT][PITCH];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int z = blockIdx.x / n_qtiles;
  const int qt = blockIdx.x % n_qtiles;
  const int hh = blockIdx.y;
  const int S = sm.S;
  const long long base = sm.base(z);
  const size_t ldq = (size_t)3 * H;
  const __nv_bfloat16* qptr = qkv + hh * HD;
  const __nv_bfloat16* kptr = qkv + H + hh * HD;
  const __nv_bfloat16* vptr = qkv + 2 * H + hh * HD;

  auto load_tile = [&](int kt, int buf) {
    for (int idx = tid; idx < KT * CHP; idx += 256) {
      int r = idx / CHP, c = idx % CHP;
      int s = kt * KT + r;
      bool ok = (s < S) && (c < CH);
      size_t tok = (size_t)(base + (long long)(ok ? s : 0) * sm.seq_stride);
      cp_async16(&Ks[buf][r][c * 8], kptr + tok * ldq + (c < CH ? c : 0) * 8, ok);
    }
    for (int idx = tid; idx < KT * CH; idx += 256) {
      int r = idx / CH, c = idx % CH;
      int s = kt * KT + r;
      bool ok = s < S;
      size_t tok = (size_t)(base + (long long)(ok ? s : 0) * sm.seq_stride);
      cp_async16(&Vs[buf][r][c * 8], vptr + tok * ldq + c * 8, ok);
    }
    cp_async_commit();
  };

  const int nkt = (S + KT - 1) / KT;
  load_tile(0, 0);

  // Q fragments straight from global memory (read once)
  const int q0 = qt * QT + warp * 16 + g;  // rows q0 and q0 + 8
  uint32_t aq[KSTEPS][4];
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int row = q0 + (i & 1) * 8;
      int d = ks * 16 + (i >> 1) * 8 + t * 2;
      uint32_t v = 0;
      if (row < S && d < HD) {
        size_t tok = (size_t)(base + (long long)row * sm.seq_stride);
        v = *reinterpret_cast<const uint32_t*>(qptr + tok * ldq + d);
      }
      aq[ks][i] = v;
    }
  }

  float m_i[2] = {-INFINITY, -INFINITY};
  float l_i[2] = {0.f, 0.f};
  float o[HD / 8][4];
#pragma unroll
  for (int d = 0; d < HD / 8; ++d) o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f;

  for (int kt = 0; kt < nkt; ++kt) {
    const int buf = kt & 1;
    cp_async_wait<0>();
    __syncthreads();  // tile kt visible to all warps; every warp is done with tile kt-1 (buffer buf^1)
    if (kt + 1 < nkt) load_tile(kt + 1, buf ^ 1);

    float s[KT / 8][4];
#pragma unroll
    for (int nt = 0; nt < KT / 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks) {
        uint32_t b0 = *reinterpret_cast<const uint32_t*>(&Ks[buf][nt * 8 + g][ks * 16 + t * 2]);
        uint32_t b1 = *reinterpret_cast<const uint32_t*>(&Ks[buf][nt * 8 + g][ks * 16 + 8 + t * 2]);
        mma_bf16_16816(s[nt], aq[ks], b0, b1);
      }
    }
    if (kt == nkt - 1) {
#pragma unroll
      for (int nt = 0; nt < KT / 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          int key = kt * KT + nt * 8 + t * 2 + (e & 1);
          if (key >= S) s[nt][e] = -INFINITY;
        }
      }
    }
    // online softmax (rows g: e = 0,1; rows g+8: e = 2,3)
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < KT / 8; ++nt) {
      mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float corr[2], mnew[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mnew[r] = fmaxf(m_i[r], mx[r]);
      corr[r] = fast_exp2(m_i[r] - mnew[r]);
      m_i[r] = mnew[r];
      l_i[r] *= corr[r];
    }
#pragma unroll
    for (int d = 0; d < HD / 8; ++d) {
      o[d][0] *= corr[0], o[d][1] *= corr[0];
      o[d][2] *= corr[1], o[d][3] *= corr[1];
    }
#pragma unroll
    for (int nt = 0; nt < KT / 8; ++nt) {
      s[nt][0] = fast_exp2(s[nt][0] - mnew[0]);
      s[nt][1] = fast_exp2(s[nt][1] - mnew[0]);
      s[nt][2] = fast_exp2(s[nt][2] - mnew[1]);
      s[nt][3] = fast_exp2(s[nt][3] - mnew[1]);
      l_i[0] += s[nt][0] + s[nt][1];
      l_i[1] += s[nt][2] + s[nt][3];
    }
    // O += P V
#pragma unroll
    for (int j = 0; j < KT / 16; ++j) {
      uint32_t ap[4];
      ap[0] = pack_bf16x2(s[2 * j][0], s[2 * j][1]);
      ap[1] = pack_bf16x2(s[2 * j][2], s[2 * j][3]);
      ap[2] = pack_bf16x2(s[2 * j + 1][0], s[2 * j + 1][1]);
      ap[3] = pack_bf16x2(s[2 * j + 1][2], s[2 * j + 1][3]);
#pragma unroll
      for (int d = 0; d < HD / 8; ++d) {
        uint32_t b0, b1;
        ldmatrix_x2_trans(b0, b1, smem_u32(&Vs[buf][j * 16 + (lane & 15)][d * 8]));
        mma_bf16_16816(o[d], ap, b0, b1);
      }
    }
  }

  // finalize: full row sums across the quad, normalise, store bf16
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_i[r] += __shfl_xor_sync(0xffffffffu, l_i[r], 1);
    l_i[r] += __shfl_xor_sync(0xffffffffu, l_i[r], 2);
  }
  const float inv0 = 1.f / l_i[0], inv1 = 1.f / l_i[1];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    int row = q0 + r * 8;
    if (row < S) {
      size_t tok = (size_t)(base + (long long)row * sm.seq_stride);
      __nv_bfloat16* op = out + tok * ldo + hh * HD + t * 2;
      float inv = r ? inv1 : inv0;
#pragma unroll
      for (int d = 0; d < HD / 8; ++d)
        *reinterpret_cast<uint32_t*>(op + d * 8) = pack_bf16x2(o[d][2 * r] * inv, o[d][2 * r + 1] * inv);
    }
  }
}

// ---- attn_seq_kernel: whole-sequence K/V resident in shared memory, no running maximum -------------------------------
// q and k are RMS-normalised (|q|, |k| <= sqrt(hd) * max|scale|), so every logit of the block is bounded by
//   bound = hd * max|q_scale| * max|k_scale| * hd^-0.5 * log2(e)          (computed on the host at pack time)
// and softmax can be evaluated as exp2(s) / sum exp2(s) directly in fp32 (shift invariance): no row maximum, no rescale of
// the accumulator, no cross-lane traffic in the key loop.  The host selects this kernel only when bound <= 64.
//   grid.x = n_seq * heads (head fastest: the 16 CTAs that share the 128-byte lines of a token row run together);
//   256 threads = 8 warps x 32 query rows (two m16 tiles share every K / V fragment); all of K and V of the (sequence, head)
//   is cp.async'ed once into shared memory (row pitch chosen so ldmatrix is conflict-free), then each warp streams over the
//   keys 16 at a time:  QK^T = m16n8k16 (+ m16n8k8 for the hd = 24 tail), 16 ex2 per thread, P.V = m16n8k16.  No block-level
//   barrier inside the loop; 2 CTAs per SM overlap one CTA's load phase with the other's math.
template <int HD>
struct AttnSeqCfg {
  static constexpr int PITCH = (HD == 32) ? 40 : 24;  // bf16 elements per K / V row: 48 B (hd 16, 24) or 80 B (hd 32)
  static __host__ __device__ int s16(int S) { return (S + 15) & ~15; }
  static __host__ __device__ size_t smem_bytes(int S) { return (size_t)2 * s16(S) * PITCH * 2; }
};

// POLY_MASK: bit (mt * 8 + nt * 4 + e) set => that one of the 16 exponentials a thread evaluates per 16-key block is
// computed with poly_exp2 on the FMA pipe instead of MUFU.EX2 (both pipes then run side by side).
// MT: m16 tiles (16 query rows) per warp; the CTA has 16 / MT warps, i.e. always covers 256 query rows per pass.
template <int HD, uint32_t POLY_MASK, int MT = 2>
__global__ void __launch_bounds__(512 / MT, 2)
attn_seq_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int H, int ldo, SeqMap sm, int heads) {
  using Cfg = AttnSeqCfg<HD>;
  constexpr int PITCH = Cfg::PITCH;
  constexpr int CH = HD / 8;       // 16-byte chunks per row
  constexpr int K16 = HD / 16;     // k16 steps of QK^T
  constexpr bool K8 = (HD % 16) == 8;
  constexpr int NT = HD / 8;       // n-tiles of P.V
  extern __shared__ __align__(16) uint8_t attn_smem[];
  const int S = sm.S;
  const int S16 = Cfg::s16(S);
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(attn_smem);
  __nv_bfloat16* Vs = Ks + (size_t)S16 * PITCH;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int z = blockIdx.x / heads;
  const int hh = blockIdx.x % heads;
  const long long base = sm.base(z);
  const size_t ldq = (size_t)3 * H;
  const __nv_bfloat16* qptr = qkv + hh * HD;
  const __nv_bfloat16* kptr = qkv + H + hh * HD;
  const __nv_bfloat16* vptr = qkv + 2 * H + hh * HD;

  // ---- K, V -> shared memory (rows >= S zero-filled)
  for (int idx = tid; idx < S16 * CH; idx += 512 / MT) {
    const int r = idx / CH, c = idx % CH;
    const bool ok = r < S;
    const size_t tok = (size_t)(base + (long long)(ok ? r : 0) * sm.seq_stride);
    cp_async16(Ks + (size_t)r * PITCH + c * 8, kptr + tok * ldq + c * 8, ok);
    cp_async16(Vs + (size_t)r * PITCH + c * 8, vptr + tok * ldq + c * 8, ok);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  // ldmatrix lane addressing.  K (non-transposed, B operand "col"): matrix m = lane / 8 -> keys (m / 2) * 8 + lane % 8,
  // d-chunk m % 2.  V (transposed): matrix m -> keys (m % 2) * 8 + lane % 8, d-chunk m / 2.
  const int lm = lane >> 3, lr = lane & 7;
  const uint32_t k_lane_off = (uint32_t)(((lm >> 1) * 8 + lr) * PITCH + (lm & 1) * 8) * 2;
  const uint32_t k_lane_off8 = (uint32_t)((((lm & 1)) * 8 + lr) * PITCH + 16) * 2;  // x2: keys (m & 1) * 8 + r, d 16..23
  const uint32_t v_lane_off = (uint32_t)(((lm & 1) * 8 + lr) * PITCH + (lm >> 1) * 8) * 2;
  const uint32_t ks_base = smem_u32(Ks), vs_base = smem_u32(Vs);
  const int nkb = S16 / 16;

  for (int q_base = warp * 16 * MT; q_base < S; q_base += 256) {
    // ---- Q fragments (two m16 tiles) straight from global memory
    uint32_t aq[MT][K16 > 0 ? K16 : 1][4];
    uint32_t aq8[MT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
      for (int hr = 0; hr < 2; ++hr) {
        const int row = q_base + mt * 16 + g + hr * 8;
        const bool ok = row < S;
        const __nv_bfloat16* qp = qptr + (size_t)(base + (long long)(ok ? row : 0) * sm.seq_stride) * ldq;
#pragma unroll
        for (int ks = 0; ks < K16; ++ks) {
          aq[mt][ks][hr] = ok ? *reinterpret_cast<const uint32_t*>(qp + ks * 16 + t * 2) : 0u;
          aq[mt][ks][hr + 2] = ok ? *reinterpret_cast<const uint32_t*>(qp + ks * 16 + 8 + t * 2) : 0u;
        }
        if constexpr (K8) aq8[mt][hr] = ok ? *reinterpret_cast<const uint32_t*>(qp + K16 * 16 + t * 2) : 0u;
      }
    }
    float o[MT][NT][4];
    float l[MT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      l[mt][0] = l[mt][1] = 0.f;
#pragma unroll
      for (int d = 0; d < NT; ++d) o[mt][d][0] = o[mt][d][1] = o[mt][d][2] = o[mt][d][3] = 0.f;
    }

#pragma unroll 2
    for (int kb = 0; kb < nkb; ++kb) {
      const uint32_t krow = (uint32_t)(kb * 16 * PITCH * 2);
      // K fragments: kf[ks][nt*2 + {0,1}] = (b0, b1) of n-tile nt for k16 step ks; kf8[nt] for the k8 tail
      uint32_t kf[K16 > 0 ? K16 : 1][4];
      uint32_t kf8[2];
#pragma unroll
      for (int ks = 0; ks < K16; ++ks) ldmatrix_x4(kf[ks], ks_base + krow + k_lane_off + ks * 32);
      if constexpr (K8) ldmatrix_x2(kf8, ks_base + krow + k_lane_off8 + (K16 - 1) * 32);
      float s[MT][2][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          mma_bf16_16816_z(s[mt][nt], aq[mt][0], kf[0][2 * nt], kf[0][2 * nt + 1]);
#pragma unroll
          for (int ks = 1; ks < K16; ++ks) mma_bf16_16816(s[mt][nt], aq[mt][ks], kf[ks][2 * nt], kf[ks][2 * nt + 1]);
          if constexpr (K8) mma_bf16_1688(s[mt][nt], aq8[mt], kf8[nt]);
        }
      }
      // V fragments: vf[2 * d + {0,1}] = (b0, b1) of d-tile d (k = 16 keys)
      uint32_t vf[2 * NT];
#pragma unroll
      for (int d2 = 0; d2 + 1 < NT; d2 += 2) ldmatrix_x4_trans(vf + 2 * d2, vs_base + krow + v_lane_off + d2 * 16);
      if constexpr (NT % 2 == 1) {
        uint32_t r0, r1;
        ldmatrix_x2_trans(r0, r1, vs_base + krow + (uint32_t)((lane & 15) * PITCH + (NT - 1) * 8) * 2);
        vf[2 * (NT - 1)] = r0, vf[2 * (NT - 1) + 1] = r1;
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            s[mt][nt][e] = ((POLY_MASK >> (mt * 8 + nt * 4 + e)) & 1u) ? poly_exp2(s[mt][nt][e]) : fast_exp2(s[mt][nt][e]);
        }
      }
      if (kb == nkb - 1 && S16 != S) {  // zero the probabilities of the padding keys (zero-filled K rows give exp2(0) = 1)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (kb * 16 + nt * 8 + t * 2 + (e & 1) >= S) s[mt][nt][e] = 0.f;
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        l[mt][0] += (s[mt][0][0] + s[mt][0][1]) + (s[mt][1][0] + s[mt][1][1]);
        l[mt][1] += (s[mt][0][2] + s[mt][0][3]) + (s[mt][1][2] + s[mt][1][3]);
        uint32_t ap[4];
        ap[0] = pack_bf16x2(s[mt][0][0], s[mt][0][1]);
        ap[1] = pack_bf16x2(s[mt][0][2], s[mt][0][3]);
        ap[2] = pack_bf16x2(s[mt][1][0], s[mt][1][1]);
        ap[3] = pack_bf16x2(s[mt][1][2], s[mt][1][3]);
#pragma unroll
        for (int d = 0; d < NT; ++d) mma_bf16_16816(o[mt][d], ap, vf[2 * d], vf[2 * d + 1]);
      }
    }

    // ---- finalize: row sums across the quad, normalise, store bf16
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
      for (int hr = 0; hr < 2; ++hr) {
        float lv = l[mt][hr];
        lv += __shfl_xor_sync(0xffffffffu, lv, 1);
        lv += __shfl_xor_sync(0xffffffffu, lv, 2);
        const float inv = 1.f / lv;
        const int row = q_base + mt * 16 + g + hr * 8;
        if (row < S) {
          __nv_bfloat16* op = out + (size_t)(base + (long long)row * sm.seq_stride) * ldo + hh * HD + t * 2;
#pragma unroll
          for (int d = 0; d < NT; ++d)
            *reinterpret_cast<uint32_t*>(op + d * 8) = pack_bf16x2(o[mt][d][2 * hr] * inv, o[mt][d][2 * hr + 1] * inv);
        }
      }
    }
  }
}

// One thread per (token, head); S <= 32 keys, contiguous or strided.  Online softmax in the exp2 domain.
template <int HD>
__global__ void __launch_bounds__(256)
attn_small_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int H, int ldo, int heads, SeqMap sm,
                  long long n_items) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_items) return;
  const int hh = (int)(idx % heads);
  long long rest = idx / heads;
  const int s_q = (int)(rest % sm.S);
  const int z = (int)(rest / sm.S);
  const long long base = sm.base(z);
  const size_t ldq = (size_t)3 * H;
  const size_t tok_q = (size_t)(base + (long long)s_q * sm.seq_stride);

  float q[HD], acc[HD];
  {
    const uint4* qp = reinterpret_cast<const uint4*>(qkv + tok_q * ldq + hh * HD);
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
      uint4 v = __ldg(qp + c);
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 f = __bfloat1622float2(h2[i]);
        q[c * 8 + 2 * i] = f.x, q[c * 8 + 2 * i + 1] = f.y;
      }
    }
  }
#pragma unroll
  for (int d = 0; d < HD; ++d) acc[d] = 0.f;
  float m = -INFINITY, l = 0.f;
  for (int s = 0; s < sm.S; ++s) {
    const size_t tok = (size_t)(base + (long long)s * sm.seq_stride);
    const uint4* kp = reinterpret_cast<const uint4*>(qkv + tok * ldq + H + hh * HD);
    const uint4* vp = reinterpret_cast<const uint4*>(qkv + tok * ldq + 2 * H + hh * HD);
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
      uint4 v = __ldg(kp + c);
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 f = __bfloat1622float2(h2[i]);
        dot = fmaf(q[c * 8 + 2 * i], f.x, dot);
        dot = fmaf(q[c * 8 + 2 * i + 1], f.y, dot);
      }
    }
    const float mn = fmaxf(m, dot);
    const float corr = fast_exp2(m - mn);
    const float p = fast_exp2(dot - mn);
    m = mn;
    l = l * corr + p;
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
      uint4 v = __ldg(vp + c);
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 f = __bfloat1622float2(h2[i]);
        acc[c * 8 + 2 * i] = fmaf(p, f.x, acc[c * 8 + 2 * i] * corr);
        acc[c * 8 + 2 * i + 1] = fmaf(p, f.y, acc[c * 8 + 2 * i + 1] * corr);
      }
    }
  }
  const float inv = 1.f / l;
  uint4* op = reinterpret_cast<uint4*>(out + tok_q * ldo + hh * HD);
#pragma unroll
  for (int c = 0; c < HD / 8; ++c) {
    uint4 v;
    v.x = pack_bf16x2(acc[c * 8 + 0] * inv, acc[c * 8 + 1] * inv);
    v.y = pack_bf16x2(acc[c * 8 + 2] * inv, acc[c * 8 + 3] * inv);
    v.z = pack_bf16x2(acc[c * 8 + 4] * inv, acc[c * 8 + 5] * inv);
    v.w = pack_bf16x2(acc[c * 8 + 6] * inv, acc[c * 8 + 7] * inv);
    op[c] = v;
  }
}

// Short CONTIGUOUS sequences (spatial attention: the S <= 32 tokens of a sequence are adjacent rows, e.g. the L = 2 latents of one
// frame): one warp per sequence.  The warp copies the S whole q|k|v rows (S * 3H bf16, contiguous in memory) into shared memory
// with fully coalesced 16-byte loads, every lane then serves (query, head) items out of shared memory — online softmax in the
// exp2 domain, as attn_small_kernel — writes its result over its own q slot, and the S output rows go out coalesced.
// attn_small_kernel (one thread per item, each gathering its 48-byte pieces from global memory) ran the 4AA spatial attention at
// 3.6 TB/s (109 us per launch for 393 MB); the point here is only the access pattern.
template <int HD>
__global__ void __launch_bounds__(256)
attn_rows_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int H, int ldo, int heads, int S, long long n_seq) {
  extern __shared__ __align__(16) uint8_t attn_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long seq = (long long)blockIdx.x * 8 + warp;
  if (seq >= n_seq) return;
  const int row_u4 = 3 * H / 8;  // 16-byte units per token row
  uint4* my = reinterpret_cast<uint4*>(attn_smem) + (size_t)warp * S * row_u4;
  const uint4* src = reinterpret_cast<const uint4*>(qkv + (size_t)seq * S * 3 * H);
  for (int i = lane; i < S * row_u4; i += 32) my[i] = __ldg(src + i);
  __syncwarp();
  constexpr int CH = HD / 8;
  for (int item = lane; item < S * heads; item += 32) {
    const int sq = item / heads, hh = item % heads;
    float q[HD], acc[HD];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const uint4 v = my[sq * row_u4 + hh * CH + c];
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h2[i]);
        q[c * 8 + 2 * i] = f.x, q[c * 8 + 2 * i + 1] = f.y;
      }
    }
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] = 0.f;
    float m = -INFINITY, l = 0.f;
    for (int sk = 0; sk < S; ++sk) {
      const uint4* kp = my + sk * row_u4 + (H / 8) + hh * CH;
      const uint4* vp = my + sk * row_u4 + 2 * (H / 8) + hh * CH;
      float dot = 0.f;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const uint4 v = kp[c];
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __bfloat1622float2(h2[i]);
          dot = fmaf(q[c * 8 + 2 * i], f.x, dot);
          dot = fmaf(q[c * 8 + 2 * i + 1], f.y, dot);
        }
      }
      const float mn = fmaxf(m, dot);
      const float corr = fast_exp2(m - mn);
      const float pr = fast_exp2(dot - mn);
      m = mn;
      l = l * corr + pr;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const uint4 v = vp[c];
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __bfloat1622float2(h2[i]);
          acc[c * 8 + 2 * i] = fmaf(pr, f.x, acc[c * 8 + 2 * i] * corr);
          acc[c * 8 + 2 * i + 1] = fmaf(pr, f.y, acc[c * 8 + 2 * i + 1] * corr);
        }
      }
    }
    const float inv = 1.f / l;
#pragma unroll
    for (int c = 0; c < CH; ++c) {  // over this item's own q slot (no other lane reads it)
      uint4 v;
      v.x = pack_bf16x2(acc[c * 8 + 0] * inv, acc[c * 8 + 1] * inv);
      v.y = pack_bf16x2(acc[c * 8 + 2] * inv, acc[c * 8 + 3] * inv);
      v.z = pack_bf16x2(acc[c * 8 + 4] * inv, acc[c * 8 + 5] * inv);
      v.w = pack_bf16x2(acc[c * 8 + 6] * inv, acc[c * 8 + 7] * inv);
      my[sq * row_u4 + hh * CH + c] = v;
    }
  }
  __syncwarp();
  const int out_u4 = H / 8;
  for (int i = lane; i < S * out_u4; i += 32) {
    const int tok = i / out_u4, c = i % out_u4;
    *reinterpret_cast<uint4*>(out + ((size_t)seq * S + tok) * ldo + c * 8) = my[tok * row_u4 + c];
  }
}

// Short sequences (S <= 32) with any token stride on the warp-level tensor path: one warp per (sequence, head).  The head's q | k | v
// pieces (S rows x HD bf16 each) go to shared memory, S = Q K^T (32 x 32 x HD) and O = P V (32 x HD x 32) are 16 mma.sync.m16n8k16 for
// HD = 16, the softmax over the <= 32 keys of a row stays in the accumulator registers of a lane quad.  The output goes back through
// the warp's Q tile and leaves with 16-byte stores.  (Measured per NBA launch at B = 1024, 336 MB of traffic: thread per (query, head)
// gathering 32-byte pieces of every key from global memory 380 us; warp per (sequence, 4 heads) with FMA dot products out of shared
// memory 268 us, shared-memory-bandwidth bound.)
template <int HD>
__global__ void __launch_bounds__(256)
attn_short_mma_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int H, int ldo, int heads, SeqMap sm, long long n_work) {
  constexpr int HDP = (HD == 24) ? 32 : HD;  // contraction dim of Q K^T padded to a multiple of 16
  constexpr int KSTEPS = HDP / 16;
  constexpr int PITCH = HDP + 8;             // bf16 elements: 48 / 80 byte rows, conflict-free fragment loads and ldmatrix
  constexpr int CH = HD / 8, CHP = HDP / 8;
  extern __shared__ __align__(16) uint8_t attn_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const long long w = (long long)blockIdx.x * 8 + warp;
  if (w >= n_work) return;
  const int z = (int)(w / heads), hh = (int)(w % heads);
  const long long base = sm.base(z);
  const int S = sm.S;
  typedef __nv_bfloat16 Row[PITCH];
  Row* Qs = reinterpret_cast<Row*>(attn_smem + (size_t)warp * 3 * 32 * PITCH * 2);
  Row* Ks = Qs + 32;
  Row* Vs = Ks + 32;
  for (int idx = lane; idx < 3 * 32 * CHP; idx += 32) {
    const int part = idx / (32 * CHP), r = (idx / CHP) % 32, c = idx % CHP;
    uint4 v = make_uint4(0, 0, 0, 0);  // rows >= S and the padding of the contraction dim are zero
    if (r < S && c < CH)
      v = __ldg(reinterpret_cast<const uint4*>(qkv + (size_t)(base + (long long)r * sm.seq_stride) * 3 * H + part * H + hh * HD) + c);
    *reinterpret_cast<uint4*>(&Qs[part * 32 + r][c * 8]) = v;
  }
  __syncwarp();
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    if (mt * 16 >= S) break;
    uint32_t aq[KSTEPS][4];
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks)
#pragma unroll
      for (int i = 0; i < 4; ++i)
        aq[ks][i] = *reinterpret_cast<const uint32_t*>(&Qs[mt * 16 + g + (i & 1) * 8][ks * 16 + (i >> 1) * 8 + t * 2]);
    float s[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&Ks[nt * 8 + g][ks * 16 + t * 2]);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&Ks[nt * 8 + g][ks * 16 + 8 + t * 2]);
        mma_bf16_16816(s[nt], aq[ks], b0, b1);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (nt * 8 + t * 2 + (e & 1) >= S) s[nt][e] = -INFINITY;
    }
    // softmax over the keys (rows g: e = 0, 1; rows g + 8: e = 2, 3), in the exp2 domain (q carries hd^-0.5 * log2 e)
    float mx[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      s[nt][0] = fast_exp2(s[nt][0] - mx[0]), s[nt][1] = fast_exp2(s[nt][1] - mx[0]);
      s[nt][2] = fast_exp2(s[nt][2] - mx[1]), s[nt][3] = fast_exp2(s[nt][3] - mx[1]);
      l[0] += s[nt][0] + s[nt][1];
      l[1] += s[nt][2] + s[nt][3];
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
      l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
    }
    float o[CH][4];
#pragma unroll
    for (int d = 0; d < CH; ++d) o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (j * 16 >= S) break;
      uint32_t ap[4];
      ap[0] = pack_bf16x2(s[2 * j][0], s[2 * j][1]);
      ap[1] = pack_bf16x2(s[2 * j][2], s[2 * j][3]);
      ap[2] = pack_bf16x2(s[2 * j + 1][0], s[2 * j + 1][1]);
      ap[3] = pack_bf16x2(s[2 * j + 1][2], s[2 * j + 1][3]);
#pragma unroll
      for (int d = 0; d < CH; ++d) {
        uint32_t b0, b1;
        ldmatrix_x2_trans(b0, b1, smem_u32(&Vs[j * 16 + (lane & 15)][d * 8]));
        mma_bf16_16816(o[d], ap, b0, b1);
      }
    }
    // the tile's Q rows are in registers: its output rows take their place
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float inv = 1.f / l[r];
#pragma unroll
      for (int d = 0; d < CH; ++d)
        *reinterpret_cast<uint32_t*>(&Qs[mt * 16 + g + r * 8][d * 8 + t * 2]) = pack_bf16x2(o[d][2 * r] * inv, o[d][2 * r + 1] * inv);
    }
  }
  __syncwarp();
  for (int idx = lane; idx < S * CH; idx += 32) {
    const int r = idx / CH, c = idx % CH;
    *reinterpret_cast<uint4*>(out + (size_t)(base + (long long)r * sm.seq_stride) * ldo + hh * HD + c * 8) = *reinterpret_cast<const uint4*>(&Qs[r][c * 8]);
  }
}

}  // namespace lam

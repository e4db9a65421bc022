// fp32 kernels of the UPT-style first stage (encoder.py, decoder.py, torch_modules.py): small per-entity MLPs,
// LayerNorms, entity-ID embedding gathers and the (masked) entity cross-attention / latent self-attention.
// The first stage is ~0.3 % of the trajectory FLOPs; it is kept in fp32 on the FMA pipe so that the
// latents that condition the ODE carry no bf16 rounding.  Rows = frames x tokens-per-frame.
#pragma once
#include "elementwise.cuh"

namespace lam {

__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// ---- Y[r, n] = epi( sum_k X[r, k] W[n, k] + bias[n] )   64x64 tile, 16-deep k slices, 4x4 outputs per thread.
//   epi: optional GELU (gelu = 1), optional + rowadd[(r % rowadd_period), n], optional + res[r, n], optional SiLU of the sum (gelu = 2)
struct LinearArgs {
  const float* X; int ldx;
  const float* W;      // [N, K] (nn.Linear layout)
  const float* bias;   // [N] or null
  float* Y; int ldy;
  const float* res; int ldr;        // residual or null
  const long long* res_idx = nullptr;  // residual row of output row r = res_idx[r] (a per-entity table) instead of r
  const float* rowadd; int rowadd_period; int ldra;  // e.g. sin/cos residue-index embedding, or null
  int rows, N, K;
  int gelu;
  // fused LayerNorm of the result (tcgen05 kernel only): ln_out[r, n] = LN_group(y[r, :])[n] * ln_w[n] + ln_b[n] over groups of
  // ln_group consecutive columns (ln_group = 0: nn.LayerNorm of the row, one n-tile; ln_group = BN / 2: the tokens a row is reshaped into).
  // Y may be null when only the normalised output is needed.  ln_w / ln_b null: no affine.
  float* ln_out = nullptr; int ld_ln = 0;
  const float* ln_w = nullptr; const float* ln_b = nullptr;
  int ln_group = 0;
  float ln_eps = 1e-5f;
  int debug = 0;  // profiling aid (LAMSLIDE_L5_DEBUG in debug builds): 1 no global stores, 2 epilogue only releases the accumulator,
                  // 4 splitters only arrive, 8 no weight-tile loads
};

__global__ void __launch_bounds__(256) linear_f32_kernel(LinearArgs a) {
  __shared__ float Xs[16][64 + 4];
  __shared__ float Ws[16][64 + 4];
  const int tid = threadIdx.x;
  const int r0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int ty = tid / 16, tx = tid % 16;  // thread computes rows ty*4..+3, cols tx*4..+3
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < a.K; k0 += 16) {
    // 64 x 16 tiles of X and W, 4 elements per thread each; k fastest in global => coalesced 64 B runs
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = tid + i * 256;
      int rr = e / 16, kk = e % 16;
      int r = r0 + rr, n = n0 + rr, k = k0 + kk;
      Xs[kk][rr] = (r < a.rows && k < a.K) ? a.X[(size_t)r * a.ldx + k] : 0.f;
      Ws[kk][rr] = (n < a.N && k < a.K) ? __ldg(a.W + (size_t)n * a.K + k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float4 xv = *reinterpret_cast<const float4*>(&Xs[kk][ty * 4]);
      float4 wv = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      float xr[4] = {xv.x, xv.y, xv.z, xv.w}, wr[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xr[i], wr[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int r = r0 + ty * 4 + i;
    if (r >= a.rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= a.N) continue;
      float v = acc[i][j] + (a.bias ? a.bias[n] : 0.f);
      if (a.gelu == 1) v = gelu_exact(v);
      if (a.rowadd) v += a.rowadd[(size_t)(r % a.rowadd_period) * a.ldra + n];
      if (a.res) v += a.res[(size_t)(a.res_idx ? a.res_idx[r] : r) * a.ldr + n];
      if (a.gelu == 2) v = v / (1.0f + expf(-v));
      a.Y[(size_t)r * a.ldy + n] = v;
    }
  }
}

// Same contract, 128 x 64 tile, 8 x 4 outputs per thread, 16-byte global loads (K fastest), register-staged double buffering of the
// shared-memory k-slices.  Needs K % 4 == 0, ldx % 4 == 0 and 16-byte aligned X / W rows; the launcher falls back to the kernel
// above otherwise (e.g. the 106-wide peptide feature rows).
__global__ void __launch_bounds__(256) linear_f32_v2_kernel(LinearArgs a) {
  constexpr int BM = 128, BN = 64, BK = 16;
  __shared__ __align__(16) float Xs[2][BK][BM + 4];
  __shared__ __align__(16) float Ws[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int r0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int ty = tid >> 4, tx = tid & 15;  // rows ty*8..+7, cols tx*4..+3
  // global -> register staging: X tile 128 x 16 = 512 float4 (2 per thread), W tile 64 x 16 = 256 float4 (1 per thread)
  const int xr0 = tid >> 2, xk = (tid & 3) * 4;  // X rows xr0 and xr0 + 64
  const int wn = tid >> 2;
  float4 xa, xb, wv;
  auto gload = [&](int k0) {
    const int k = k0 + xk;
    const bool kin = k < a.K;  // K % 4 == 0: a float4 is all-in or all-out
    const int ra = r0 + xr0, rb = ra + 64, n = n0 + wn;
    xa = (kin && ra < a.rows) ? *reinterpret_cast<const float4*>(a.X + (size_t)ra * a.ldx + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    xb = (kin && rb < a.rows) ? *reinterpret_cast<const float4*>(a.X + (size_t)rb * a.ldx + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    wv = (kin && n < a.N) ? __ldg(reinterpret_cast<const float4*>(a.W + (size_t)n * a.K + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto sstore = [&](int buf) {
    Xs[buf][xk + 0][xr0] = xa.x, Xs[buf][xk + 1][xr0] = xa.y, Xs[buf][xk + 2][xr0] = xa.z, Xs[buf][xk + 3][xr0] = xa.w;
    Xs[buf][xk + 0][xr0 + 64] = xb.x, Xs[buf][xk + 1][xr0 + 64] = xb.y, Xs[buf][xk + 2][xr0 + 64] = xb.z, Xs[buf][xk + 3][xr0 + 64] = xb.w;
    Ws[buf][xk + 0][wn] = wv.x, Ws[buf][xk + 1][wn] = wv.y, Ws[buf][xk + 2][wn] = wv.z, Ws[buf][xk + 3][wn] = wv.w;
  };
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int nk = (a.K + BK - 1) / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * BK);  // in flight during the FMAs below
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 x0 = *reinterpret_cast<const float4*>(&Xs[buf][kk][ty * 8]);
      const float4 x1 = *reinterpret_cast<const float4*>(&Xs[buf][kk][ty * 8 + 4]);
      const float4 w4 = *reinterpret_cast<const float4*>(&Ws[buf][kk][tx * 4]);
      const float xr[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w}, wr[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xr[i], wr[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);  // the other buffer was last read in iteration kt - 1 (barrier at its end)
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = r0 + ty * 8 + i;
    if (r >= a.rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= a.N) continue;
      float v = acc[i][j] + (a.bias ? a.bias[n] : 0.f);
      if (a.gelu == 1) v = gelu_exact(v);
      if (a.rowadd) v += a.rowadd[(size_t)(r % a.rowadd_period) * a.ldra + n];
      if (a.res) v += a.res[(size_t)(a.res_idx ? a.res_idx[r] : r) * a.ldr + n];
      if (a.gelu == 2) v = v / (1.0f + expf(-v));
      a.Y[(size_t)r * a.ldy + n] = v;
    }
  }
}

// Same contract on the (legacy) tensor path with fp32 accuracy: 3xTF32.  Every operand is split into a TF32 head and a TF32 tail
// (x = hi + lo, both exactly representable), and x.w is accumulated as lo.hi + hi.lo + hi.hi with mma.sync.m16n8k8 (fp32
// accumulate) — the dropped lo.lo term is 2^-22 relative, i.e. the result is as good as an fp32 FMA chain (the reference's
// first stage is fp32 and its outputs are checked to 1e-4).  128 x 64 CTA tile, 8 warps of 32 x 32 (2 m16 x 4 n8 tiles),
// 16-deep k slices staged through registers into K-contiguous shared-memory rows (pitch 20 floats: conflict-free fragment
// loads).  Per k8 step a warp issues 24 MMAs for 16 k flop of useful work: ~90 TFLOP/s of fp32-accurate throughput against
// ~31 TFLOP/s for the FMA kernel above (B200).  Needs K % 4 == 0, ldx % 4 == 0, 16-byte aligned rows, N % 2 == 0, ldy % 2 == 0.
// hi = x truncated to TF32 (10 explicit mantissa bits), lo = x - hi (exact in fp32; the MMA reads only its upper 19 bits, a
// 2^-11 relative truncation of lo = 2^-21 of x).  (cvt.rna.tf32.f32 is emulated on sm_100a — ~6 instructions with its NaN / Inf
// handling, which made the first version of this kernel issue-bound: 277 instructions per k8 step for 24 MMAs.)
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32_1688(float* c, const uint32_t* a, const uint32_t* b) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__global__ void __launch_bounds__(256, 2) linear_f32_tc_kernel(LinearArgs a) {
  constexpr int BM = 128, BN = 64, BK = 16, P = BK + 4;
  __shared__ __align__(16) float Xs[2][BM][P];
  __shared__ __align__(16) float Ws[2][BN][P];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;  // warp tile origin inside the CTA tile
  const int r0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int xr0 = tid >> 2, xk = (tid & 3) * 4;  // staging: X rows xr0 and xr0 + 64, W row xr0 (tid < 256 -> 64 rows)
  float4 xa, xb, wv;
  auto gload = [&](int k0) {
    const int k = k0 + xk;
    const bool kin = k < a.K;  // K % 4 == 0: a float4 is all-in or all-out
    const int ra = r0 + xr0, rb = ra + 64, n = n0 + xr0;
    xa = (kin && ra < a.rows) ? *reinterpret_cast<const float4*>(a.X + (size_t)ra * a.ldx + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    xb = (kin && rb < a.rows) ? *reinterpret_cast<const float4*>(a.X + (size_t)rb * a.ldx + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    wv = (kin && n < a.N) ? __ldg(reinterpret_cast<const float4*>(a.W + (size_t)n * a.K + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto sstore = [&](int buf) {
    *reinterpret_cast<float4*>(&Xs[buf][xr0][xk]) = xa;
    *reinterpret_cast<float4*>(&Xs[buf][xr0 + 64][xk]) = xb;
    *reinterpret_cast<float4*>(&Ws[buf][xr0][xk]) = wv;
  };
  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;
  const int nk = (a.K + BK - 1) / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * BK);  // in flight during the MMAs below
#pragma unroll
    for (int k8 = 0; k8 < BK; k8 += 8) {
      uint32_t ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const float* xp = &Xs[buf][wm + mt * 16 + g][k8 + t];
        split_tf32(xp[0], ah[mt][0], al[mt][0]);
        split_tf32(xp[8 * P], ah[mt][1], al[mt][1]);
        split_tf32(xp[4], ah[mt][2], al[mt][2]);
        split_tf32(xp[8 * P + 4], ah[mt][3], al[mt][3]);
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const float* wp = &Ws[buf][wn + nt * 8 + g][k8 + t];
        split_tf32(wp[0], bh[nt][0], bl[nt][0]);
        split_tf32(wp[4], bh[nt][1], bl[nt][1]);
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          mma_tf32_1688(acc[mt][nt], al[mt], bh[nt]);
          mma_tf32_1688(acc[mt][nt], ah[mt], bl[nt]);
          mma_tf32_1688(acc[mt][nt], ah[mt], bh[nt]);
        }
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);  // the other buffer was last read in iteration kt - 1 (barrier at its end)
      __syncthreads();
    }
  }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      const int r = r0 + wm + mt * 16 + g + hr * 8;
      if (r >= a.rows) continue;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int n = n0 + wn + nt * 8 + t * 2;
        if (n >= a.N) continue;  // N % 2 == 0: a column pair is all-in or all-out
        float v0 = acc[mt][nt][2 * hr] + (a.bias ? a.bias[n] : 0.f);
        float v1 = acc[mt][nt][2 * hr + 1] + (a.bias ? a.bias[n + 1] : 0.f);
        if (a.gelu == 1) v0 = gelu_erf(v0), v1 = gelu_erf(v1);  // erf to 3e-7 abs in ~14 instructions (libm erff: ~40, half of this kernel at K = 128)
        if (a.rowadd) {
          const float* ra = a.rowadd + (size_t)(r % a.rowadd_period) * a.ldra + n;
          v0 += ra[0], v1 += ra[1];
        }
        if (a.res) {
          const float* rs = a.res + (size_t)(a.res_idx ? a.res_idx[r] : r) * a.ldr + n;
          v0 += rs[0], v1 += rs[1];
        }
        if (a.gelu == 2) v0 = v0 / (1.0f + expf(-v0)), v1 = v1 / (1.0f + expf(-v1));
        *reinterpret_cast<float2*>(a.Y + (size_t)r * a.ldy + n) = make_float2(v0, v1);
      }
    }
}

// ---- LayerNorm over rows (nn.LayerNorm semantics, eps given, optional affine), out-of-place with pitches; warp per row.
// x_bcast_period > 0: the input row is x[r % period] (used to normalise the learned latents once per frame without
// materialising the broadcast — encoder.py:39).
__global__ void __launch_bounds__(256)
layernorm_f32_kernel(const float* __restrict__ x, int ldx, int x_bcast_period, float* __restrict__ y, int ldy,
                     const float* __restrict__ w, const float* __restrict__ b, int rows, int dim, float eps) {
  int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* p = x + (size_t)(x_bcast_period > 0 ? row % x_bcast_period : row) * ldx;
  float s = 0.f;
  for (int j = lane; j < dim; j += 32) s += p[j];
  float mean = warp_sum(s) / dim;
  float v = 0.f;
  for (int j = lane; j < dim; j += 32) {
    float d = p[j] - mean;
    v = fmaf(d, d, v);
  }
  float rstd = rsqrtf(warp_sum(v) / dim + eps);
  float* o = y + (size_t)row * ldy;
  for (int j = lane; j < dim; j += 32) {
    float t = (p[j] - mean) * rstd;
    o[j] = w ? fmaf(t, w[j], b[j]) : t;
  }
}

// ---- dst[r, col0 + j] = table[idx[r], j]   (embedding gather into a column block of a wider row-major matrix)
__global__ void gather_cols_kernel(float* __restrict__ dst, int ldd, int col0, const float* __restrict__ table, int width,
                                   const long long* __restrict__ idx, long long rows, int pad = 0) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int wp = width + pad;  // `pad` zero columns behind the gathered ones (see copy_cols_kernel)
  if (i >= rows * wp) return;
  long long r = i / wp;
  int j = (int)(i % wp);
  dst[(size_t)r * ldd + col0 + j] = j < width ? table[(size_t)idx[r] * width + j] : 0.f;
}
// same for width, col0, ldd multiples of 4 and 16-byte aligned pointers: one float4 per thread
__global__ void gather_cols4_kernel(float* __restrict__ dst, int ldd, int col0, const float* __restrict__ table, int width,
                                    const long long* __restrict__ idx, long long rows) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int w4 = width >> 2;
  if (i >= rows * w4) return;
  long long r = i / w4;
  int j = (int)(i % w4) * 4;
  *reinterpret_cast<float4*>(dst + (size_t)r * ldd + col0 + j) = __ldg(reinterpret_cast<const float4*>(table + (size_t)idx[r] * width + j));
}
// ---- dst[r, col0 + j] = src[r, j] for j < width, 0 for width <= j < width + pad (zero columns that round a feature row up to a
// multiple of 4 floats, so the following linear layer can take the vectorised / tensor-core kernels)
__global__ void copy_cols_kernel(float* __restrict__ dst, int ldd, int col0, const float* __restrict__ src, int width, int pad, long long rows) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int wp = width + pad;
  if (i >= rows * wp) return;
  long long r = i / wp;
  int j = (int)(i % wp);
  dst[(size_t)r * ldd + col0 + j] = j < width ? src[(size_t)r * width + j] : 0.f;
}
// ---- dst[r, :] = src[r % period, :]   (broadcast of the learned latents over frames)
__global__ void bcast_rows_kernel(float* __restrict__ dst, const float* __restrict__ src, int width, int period, long long rows) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * width) return;
  long long r = i / width;
  int j = (int)(i % width);
  dst[i] = src[(size_t)(r % period) * width + j];
}
// ---- PointEmbed features (embeddings.py:80-88): [sin(pos . basis) (nb), cos(pos . basis) (nb), pos (3)], basis [3, nb]
__global__ void point_feats_kernel(const float* __restrict__ pos, const float* __restrict__ basis, float* __restrict__ out, int nb,
                                   long long rows) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int width = 2 * nb + 3;
  if (i >= rows * width) return;
  long long r = i / width;
  int j = (int)(i % width);
  const float* p = pos + (size_t)r * 3;
  float v;
  if (j >= 2 * nb) {
    v = p[j - 2 * nb];
  } else {
    int e = j < nb ? j : j - nb;
    float proj = p[0] * basis[e] + p[1] * basis[nb + e] + p[2] * basis[2 * nb + e];
    v = j < nb ? sinf(proj) : cosf(proj);
  }
  out[i] = v;
}

// ---- attention of the Perceiver-style blocks (torch_modules.py:147-186, 221-253), dim_head = 16, fp32.
// One thread per (frame, head, query).  q/k are RMS-normalised per head (eps 1e-6, learned scale) when qk_norm;
// keys with mask == 0 are excluded (bool key mask, True = keep).  q_frame_stride may be 0 (queries shared by all frames).
struct SmallAttnArgs {
  const float* q; long long q_frame_stride; int ldq;   // q[frame, sq, head*16 + d]
  const long long* q_index = nullptr;                  // when set: the query row of (frame, sq) is q[q_index[frame * Sq + sq]] (a per-entity table)
  const float* k; const float* v; long long kv_frame_stride; int ldkv;  // k/v[frame, sk, head*16 + d] (pointers pre-offset)
  const float* gq; const float* gk;                    // RMSNorm scales [16] or null
  const unsigned char* mask;                           // [frames, Sk] or null
  float* out; int ldo;                                 // out[frame, sq, head*16 + d]
  int frames, Sq, Sk, heads;
  float scale;
};

__global__ void __launch_bounds__(128) small_attn_f32_kernel(SmallAttnArgs a) {
  constexpr int DH = 16;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long tot = (long long)a.frames * a.heads * a.Sq;
  if (idx >= tot) return;
  const int sq = (int)(idx % a.Sq);
  const int hh = (int)((idx / a.Sq) % a.heads);
  const long long f = idx / ((long long)a.Sq * a.heads);
  float q[DH], acc[DH];
  {
    const float4* qp = reinterpret_cast<const float4*>(
        a.q_index ? a.q + (size_t)a.q_index[f * a.Sq + sq] * a.ldq + hh * DH : a.q + f * a.q_frame_stride + (size_t)sq * a.ldq + hh * DH);
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float4 t = qp[c];
      q[4 * c] = t.x, q[4 * c + 1] = t.y, q[4 * c + 2] = t.z, q[4 * c + 3] = t.w;
      ss += t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w;
    }
    if (a.gq) {
      float r = rsqrtf(ss * (1.0f / DH) + 1e-6f);
#pragma unroll
      for (int d = 0; d < DH; ++d) q[d] = q[d] * r * a.gq[d];
    }
#pragma unroll
    for (int d = 0; d < DH; ++d) q[d] *= a.scale;
  }
#pragma unroll
  for (int d = 0; d < DH; ++d) acc[d] = 0.f;
  float m = -INFINITY, l = 0.f;
  const float* kb = a.k + f * a.kv_frame_stride + hh * DH;
  const float* vb = a.v + f * a.kv_frame_stride + hh * DH;
  for (int s = 0; s < a.Sk; ++s) {
    if (a.mask && !a.mask[f * a.Sk + s]) continue;
    const float4* kp = reinterpret_cast<const float4*>(kb + (size_t)s * a.ldkv);
    float kk[DH];
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float4 t = kp[c];
      kk[4 * c] = t.x, kk[4 * c + 1] = t.y, kk[4 * c + 2] = t.z, kk[4 * c + 3] = t.w;
      ss += t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w;
    }
    float dot = 0.f;
    if (a.gk) {
      float r = rsqrtf(ss * (1.0f / DH) + 1e-6f);
#pragma unroll
      for (int d = 0; d < DH; ++d) dot = fmaf(q[d], kk[d] * r * a.gk[d], dot);
    } else {
#pragma unroll
      for (int d = 0; d < DH; ++d) dot = fmaf(q[d], kk[d], dot);
    }
    const float mn = fmaxf(m, dot);
    const float corr = expf(m - mn);
    const float p = expf(dot - mn);
    m = mn;
    l = l * corr + p;
    const float4* vp = reinterpret_cast<const float4*>(vb + (size_t)s * a.ldkv);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float4 t = vp[c];
      acc[4 * c] = fmaf(p, t.x, acc[4 * c] * corr);
      acc[4 * c + 1] = fmaf(p, t.y, acc[4 * c + 1] * corr);
      acc[4 * c + 2] = fmaf(p, t.z, acc[4 * c + 2] * corr);
      acc[4 * c + 3] = fmaf(p, t.w, acc[4 * c + 3] * corr);
    }
  }
  const float inv = 1.0f / l;
  float4* op = reinterpret_cast<float4*>(a.out + ((size_t)f * a.Sq + sq) * a.ldo + hh * DH);
#pragma unroll
  for (int c = 0; c < 4; ++c) op[c] = make_float4(acc[4 * c] * inv, acc[4 * c + 1] * inv, acc[4 * c + 2] * inv, acc[4 * c + 3] * inv);
}

}  // namespace lam

// Vectorised / warp-shuffle kernels around the GEMMs of the second stage: input embedding, the per-sample vector path
// (timestep embedding -> time_in [+ vec_in] -> all adaLN modulations), LayerNorm+modulate, drift + Euler update,
// conditioning, RoPE tables.  All statistics and the ODE state stay in fp32.
#pragma once
#include "ptx.cuh"

namespace lam {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }

// ---- RoPE tables (mmdit.py:75-82): angle = p * theta^(-2i/hd) in fp64, cos/sin stored fp32 [S, hd/2]
__global__ void rope_table_kernel(float* __restrict__ cs, float* __restrict__ sn, int S, int half, double theta) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= S * half) return;
  int p = idx / half, i = idx % half;
  double omega = 1.0 / pow(theta, (double)(2 * i) / (double)(2 * half));
  double a = (double)p * omega;
  cs[idx] = (float)cos(a);
  sn[idx] = (float)sin(a);
}

// ---- timestep_embedding (mmdit.py:93-113): e = cat[cos(1000 t f), sin(1000 t f)], f_i = exp(-ln(1e4) i / 128); [B, 256]
__global__ void timestep_embed_kernel(const float* __restrict__ t, float* __restrict__ e, int B) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * 128) return;
  int b = idx / 128, i = idx % 128;
  float f = expf(-9.210340371976184f * (float)i / 128.0f);
  float arg = (1000.0f * t[b]) * f;
  e[b * 256 + i] = cosf(arg);
  e[b * 256 + 128 + i] = sinf(arg);
}

// ---- input embedding (latent_si_v31.py:172-174): h = x Wx^T + x_cond Wc^T + (bx + bc) + E_mask[m]
// Wt is the pre-transposed, concatenated weight [2D, H] (k-major so a warp reads consecutive output columns).
// 32 tokens per block; each thread owns output columns tid and tid + 256.
template <int NCOL>
__global__ void __launch_bounds__(256)
embed_in_kernel(const float* __restrict__ x, const float* __restrict__ xc, const long long* __restrict__ mask,
                const float* __restrict__ Wt, const float* __restrict__ bias, const float* __restrict__ emask,
                float* __restrict__ h, int n_tok, int D, int H) {
  extern __shared__ float xs[];  // [2D][36]
  const int tok0 = blockIdx.x * 32;
  const int K2 = 2 * D;
  for (int i = threadIdx.x; i < 32 * K2; i += 256) {
    int tt = i / K2, k = i % K2;
    int tok = tok0 + tt;
    float v = 0.f;
    if (tok < n_tok) v = k < D ? x[(size_t)tok * D + k] : xc[(size_t)tok * D + (k - D)];
    xs[k * 36 + tt] = v;
  }
  __syncthreads();
  float acc[NCOL][32];
#pragma unroll
  for (int c = 0; c < NCOL; ++c)
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[c][i] = 0.f;
  int col[NCOL];
#pragma unroll
  for (int c = 0; c < NCOL; ++c) col[c] = threadIdx.x + c * 256;
  for (int k = 0; k < K2; ++k) {
    float w[NCOL];
#pragma unroll
    for (int c = 0; c < NCOL; ++c) w[c] = col[c] < H ? __ldg(Wt + (size_t)k * H + col[c]) : 0.f;
    const float4* xr = reinterpret_cast<const float4*>(xs + k * 36);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 v = xr[i];
#pragma unroll
      for (int c = 0; c < NCOL; ++c) {
        acc[c][4 * i + 0] = fmaf(w[c], v.x, acc[c][4 * i + 0]);
        acc[c][4 * i + 1] = fmaf(w[c], v.y, acc[c][4 * i + 1]);
        acc[c][4 * i + 2] = fmaf(w[c], v.z, acc[c][4 * i + 2]);
        acc[c][4 * i + 3] = fmaf(w[c], v.w, acc[c][4 * i + 3]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < NCOL; ++c) {
    if (col[c] >= H) continue;
    const float bj = bias[col[c]];
    const float e0 = emask[col[c]], e1 = emask[H + col[c]];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      int tok = tok0 + i;
      if (tok < n_tok) h[(size_t)tok * H + col[c]] = acc[c][i] + bj + (mask[tok] ? e1 : e0);
    }
  }
}

// ---- A operand of the tensor-core input embedding: row = [x_hi | x_lo | x_hi | xc_hi | xc_lo | xc_hi] (6D bf16), hi = bf16(v),
// lo = bf16(v - hi).  With the weight packed as [Wx_hi | Wx_hi | Wx_lo | Wc_hi | Wc_hi | Wc_lo] an ordinary bf16 GEMM evaluates
// v_hi w_hi + v_lo w_hi + v_hi w_lo for both inputs (~16 mantissa bits).  One thread per 4 consecutive input floats.
__global__ void __launch_bounds__(256)
split3_embed_kernel(const float4* __restrict__ x, const float4* __restrict__ xc, __nv_bfloat16* __restrict__ a, long long n_tok, int D) {
  const int q = D / 4;  // float4 per row per input
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_tok * 2 * q) return;
  const long long tok = idx / (2 * q);
  const int r = (int)(idx % (2 * q));
  const int which = r / q, j = r % q;
  const float4 v = which == 0 ? x[tok * q + j] : xc[tok * q + j];
  const uint2 hi = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&hi);
  const float2 f01 = __bfloat1622float2(h2[0]), f23 = __bfloat1622float2(h2[1]);
  const uint2 lo = make_uint2(pack_bf16x2(v.x - f01.x, v.y - f01.y), pack_bf16x2(v.z - f23.x, v.w - f23.y));
  uint2* row = reinterpret_cast<uint2*>(a + tok * 6 * D + (long long)which * 3 * D);
  row[j] = hi;
  row[q + j] = lo;
  row[2 * q + j] = hi;
}

// ---- in-place LayerNorm without affine over rows of width H (F.layer_norm at latent_si_v31.py:174, eps 1e-5); warp per row
__global__ void __launch_bounds__(256) layernorm_rows_kernel(float* __restrict__ h, int rows, int H, float eps) {
  int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float* p = h + (size_t)row * H;
  float s = 0.f;
  for (int j = lane; j < H; j += 32) s += p[j];
  float mean = warp_sum(s) / H;
  float v = 0.f;
  for (int j = lane; j < H; j += 32) {
    float d = p[j] - mean;
    v = fmaf(d, d, v);
  }
  float rstd = rsqrtf(warp_sum(v) / H + eps);
  for (int j = lane; j < H; j += 32) p[j] = (p[j] - mean) * rstd;
}

// ---- LayerNorm (no affine, eps 1e-6) + adaLN modulate (mmdit.py:21-22) -> bf16 A operand of the next GEMM
//   u[row, :] = LN(h[row, :]) * (1 + scale[b, :]) + shift[b, :],  b = row / rows_per_sample;  warp per row, H = 128 * HV
// SPLIT3 (final head only): the row is written as [hi | lo | hi] (3H bf16), hi = bf16(u), lo = bf16(u - hi); together with
// a [W_hi | W_hi | W_lo] weight the ordinary bf16 GEMM then evaluates u_hi W_hi + u_lo W_hi + u_hi W_lo, i.e. the head
// projection at ~16 mantissa bits.  The head is 0.4 % of the FLOPs but its operand rounding would otherwise dominate the
// error of the predicted latent (and hence the decoded coordinates).
// (A variant with two rows per warp — 16 lanes per row, 32-byte loads and 16-byte stores per lane — measured 9.2 ms per step against
// 8.4 ms for this one on B200, 4AA: more bytes in flight per warp did not help, the kernel runs at 4.7 TB/s.)
template <int HV, bool SPLIT3 = false>
__global__ void __launch_bounds__(256)
ln_modulate_kernel(const float* __restrict__ h, __nv_bfloat16* __restrict__ u, const float* __restrict__ shift,
                   const float* __restrict__ scale, int mod_stride, int rows, int rows_per_sample) {
  constexpr int H = HV * 128;
  int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* p = reinterpret_cast<const float4*>(h + (size_t)row * H);
  float4 v[HV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < HV; ++i) {
    v[i] = p[lane + 32 * i];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / H);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < HV; ++i) {
    v[i].x -= mean, v[i].y -= mean, v[i].z -= mean, v[i].w -= mean;
    q = fmaf(v[i].x, v[i].x, q), q = fmaf(v[i].y, v[i].y, q), q = fmaf(v[i].z, v[i].z, q), q = fmaf(v[i].w, v[i].w, q);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / H) + 1e-6f);
  const int b = row / rows_per_sample;
  const float4* sh = reinterpret_cast<const float4*>(shift + (size_t)b * mod_stride);
  const float4* sc = reinterpret_cast<const float4*>(scale + (size_t)b * mod_stride);
  uint2* up = reinterpret_cast<uint2*>(u + (size_t)row * H * (SPLIT3 ? 3 : 1));
#pragma unroll
  for (int i = 0; i < HV; ++i) {
    float4 a = __ldg(sc + lane + 32 * i), c = __ldg(sh + lane + 32 * i);
    float y0 = fmaf(v[i].x * rstd, 1.0f + a.x, c.x);
    float y1 = fmaf(v[i].y * rstd, 1.0f + a.y, c.y);
    float y2 = fmaf(v[i].z * rstd, 1.0f + a.z, c.z);
    float y3 = fmaf(v[i].w * rstd, 1.0f + a.w, c.w);
    uint2 hi = make_uint2(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3));
    up[lane + 32 * i] = hi;
    if constexpr (SPLIT3) {
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&hi);
      float2 f01 = __bfloat1622float2(h2[0]), f23 = __bfloat1622float2(h2[1]);
      up[H / 4 + lane + 32 * i] = make_uint2(pack_bf16x2(y0 - f01.x, y1 - f01.y), pack_bf16x2(y2 - f23.x, y3 - f23.y));
      up[2 * (H / 4) + lane + 32 * i] = hi;
    }
  }
}

// ---- probability-flow drift + explicit Euler (transport.py:158-202, integrators.py:119 / torchdiffeq fixed grid)
// Every drift of the reference is linear in (net output m, state x) with time-only coefficients, computed on the host
// in fp64:  v = cm * m + cx * x;   x <- x + dt * v.   Optionally records v and the new state.
__global__ void __launch_bounds__(256)
drift_euler_kernel(const float4* __restrict__ m, float4* __restrict__ x, float4* __restrict__ v_out, float4* __restrict__ state_out,
                   float cm, float cx, float dt, long long n4) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 mv = m[i], xv = x[i], v;
  v.x = fmaf(cm, mv.x, cx * xv.x), v.y = fmaf(cm, mv.y, cx * xv.y);
  v.z = fmaf(cm, mv.z, cx * xv.z), v.w = fmaf(cm, mv.w, cx * xv.w);
  xv.x = fmaf(dt, v.x, xv.x), xv.y = fmaf(dt, v.y, xv.y), xv.z = fmaf(dt, v.z, xv.z), xv.w = fmaf(dt, v.w, xv.w);
  x[i] = xv;
  if (v_out) v_out[i] = v;
  if (state_out) state_out[i] = xv;
}

// ---- setup_conditioning (lightning_base.py:240-263): mask[:, c0:c1] = 1; x_cond = where(mask, latents, mean_{t in cond} latents | 0)
__global__ void __launch_bounds__(256)
conditioning_kernel(const float* __restrict__ lat, float* __restrict__ x_cond, long long* __restrict__ mask, int B, int T, int L,
                    int D, int c0, int c1, int use_mean) {
  // one thread per (b, l, d); loops over T
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long tot = (long long)B * L * D;
  if (idx >= tot) return;
  int d = (int)(idx % D);
  int l = (int)((idx / D) % L);
  int b = (int)(idx / ((long long)D * L));
  const size_t s_t = (size_t)L * D;
  const float* p = lat + (size_t)b * T * s_t + (size_t)l * D + d;
  float fill = 0.f;
  if (use_mean) {
    float s = 0.f;
    for (int t = c0; t < c1; ++t) s += p[(size_t)t * s_t];
    fill = s / (float)(c1 - c0);
  }
  float* o = x_cond + (size_t)b * T * s_t + (size_t)l * D + d;
  for (int t = 0; t < T; ++t) {
    bool c = t >= c0 && t < c1;
    o[(size_t)t * s_t] = c ? p[(size_t)t * s_t] : fill;
    if (d == 0) mask[((size_t)b * T + t) * L + l] = c ? 1 : 0;
  }
}

// ---- K-sample trajectory errors: ADE / FDE of K sampled futures against the ground truth.
// preds [K, B, T, A, D] (sample-major, as K batched sample() calls produce them), target [B, T, A, D]; T = frames after the
// conditioning window, A agents / atoms, D coordinates.  err(k, b, t, a) = || preds - target ||_2 over D.
//   mode 0 (nba.py:191-197,220-225; pedestrian.py idem): per (b, a): ade = min_{k < num_runs} mean_t err, fde = min_{k < num_runs} err(t = T-1)
//   mode 1 (md17.py:158-168):                            per b:      ade = mean_k mean_{t,a} err,      fde = mean_k mean_a err(t = T-1)
// One thread per output element; the reductions are sequential in the reference's index order (k, then t, then a) in fp32.
__global__ void __launch_bounds__(128)
ksample_errors_kernel(const float* __restrict__ preds, const float* __restrict__ target, float* __restrict__ ades, float* __restrict__ fdes,
                      int K, int num_runs, int B, int T, int A, int D, int mode) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_out = mode == 0 ? B * A : B;
  if (idx >= n_out) return;
  const size_t frame = (size_t)A * D, traj = (size_t)T * frame, sample = (size_t)B * traj;
  if (mode == 0) {
    const int b = idx / A, a = idx % A;
    float best_a = INFINITY, best_f = INFINITY;
    for (int k = 0; k < num_runs; ++k) {
      const float* p = preds + (size_t)k * sample + (size_t)b * traj + (size_t)a * D;
      const float* q = target + (size_t)b * traj + (size_t)a * D;
      float sum = 0.f, last = 0.f;
      for (int t = 0; t < T; ++t) {
        float e2 = 0.f;
        for (int d = 0; d < D; ++d) {
          const float df = p[(size_t)t * frame + d] - q[(size_t)t * frame + d];
          e2 = fmaf(df, df, e2);
        }
        last = sqrtf(e2);
        sum += last;
      }
      best_a = fminf(best_a, sum / (float)T);
      best_f = fminf(best_f, last);
    }
    ades[idx] = best_a, fdes[idx] = best_f;
  } else {
    const int b = idx;
    float acc_a = 0.f, acc_f = 0.f;
    for (int k = 0; k < K; ++k) {
      const float* p = preds + (size_t)k * sample + (size_t)b * traj;
      const float* q = target + (size_t)b * traj;
      float sum = 0.f, sum_last = 0.f;
      for (int t = 0; t < T; ++t)
        for (int a = 0; a < A; ++a) {
          float e2 = 0.f;
          for (int d = 0; d < D; ++d) {
            const float df = q[(size_t)t * frame + (size_t)a * D + d] - p[(size_t)t * frame + (size_t)a * D + d];
            e2 = fmaf(df, df, e2);
          }
          const float e = sqrtf(e2);
          sum += e;
          if (t == T - 1) sum_last += e;
        }
      acc_a += sum / (float)(T * A);
      acc_f += sum_last / (float)A;
    }
    ades[idx] = acc_a / (float)K, fdes[idx] = acc_f / (float)K;
  }
}

// ---- out = px * x + pm * m + pw * w  (w nullable) — the update rules of the SDE sampler (integrators.py:29-52: Euler-Maruyama,
// Heun; transport.py:266-299: last step) are all linear in (state, network output, noise) with time-only coefficients, which the
// host evaluates in fp64 (lam_slide_b200/transport.py).  out may alias x.
__global__ void __launch_bounds__(256)
lincomb3_kernel(float4* out, const float4* x, const float4* __restrict__ m, const float4* __restrict__ w,
                float px, float pm, float pw, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 xv = x[i], mv = m ? m[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 o;
  o.x = fmaf(pm, mv.x, px * xv.x), o.y = fmaf(pm, mv.y, px * xv.y), o.z = fmaf(pm, mv.z, px * xv.z), o.w = fmaf(pm, mv.w, px * xv.w);
  if (w) {
    const float4 wv = w[i];
    o.x = fmaf(pw, wv.x, o.x), o.y = fmaf(pw, wv.y, o.y), o.z = fmaf(pw, wv.z, o.z), o.w = fmaf(pw, wv.w, o.w);
  }
  out[i] = o;
}

// ---- out = sum_j coef[j] * src[j], j < n <= 8 (out may be one of the sources): stage combinations y0 + sum_j (dt beta_ij) k_j, the
// solution / mid-point combinations and the interpolant of the Runge-Kutta integrators (lam_slide_b200/odeint.py; torchdiffeq
// rk_common.py: _runge_kutta_step, interp.py).  Pointers and coefficients travel by value.
struct LincombN {
  const float4* src[8];
  float coef[8];
  int n;
};
__global__ void __launch_bounds__(256) lincomb_n_kernel(float4* out, LincombN a, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (j < a.n) {
      const float4 v = a.src[j][i];
      const float c = a.coef[j];
      o.x = fmaf(c, v.x, o.x), o.y = fmaf(c, v.y, o.y), o.z = fmaf(c, v.z, o.z), o.w = fmaf(c, v.w, o.w);
    }
  }
  out[i] = o;
}

// ---- *acc += sum_i ( (sum_j coef[j] src[j][i]) / (atol + rtol * max(|a[i]|, |b[i]|)) )^2 in fp64: the squared RMS error norm of an
// embedded Runge-Kutta step (rk_common.py: _compute_error_ratio) and the norms of the initial step selection (misc.py:
// _select_initial_step, with a = b = y0), fused so that the error vector is never written.  One fp64 atomic per block.
__global__ void __launch_bounds__(256)
rk_error_sumsq_kernel(LincombN e, const float4* __restrict__ a, const float4* __restrict__ b, double rtol, double atol, long long n4,
                      double* __restrict__ acc) {
  double local = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j < e.n) {
        const float4 v = e.src[j][i];
        const float c = e.coef[j];
        o.x = fmaf(c, v.x, o.x), o.y = fmaf(c, v.y, o.y), o.z = fmaf(c, v.z, o.z), o.w = fmaf(c, v.w, o.w);
      }
    }
    const float4 av = a[i], bv = b[i];
    const double r0 = (double)o.x / (atol + rtol * (double)fmaxf(fabsf(av.x), fabsf(bv.x)));
    const double r1 = (double)o.y / (atol + rtol * (double)fmaxf(fabsf(av.y), fabsf(bv.y)));
    const double r2 = (double)o.z / (atol + rtol * (double)fmaxf(fabsf(av.z), fabsf(bv.z)));
    const double r3 = (double)o.w / (atol + rtol * (double)fmaxf(fabsf(av.w), fabsf(bv.w)));
    local += (r0 * r0 + r1 * r1) + (r2 * r2 + r3 * r3);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
  __shared__ double warp_sum[8];
  if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += warp_sum[w];
    atomicAdd(acc, t);
  }
}

}  // namespace lam

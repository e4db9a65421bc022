// Micro-benchmarks (GPU box) behind the design of the tcgen05 temporal-attention kernel (attn_tc.cuh):
//   A  tcgen05.ld throughput per SM by shape (x16 / x32 / x64), warps issuing and loads in flight before tcgen05.wait::ld
//   B  exp2 throughput per SM: MUFU.EX2 alone, FMA-pipe polynomial alone (scalar and packed f32x2), and mixes of the two
//      including the row-sum add and the bf16x2 pack of a softmax inner loop
//   C  the softmax inner loop on TMEM (tcgen05.ld S -> exp2 mix -> sum -> bf16 -> tcgen05.st P), plain and software pipelined
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I lam_slide_b200/csrc scripts/attn_ubench.cu -o /tmp/attn_ubench && /tmp/attn_ubench
#include <cstdio>
#include <cstdlib>
#include "ptx.cuh"
using namespace lam;

__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
      "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),
        "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]),
        "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]),
        "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16_(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
               "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait_() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- packed f32x2 helpers
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float a, float b) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void up2(u64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// 2^x for two values on the FMA pipe (same polynomial as ptx.cuh: poly_exp2)
__device__ __forceinline__ void poly_exp2_x2(float x0, float x1, float& e0, float& e1) {
  const u64 X = pk2(x0, x1);
  const u64 MAGIC = pk2(12582912.f, 12582912.f), NMAGIC = pk2(-12582912.f, -12582912.f), NEG1 = pk2(-1.f, -1.f);
  const u64 t = add2(X, MAGIC);
  const u64 u = add2(t, NMAGIC);
  const u64 f = fma2(u, NEG1, X);
  u64 p = fma2(f, pk2(0.05517146f, 0.05517146f), pk2(0.24261086f, 0.24261086f));
  p = fma2(p, f, pk2(0.69326099f, 0.69326099f));
  p = fma2(p, f, pk2(0.99992809f, 0.99992809f));
  float p0, p1, t0, t1;
  up2(p, p0, p1);
  up2(t, t0, t1);
  e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

struct Span {
  long long t0, t1;
};
__device__ __forceinline__ void span_record(long long* cyc, long long t0, long long t1) {
  __shared__ unsigned long long s_lo, s_hi;
  if (threadIdx.x == 0) s_lo = ~0ull, s_hi = 0ull;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&s_lo, (unsigned long long)t0);
    atomicMax(&s_hi, (unsigned long long)t1);
  }
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = (long long)(s_hi - s_lo);
}

// ================================================================================================ A: tcgen05.ld throughput
// SHAPE: columns per load (16 / 32 / 64); INFL: loads issued before each tcgen05.wait::ld (registers: SHAPE * INFL <= 128)
template <int SHAPE, int INFL, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) ldtm_kernel(long long* cyc, uint32_t* sink, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t taddr = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + ((warp >> 2) & 3) * 128;
  uint32_t r[SHAPE * INFL];
  uint32_t acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < INFL; ++j) {
      const uint32_t a = taddr + ((j * SHAPE) & 127);
      if constexpr (SHAPE == 16) tmem_ld16(a, r + j * SHAPE);
      else if constexpr (SHAPE == 32) tmem_ld32(a, r + j * SHAPE);
      else tmem_ld64(a, r + j * SHAPE);
    }
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < SHAPE * INFL; j += 8) acc += r[j];
  }
  const long long t1 = clock64();
  span_record(cyc, t0, t1);
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc<512>(slot);
  }
}

// ================================================================================================ B: exp2 mixes (registers only)
// Per iteration a thread turns 32 fp32 "logits" into 16 packed bf16x2 "probabilities" + a running sum.
// POLY = number of every 8 values evaluated by the FMA-pipe polynomial (0 = all MUFU, 8 = all polynomial); PACKED: f32x2 polynomial.
template <int POLY, bool PACKED, bool SUMPACK>
__device__ __forceinline__ void softmax32(const float* x, uint32_t* pk, float& l0, float& l1) {
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    float e0, e1;
    const bool poly = ((i >> 1) & 7) * 1 < POLY;  // pairs 0..POLY-1 of every 8 pairs
    if (poly) {
      if constexpr (PACKED) poly_exp2_x2(x[i], x[i + 1], e0, e1);
      else e0 = poly_exp2(x[i]), e1 = poly_exp2(x[i + 1]);
    } else {
      e0 = fast_exp2(x[i]), e1 = fast_exp2(x[i + 1]);
    }
    if constexpr (SUMPACK) {
      u64 l = add2(pk2(l0, l1), pk2(e0, e1));
      up2(l, l0, l1);
      pk[i >> 1] = pack_bf16x2(e0, e1);
    } else {
      pk[i >> 1] = __float_as_uint(e0) ^ __float_as_uint(e1);
    }
  }
}
template <int POLY, bool PACKED, bool SUMPACK>
__global__ void __launch_bounds__(512, 1) exp_kernel(long long* cyc, uint32_t* sink, int iters, float seed) {
  float x[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) x[i] = seed * (float)(i - 16) + 1e-3f * (threadIdx.x & 31);
  float l0 = 0.f, l1 = 0.f;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t pk[16];
    softmax32<POLY, PACKED, SUMPACK>(x, pk, l0, l1);
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= pk[i];
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] += 1e-4f;  // new inputs every iteration (cost: 32 FADD, reported separately by POLY = -1)
  }
  const long long t1 = clock64();
  span_record(cyc, t0, t1);
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc + __float_as_uint(l0 + l1);
}
// the loop overhead of exp_kernel alone (input refresh + xor-reduce), to subtract
__global__ void __launch_bounds__(512, 1) exp_overhead_kernel(long long* cyc, uint32_t* sink, int iters, float seed) {
  float x[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) x[i] = seed * (float)(i - 16) + 1e-3f * (threadIdx.x & 31);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; i += 2) acc ^= __float_as_uint(x[i]) ^ __float_as_uint(x[i + 1]);
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] += 1e-4f;
  }
  const long long t1 = clock64();
  span_record(cyc, t0, t1);
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// ================================================================================================ C: softmax loop on TMEM
// Each warp owns TMEM lanes 32 * (warp % 4) and a 128-column group (warp / 4): per iteration it walks the group in 32-column
// steps: tcgen05.ld S -> exp2 mix -> row sum -> bf16x2 -> tcgen05.st P (16 columns, over the first half of what it just read).
// PIPE: the tcgen05.ld of step k + 1 is issued before the math of step k.
template <int POLY, bool PIPE, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) tmem_softmax_kernel(long long* cyc, uint32_t* sink, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t lane_t = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  // fill all 512 columns of this warp's lanes with small logits (only warps 0..3 of each lane quarter need to do it; all do)
  {
    uint32_t v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(0.25f * (float)(i - 8) + 1e-3f * (threadIdx.x & 31));
    for (int c = 0; c < 512; c += 16) tmem_st16_(lane_t + c, v);
    tmem_st_wait_();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t grp = lane_t + ((warp >> 2) & 3) * 128;
  float l0 = 0.f, l1 = 0.f;
  uint32_t acc = 0;
  const long long t0 = clock64();
  if constexpr (!PIPE) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
      for (int s = 0; s < 4; ++s) {
        uint32_t sv[32], pk[16];
        tmem_ld32(grp + s * 32, sv);
        tmem_ld_wait();
        softmax32<POLY, true, true>(reinterpret_cast<const float*>(sv), pk, l0, l1);
        tmem_st16_(grp + s * 32, pk);
      }
    }
  } else {
    uint32_t sa[32], sb[32], pk[16];
    tmem_ld32(grp, sa);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int s = 0; s < 4; s += 2) {
        tmem_ld_wait();
        tmem_ld32(grp + (s + 1) * 32, sb);
        softmax32<POLY, true, true>(reinterpret_cast<const float*>(sa), pk, l0, l1);
        tmem_st16_(grp + s * 32, pk);
        tmem_ld_wait();
        tmem_ld32(grp + ((s + 2) & 3) * 32, sa);
        softmax32<POLY, true, true>(reinterpret_cast<const float*>(sb), pk, l0, l1);
        tmem_st16_(grp + (s + 1) * 32, pk);
      }
    }
    tmem_ld_wait();
    acc += sa[0];
  }
  tmem_st_wait_();
  const long long t1 = clock64();
  span_record(cyc, t0, t1);
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc + __float_as_uint(l0 + l1);
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc<512>(slot);
  }
}

// ================================================================================================ host
static long long* d_cyc;
static uint32_t* d_sink;
static double avg_cycles(int grid) {
  static long long h[148];
  cudaDeviceSynchronize();
  cudaMemcpy(h, d_cyc, grid * 8, cudaMemcpyDeviceToHost);
  double a = 0;
  for (int i = 0; i < grid; ++i) a += (double)h[i];
  return a / grid;
}
#define CHECK()                                                                  \
  do {                                                                           \
    cudaError_t e = cudaDeviceSynchronize();                                     \
    if (e != cudaSuccess) {                                                      \
      printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__);     \
      exit(1);                                                                   \
    }                                                                            \
  } while (0)

template <int SHAPE, int INFL>
static void run_ldtm() {
  const int iters = 2000;
  for (int nw : {4, 8, 16}) {
    if (nw <= 8) {
      ldtm_kernel<SHAPE, INFL, 256><<<148, nw * 32>>>(d_cyc, d_sink, 10);
      ldtm_kernel<SHAPE, INFL, 256><<<148, nw * 32>>>(d_cyc, d_sink, iters);
    } else if (SHAPE * INFL <= 64) {
      ldtm_kernel<SHAPE, INFL, 512><<<148, nw * 32>>>(d_cyc, d_sink, 10);
      ldtm_kernel<SHAPE, INFL, 512><<<148, nw * 32>>>(d_cyc, d_sink, iters);
    } else {
      continue;
    }
    CHECK();
    const double c = avg_cycles(148);
    printf("A ldtm x%-3d inflight %d  %2d warps: %7.1f B/clk/SM  (%6.1f clk per load)\n", SHAPE, INFL, nw,
           (double)SHAPE * 4 * 32 * INFL * nw * iters / c, c / iters / INFL);
  }
}
template <int POLY, bool PACKED, bool SUMPACK>
static void run_exp(double overhead[3]) {
  const int iters = 2000;
  int k = 0;
  for (int nw : {4, 8, 16}) {
    exp_kernel<POLY, PACKED, SUMPACK><<<148, nw * 32>>>(d_cyc, d_sink, 10, 0.11f);
    exp_kernel<POLY, PACKED, SUMPACK><<<148, nw * 32>>>(d_cyc, d_sink, iters, 0.11f);
    CHECK();
    const double c = avg_cycles(148);
    printf("B exp2 poly %d/8 %s %s  %2d warps: %6.2f exp/clk/SM raw, %6.2f net of loop overhead\n", POLY, PACKED ? "f32x2 " : "scalar",
           SUMPACK ? "+sum+pack" : "         ", nw, 32.0 * 32 * nw * iters / c, 32.0 * 32 * nw * iters / (c - overhead[k]));
    ++k;
  }
}
template <int POLY, bool PIPE>
static void run_tmem_softmax() {
  const int iters = 500;
  for (int nw : {4, 8, 16}) {
    if (nw <= 8) {
      tmem_softmax_kernel<POLY, PIPE, 256><<<148, nw * 32>>>(d_cyc, d_sink, 5);
      tmem_softmax_kernel<POLY, PIPE, 256><<<148, nw * 32>>>(d_cyc, d_sink, iters);
    } else {
      tmem_softmax_kernel<POLY, PIPE, 512><<<148, nw * 32>>>(d_cyc, d_sink, 5);
      tmem_softmax_kernel<POLY, PIPE, 512><<<148, nw * 32>>>(d_cyc, d_sink, iters);
    }
    CHECK();
    const double c = avg_cycles(148);
    printf("C tmem softmax poly %d/8 %s  %2d warps: %6.2f exp/clk/SM  (%6.1f B/clk/SM of tcgen05.ld)\n", POLY, PIPE ? "pipelined" : "plain    ", nw,
           128.0 * 32 * nw * iters / c, 128.0 * 32 * nw * iters * 4 / c);
  }
}

int main() {
  cudaMalloc(&d_cyc, 148 * 8);
  cudaMalloc(&d_sink, 148 * 512 * 4);
  run_ldtm<16, 1>();
  run_ldtm<16, 4>();
  run_ldtm<32, 1>();
  run_ldtm<32, 2>();
  run_ldtm<32, 4>();
  run_ldtm<64, 1>();
  run_ldtm<64, 2>();
  double ov[3];
  {
    int k = 0;
    for (int nw : {4, 8, 16}) {
      exp_overhead_kernel<<<148, nw * 32>>>(d_cyc, d_sink, 10, 0.11f);
      exp_overhead_kernel<<<148, nw * 32>>>(d_cyc, d_sink, 2000, 0.11f);
      CHECK();
      ov[k] = avg_cycles(148);
      printf("B loop overhead %2d warps: %.0f clk per 2000 iterations\n", nw, ov[k]);
      ++k;
    }
  }
  run_exp<0, true, false>(ov);
  run_exp<8, false, false>(ov);
  run_exp<8, true, false>(ov);
  run_exp<0, true, true>(ov);
  run_exp<2, true, true>(ov);
  run_exp<3, true, true>(ov);
  run_exp<4, true, true>(ov);
  run_exp<5, true, true>(ov);
  run_exp<4, false, true>(ov);
  run_exp<8, true, true>(ov);
  run_tmem_softmax<0, false>();
  run_tmem_softmax<0, true>();
  run_tmem_softmax<2, true>();
  run_tmem_softmax<3, true>();
  run_tmem_softmax<4, true>();
  run_tmem_softmax<4, false>();
  return 0;
}

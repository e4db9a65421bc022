// liblamslide.so — host runtime + C ABI (include/lamslide.h) of the B200-native LaM-SLidE sampling hot path.
// One translation unit: weight packing, TMA descriptors, kernel orchestration for LatentSIV3.forward, the SiT Euler
// ODE loop, setup_conditioning and the first-stage encode / decode.  No CPU fallback: every entry point launches the
// sm_100a kernels in gemm_tc.cuh / attn.cuh / elementwise.cuh / first_stage.cuh or fails with an error code.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/lamslide.h"
#include "gemm_tc.cuh"
#include "gemm_ws.cuh"
#include "mlp_fused.cuh"
#include "attn.cuh"
#include "attn_tc.cuh"
#include "attn_tc3.cuh"
#include "elementwise.cuh"
#include "first_stage.cuh"
#include "linear_tc5.cuh"

using namespace lam;

// ================================================================================================ error handling
static thread_local std::string g_err;
static thread_local int64_t g_launches = 0;
static thread_local std::unordered_map<std::string, int64_t> g_named_launches;  // lamslide_debug_kernel_count (tests)
#define COUNT_KERNEL(name) (++g_named_launches[name])

static int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CUDA_TRY(expr)                                                                                          \
  do {                                                                                                          \
    cudaError_t _e = (expr);                                                                                    \
    if (_e != cudaSuccess) return fail(LAMSLIDE_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                       __FILE__, __LINE__);                                                     \
  } while (0)
#define LAUNCH_CHECK()                                                                                          \
  do {                                                                                                          \
    ++g_launches;                                                                                               \
    cudaError_t _e = cudaGetLastError();                                                                        \
    if (_e != cudaSuccess) return fail(LAMSLIDE_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                                       __FILE__, __LINE__);                                                     \
  } while (0)
#define TRY(expr)            \
  do {                       \
    int _r = (expr);         \
    if (_r != 0) return _r;  \
  } while (0)

// ================================================================================================ per-kernel-class timing
// Optional CUDA-event timing of kernel classes on the launching stream (bench.py's roofline / time shares).  Off by default.
enum ProfCat { PC_EMBED = 0, PC_VEC, PC_LNMOD, PC_LINEAR1, PC_ATTN_SPATIAL, PC_ATTN_TEMPORAL, PC_LINEAR2, PC_HEAD, PC_EULER,
               PC_ENCODE, PC_DECODE, PC_COND, PC_COUNT };
static const char* kProfNames[PC_COUNT] = {"embed_in", "vec_path", "ln_modulate", "gemm_linear1", "attn_spatial", "attn_temporal",
                                           "gemm_linear2", "gemm_head", "drift_euler", "fs_encode", "fs_decode", "conditioning"};
struct ProfRec {
  int cat;
  cudaEvent_t a, b;
};
static thread_local bool g_prof_on = false;
static thread_local std::vector<ProfRec> g_prof;
struct ProfScope {
  cudaStream_t st;
  bool on;
  ProfRec r;
  ProfScope(int cat, cudaStream_t s) : st(s), on(g_prof_on) {
    if (!on) return;
    r.cat = cat;
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
  }
  ~ProfScope() {
    if (!on) return;
    cudaEventRecord(r.b, st);
    g_prof.push_back(r);
  }
};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ================================================================================================ state dict / device arena
struct StateDict {
  std::unordered_map<std::string, const lamslide_tensor*> m;
  StateDict(const lamslide_tensor* t, int n) {
    for (int i = 0; i < n; ++i) m[t[i].name] = &t[i];
  }
  // returns the tensor or nullptr (+ error text) — shape is checked when dims are given (0-terminated list)
  const lamslide_tensor* get(const std::string& name, std::initializer_list<int64_t> shape) const {
    auto it = m.find(name);
    if (it == m.end()) {
      fail(LAMSLIDE_ERR_MISSING, "state dict key '%s' is missing", name.c_str());
      return nullptr;
    }
    const lamslide_tensor* t = it->second;
    int64_t want = 1, have = 1;
    for (auto s : shape) want *= s;
    for (int i = 0; i < t->ndim; ++i) have *= t->shape[i];
    if (want != have) {
      fail(LAMSLIDE_ERR_MISSING, "state dict key '%s' has %lld elements, expected %lld", name.c_str(), (long long)have,
           (long long)want);
      return nullptr;
    }
    return t;
  }
};

struct Arena {
  std::vector<void*> blocks;
  ~Arena() {
    for (void* p : blocks) cudaFree(p);
  }
  int upload(const void* host, size_t bytes, void** out) {
    void* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, align_up(bytes, 256)));
    blocks.push_back(d);
    CUDA_TRY(cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice));
    *out = d;
    return 0;
  }
  int upload_f32(const float* host, size_t n, float** out) { return upload(host, n * sizeof(float), (void**)out); }
  int upload_bf16(const float* host, size_t n, __nv_bfloat16** out) {
    std::vector<__nv_bfloat16> tmp(n);
    for (size_t i = 0; i < n; ++i) tmp[i] = __float2bfloat16_rn(host[i]);
    return upload(tmp.data(), n * sizeof(__nv_bfloat16), (void**)out);
  }
};

// ================================================================================================ TMA descriptors
static PFN_cuTensorMapEncodeTiled_v12000 get_tmap_encoder() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

// row-major [rows, cols] bf16, tile box_rows x 64 columns, 128-byte swizzle (matches umma_desc_sw128)
static int make_tmap(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  auto enc = get_tmap_encoder();
  if (!enc) return fail(LAMSLIDE_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  if (cols % 64 != 0) return fail(LAMSLIDE_ERR_INVALID, "GEMM K = %llu must be a multiple of 64", (unsigned long long)cols);
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(LAMSLIDE_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

// general 2-D row-major tensor map: [rows, cols] elements of `elem_bytes`, box = box_cols x box_rows (TMA stores / reduces
// of the warp-specialised GEMM epilogues; the swizzle mode must match gemm_ws.cuh: stage_off)
static int make_tmap_ex(CUtensorMap* map, const void* ptr, CUtensorMapDataType dt, int elem_bytes, uint64_t rows, uint64_t cols,
                        uint32_t box_cols, uint32_t box_rows, CUtensorMapSwizzle sw) {
  auto enc = get_tmap_encoder();
  if (!enc) return fail(LAMSLIDE_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, dt, 2, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(LAMSLIDE_ERR_CUDA, "cuTensorMapEncodeTiled (store map) failed with CUresult %d", (int)r);
  return 0;
}
// row-major fp32 [rows, cols] with a row pitch of `ld` floats, tile box_rows x 32 columns (128 bytes), 128-byte swizzle: the
// kind::tf32 operands of linear_tc5.cuh.  Columns / rows past the tensor are zero-filled by the TMA.
static int make_tmap_f32(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  auto enc = get_tmap_encoder();
  if (!enc) return fail(LAMSLIDE_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 4};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(LAMSLIDE_ERR_CUDA, "cuTensorMapEncodeTiled (fp32 operand) failed with CUresult %d", (int)r);
  return 0;
}
// A/B switches and profiling aids read from the environment exist only in debug builds (-DLAMSLIDE_DEBUG_KNOBS, see
// lam_slide_b200/build.py); the shipped library takes no behaviour from the environment.
#ifdef LAMSLIDE_DEBUG_KNOBS
static bool env_flag(const char* name) {
  const char* v = getenv(name);
  return v && v[0] && v[0] != '0';
}
static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v && v[0] ? atoi(v) : dflt;
}
#else
static constexpr bool env_flag(const char*) { return false; }
static constexpr int env_int(const char*, int dflt) { return dflt; }
#endif
// cudaFuncSetAttribute is per device: remember (device, kernel) pairs, not a process-wide flag — one process may hold handles on
// several GPUs (include/lamslide.h: a handle belongs to the device that was current at *_create).
static int ensure_dynamic_smem(const void* kernel, int bytes, bool carveout_max = false) {
  static std::mutex mu;
  static std::set<std::pair<int, const void*>> done;
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  if (done.count({dev, kernel})) return 0;
  CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  if (carveout_max) CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  done.insert({dev, kernel});
  return 0;
}
static int num_sms() {  // of the current device
  static int cache[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return 148;
  if (!cache[dev]) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cache[dev] = n > 0 ? n : 148;
  }
  return cache[dev];
}

// ================================================================================================ GEMM launch
template <int BN>
struct StagesFor {
  static constexpr int value = BN >= 192 ? 2 : (BN >= 96 ? 3 : 4);
};

template <int BN, class Epi>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, int rows, int N, int K, const typename Epi::Params& ep,
                       cudaStream_t st) {
  constexpr int STAGES = StagesFor<BN>::value;
  using SM = GemmSmem<BN, STAGES>;
  auto kern = gemm_tc_kernel<BN, STAGES, Epi>;
  TRY(ensure_dynamic_smem((const void*)kern, SM::kTotal));
  dim3 grid(N / BN, cdiv(rows, kBlockM));
  kern<<<grid, kGemmThreads, SM::kTotal, st>>>(ta, tb, K / kBlockK, ep);
  LAUNCH_CHECK();
  return 0;
}

// persistent warp-specialised GEMM (gemm_ws.cuh).  Returns 1 (and launches nothing) when the shape does not fit its
// shared-memory plan; the caller then uses the one-tile-per-CTA kernel above.  tb_half: the weight map with a box of BN / 2
// rows — when given (and the grid is large enough) the kernel runs as 2-CTA clusters that multicast the weight tiles.
template <int BN, class Epi>
static int launch_gemm_ws(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap* tb_half, const CUtensorMap& o0,
                          const CUtensorMap& o1, int rows, int N, int K, const typename Epi::Params& ep, cudaStream_t st) {
  constexpr int kSmemMax = 232448;  // 227 KB of dynamic shared memory per CTA on sm_100
  const int kblocks = K / kBlockK;
  const int cf = Epi::smem_floats(ep);
  const int mblocks = cdiv(rows, kBlockM);
  static const bool no_cluster = env_flag("LAMSLIDE_NO_CLUSTER");
  const bool use_pair = tb_half && !no_cluster && mblocks >= 2 && num_sms() >= 2 && BN % 16 == 0;
  const int cl = use_pair ? 2 : 1;
  int a_res = kblocks <= kWsMaxKBlocksResident ? 1 : 0, stages = 0;
  static const int max_stages = env_int("LAMSLIDE_WS_STAGES", 8);  // profiling aid
  static const bool no_resident = env_flag("LAMSLIDE_WS_NO_RESIDENT");
  if (no_resident) a_res = 0;
  for (int pass = 0; pass < 2 && !stages; ++pass) {
    for (int s = std::min(8, max_stages); s >= 2; --s)
      if (ws_smem_plan(BN, kblocks, s, a_res, cf, Epi::kStageBytes, cl).total <= kSmemMax) {
        stages = s;
        break;
      }
    if (!stages) {
      if (!a_res) return 1;
      a_res = 0;
    }
  }
  if (!stages) return 1;
  const WsSmemPlan plan = ws_smem_plan(BN, kblocks, stages, a_res, cf, Epi::kStageBytes, cl);
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kWsThreads);
  cfg.dynamicSmemBytes = plan.total;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  if (use_pair) {
    auto kern = gemm_ws_kernel<BN, 2, Epi>;
    TRY(ensure_dynamic_smem((const void*)kern, kSmemMax));
    static const int grid_cap = env_int("LAMSLIDE_WS_GRID", 1 << 30);  // profiling aid
    const int grid = std::min(std::min(num_sms(), grid_cap) / 2 * 2, cdiv(mblocks, 2) * 2);
    cfg.gridDim = dim3(grid);
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, ta, *tb_half, o0, o1, mblocks, N / BN, kblocks, stages, a_res, ep));
  } else {
    auto kern = gemm_ws_kernel<BN, 1, Epi>;
    TRY(ensure_dynamic_smem((const void*)kern, kSmemMax));
    static const int grid_cap = env_int("LAMSLIDE_WS_GRID", 1 << 30);  // profiling aid
    cfg.gridDim = dim3(std::min(std::min(num_sms(), grid_cap), mblocks));
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, ta, tb, o0, o1, mblocks, N / BN, kblocks, stages, a_res, ep));
  }
  LAUNCH_CHECK();
  return 0;
}

// ================================================================================================ second stage handle
struct BlockWeights {
  __nv_bfloat16* w1 = nullptr;  // [3H+M, H]
  __nv_bfloat16* w2 = nullptr;  // [H, H+M]
  float *b1 = nullptr, *b2 = nullptr, *gq = nullptr, *gk = nullptr;
  float gq_h[32] = {0}, gk_h[32] = {0};  // host copies of the QK-norm scales (passed by value to the persistent linear1 kernel)
  float logit_bound = 0.f;  // max |q.k| * hd^-0.5 * log2(e) after QK-RMSNorm: hd^0.5 * log2(e) * max|gq| * max|gk|
  CUtensorMap tm_w1, tm_w2;
  CUtensorMap tm_w1_h, tm_w2_h;  // boxes of half a tile: in a CTA pair each CTA stages half of every weight tile
  CUtensorMap tm_w1_u, tm_w2_u;  // half-unit boxes of the fused MLP kernel (mlp_fused.cuh): 64 / NU/2 rows x 64 columns
};

struct lamslide_backbone {
  lamslide_backbone_config cfg;
  int device = 0;
  int H = 0, M = 0, D = 0, hd = 0, heads = 0, depth = 0;
  int bn1 = 0, bn2 = 0, bn_out = 0;
  int mod_width = 0;  // depth * 6H + 2H
  Arena arena;
  float *wt_in = nullptr, *b_in = nullptr, *emask = nullptr;
  __nv_bfloat16* w_emb = nullptr;  // [H, 6D] hi/lo split of [Wx | Wc] for the tensor-core input embedding (0 = not available)
  CUtensorMap tm_wemb, tm_wemb_h;
  int bn_emb = 0;
  float *time_w1 = nullptr, *time_b1 = nullptr, *time_w2 = nullptr, *time_b2 = nullptr;
  float *vec_w1 = nullptr, *vec_b1 = nullptr, *vec_w2 = nullptr, *vec_b2 = nullptr;
  float *mod_w = nullptr, *mod_b = nullptr;
  std::vector<BlockWeights> blocks;  // 2 * depth: spatial, temporal, spatial, ...
  __nv_bfloat16* w_out = nullptr;
  float* b_out = nullptr;
  CUtensorMap tm_wout;
};

struct BackboneWorkspace {
  float* h;
  __nv_bfloat16* u;
  __nv_bfloat16* qkv;
  __nv_bfloat16* act;
  float* m_out;
  float *e, *hid, *svec, *yhid, *yvec, *mod, *tvec;  // per-sample vector path, sized for n_vec network evaluations at once
  int n_vec;
  float *cos_s, *sin_s, *cos_t, *sin_t;
  size_t bytes;
};

// n_vec: network evaluations whose per-sample vectors (timestep embedding -> time_in -> all modulations) are computed in one pass
// (lamslide_ode_sample knows its whole time grid in advance)
static BackboneWorkspace plan_workspace(const lamslide_backbone* bb, void* base, int B, int T, int L, int n_vec = 1) {
  BackboneWorkspace w;
  size_t off = 0;
  const size_t n = (size_t)B * T * L;
  auto take = [&](size_t bytes) {
    void* p = base ? (void*)((uint8_t*)base + off) : nullptr;
    off += align_up(bytes, 1024);
    return p;
  };
  const int H = bb->H, M = bb->M, D = bb->D, half = bb->hd / 2;
  w.h = (float*)take(n * H * 4);
  w.u = (__nv_bfloat16*)take(n * H * 2);
  w.qkv = (__nv_bfloat16*)take(n * 3 * H * 2);
  w.act = (__nv_bfloat16*)take(n * (size_t)(H + M) * 2);
  w.m_out = (float*)take(n * D * 4);
  const size_t R = (size_t)B * n_vec;
  w.n_vec = n_vec;
  w.e = (float*)take(R * 256 * 4);
  w.hid = (float*)take(R * H * 4);
  w.svec = (float*)take(R * H * 4);
  w.yhid = (float*)take((size_t)B * H * 4);
  w.yvec = (float*)take((size_t)B * H * 4);
  w.mod = (float*)take(R * bb->mod_width * 4);
  w.tvec = (float*)take(R * 4);
  w.cos_s = (float*)take((size_t)L * half * 4);
  w.sin_s = (float*)take((size_t)L * half * 4);
  w.cos_t = (float*)take((size_t)T * half * 4);
  w.sin_t = (float*)take((size_t)T * half * 4);
  w.bytes = off;
  return w;
}

static int plain_bn_for(int N) {
  for (int bn : {256, 192, 128, 96, 64, 48, 32, 16})
    if (N % bn == 0) return bn;
  return 0;
}

static int pick_bn(std::initializer_list<int> cands, int a, int b, int c) {
  for (int bn : cands)
    if (a % bn == 0 && (b == 0 || b % bn == 0) && (c == 0 || bn % c == 0)) return bn;
  return 0;
}

extern "C" int lamslide_backbone_create(const lamslide_backbone_config* cfg, const lamslide_tensor* tensors, int32_t n_tensors,
                                        lamslide_backbone** out) {
  if (!cfg || !tensors || !out) return fail(LAMSLIDE_ERR_INVALID, "null argument");
  *out = nullptr;
  const int H = cfg->hidden_size, D = cfg->in_dim, heads = cfg->num_heads, M = cfg->mlp_hidden, depth = cfg->depth;
  if (heads <= 0 || H % heads != 0)  // latent_si_v31.py:92-95
    return fail(LAMSLIDE_ERR_INVALID, "Hidden size %d must be divisible by num_heads %d", H, heads);
  const int hd = H / heads;
  if (hd != 16 && hd != 24 && hd != 32) return fail(LAMSLIDE_ERR_INVALID, "head_dim %d unsupported (16, 24, 32)", hd);
  if (H % 128 != 0 || H > 512) return fail(LAMSLIDE_ERR_INVALID, "hidden_size %d unsupported (multiple of 128, <= 512)", H);
  if (M % 64 != 0 || M <= 0) return fail(LAMSLIDE_ERR_INVALID, "mlp hidden %d must be a positive multiple of 64", M);
  if (D % 16 != 0 || D > 256 || D <= 0) return fail(LAMSLIDE_ERR_INVALID, "in_dim %d unsupported (multiple of 16, <= 256)", D);
  if (depth <= 0) return fail(LAMSLIDE_ERR_INVALID, "depth must be positive");
  const int bn1 = hd == 24 ? pick_bn({192, 96}, H, M, hd) : pick_bn({128, 64}, H, M, hd);
  const int bn2 = pick_bn({192, 128, 96, 64}, H, 0, 0);
  const int bn_out = plain_bn_for(D);
  if (!bn1 || !bn2 || !bn_out) return fail(LAMSLIDE_ERR_INVALID, "no GEMM tiling for hidden_size %d / mlp %d / head_dim %d", H, M, hd);

  auto* bb = new lamslide_backbone();
  std::unique_ptr<lamslide_backbone> guard(bb);
  bb->cfg = *cfg;
  bb->H = H, bb->M = M, bb->D = D, bb->hd = hd, bb->heads = heads, bb->depth = depth;
  bb->bn1 = bn1, bb->bn2 = bn2, bb->bn_out = bn_out;
  bb->mod_width = depth * 6 * H + 2 * H;
  CUDA_TRY(cudaGetDevice(&bb->device));
  StateDict sd(tensors, n_tensors);
  Arena& A = bb->arena;

#define GET(var, name, ...)                         \
  const lamslide_tensor* var = sd.get(name, {__VA_ARGS__}); \
  if (!var) return LAMSLIDE_ERR_MISSING;

  {  // input embedding: Wt[k, j] with k over [x (D) | x_cond (D)], bias = bx + bc
    GET(wx, "x_in.weight", H, D);
    GET(bx, "x_in.bias", H);
    GET(wc, "cond_to_emb.weight", H, D);
    GET(bc, "cond_to_emb.bias", H);
    GET(em, "mask_to_emb.weight", 2, H);
    std::vector<float> wt((size_t)2 * D * H), bsum(H);
    for (int j = 0; j < H; ++j) {
      for (int k = 0; k < D; ++k) {
        wt[(size_t)k * H + j] = wx->data[(size_t)j * D + k];
        wt[(size_t)(D + k) * H + j] = wc->data[(size_t)j * D + k];
      }
      bsum[j] = bx->data[j] + bc->data[j];
    }
    TRY(A.upload_f32(wt.data(), wt.size(), &bb->wt_in));
    TRY(A.upload_f32(bsum.data(), H, &bb->b_in));
    TRY(A.upload_f32(em->data, 2 * H, &bb->emask));
    // tensor-core variant: [Wx_hi | Wx_hi | Wx_lo | Wc_hi | Wc_hi | Wc_lo], K = 6D (must tile by 64 and fit the act buffer)
    bb->bn_emb = pick_bn({192, 128, 64}, H, 0, 0);
    if ((6 * D) % 64 == 0 && 6 * D <= H + M && bb->bn_emb) {
      std::vector<__nv_bfloat16> we((size_t)H * 6 * D);
      for (int j = 0; j < H; ++j)
        for (int k = 0; k < D; ++k) {
          const float vx = wx->data[(size_t)j * D + k], vc = wc->data[(size_t)j * D + k];
          const __nv_bfloat16 xh = __float2bfloat16_rn(vx), ch = __float2bfloat16_rn(vc);
          const __nv_bfloat16 xl = __float2bfloat16_rn(vx - __bfloat162float(xh)), cl = __float2bfloat16_rn(vc - __bfloat162float(ch));
          __nv_bfloat16* r = &we[(size_t)j * 6 * D];
          r[k] = xh, r[D + k] = xh, r[2 * D + k] = xl;
          r[3 * D + k] = ch, r[4 * D + k] = ch, r[5 * D + k] = cl;
        }
      TRY(A.upload(we.data(), we.size() * sizeof(__nv_bfloat16), (void**)&bb->w_emb));
      TRY(make_tmap(&bb->tm_wemb, bb->w_emb, H, 6 * D, bb->bn_emb));
      TRY(make_tmap(&bb->tm_wemb_h, bb->w_emb, H, 6 * D, bb->bn_emb / 2));
    }
  }
  {
    GET(w1, "time_in.in_layer.weight", H, 256);
    GET(b1, "time_in.in_layer.bias", H);
    GET(w2, "time_in.out_layer.weight", H, H);
    GET(b2, "time_in.out_layer.bias", H);
    TRY(A.upload_f32(w1->data, (size_t)H * 256, &bb->time_w1));
    TRY(A.upload_f32(b1->data, H, &bb->time_b1));
    TRY(A.upload_f32(w2->data, (size_t)H * H, &bb->time_w2));
    TRY(A.upload_f32(b2->data, H, &bb->time_b2));
  }
  if (cfg->vec_in_dim > 0) {
    const int V = cfg->vec_in_dim;
    GET(w1, "vec_in.in_layer.weight", H, V);
    GET(b1, "vec_in.in_layer.bias", H);
    GET(w2, "vec_in.out_layer.weight", H, H);
    GET(b2, "vec_in.out_layer.bias", H);
    TRY(A.upload_f32(w1->data, (size_t)H * V, &bb->vec_w1));
    TRY(A.upload_f32(b1->data, H, &bb->vec_b1));
    TRY(A.upload_f32(w2->data, (size_t)H * H, &bb->vec_w2));
    TRY(A.upload_f32(b2->data, H, &bb->vec_b2));
  }
  {  // every modulation.lin of every layer + the final adaLN in ONE [depth*6H + 2H, H] matrix: one launch per ODE step
    std::vector<float> mw((size_t)bb->mod_width * H), mb(bb->mod_width);
    for (int i = 0; i < depth; ++i) {
      std::string p = "blocks." + std::to_string(i) + ".modulation.lin.";
      GET(w, p + "weight", 6 * H, H);
      GET(b, p + "bias", 6 * H);
      memcpy(&mw[(size_t)i * 6 * H * H], w->data, (size_t)6 * H * H * 4);
      memcpy(&mb[(size_t)i * 6 * H], b->data, (size_t)6 * H * 4);
    }
    GET(w, "adaLN_modulation.1.weight", 2 * H, H);
    GET(b, "adaLN_modulation.1.bias", 2 * H);
    memcpy(&mw[(size_t)depth * 6 * H * H], w->data, (size_t)2 * H * H * 4);
    memcpy(&mb[(size_t)depth * 6 * H], b->data, (size_t)2 * H * 4);
    TRY(A.upload_f32(mw.data(), mw.size(), &bb->mod_w));
    TRY(A.upload_f32(mb.data(), mb.size(), &bb->mod_b));
  }
  bb->blocks.resize(2 * depth);
  for (int i = 0; i < depth; ++i) {
    for (int s = 0; s < 2; ++s) {
      std::string p = "blocks." + std::to_string(i) + (s == 0 ? ".spatial_block." : ".temporal_block.");
      BlockWeights& bw = bb->blocks[2 * i + s];
      GET(w1, p + "linear1.weight", 3 * H + M, H);
      GET(b1, p + "linear1.bias", 3 * H + M);
      GET(w2, p + "linear2.weight", H, H + M);
      GET(b2, p + "linear2.bias", H);
      GET(gq, p + "norm.query_norm.scale", hd);
      GET(gk, p + "norm.key_norm.scale", hd);
      TRY(A.upload_bf16(w1->data, (size_t)(3 * H + M) * H, &bw.w1));
      TRY(A.upload_bf16(w2->data, (size_t)H * (H + M), &bw.w2));
      TRY(A.upload_f32(b1->data, 3 * H + M, &bw.b1));
      TRY(A.upload_f32(b2->data, H, &bw.b2));
      TRY(A.upload_f32(gq->data, hd, &bw.gq));
      TRY(A.upload_f32(gk->data, hd, &bw.gk));
      for (int j = 0; j < hd; ++j) bw.gq_h[j] = gq->data[j], bw.gk_h[j] = gk->data[j];
      {
        float mq = 0.f, mk = 0.f;
        for (int j = 0; j < hd; ++j) mq = std::max(mq, std::fabs(gq->data[j])), mk = std::max(mk, std::fabs(gk->data[j]));
        bw.logit_bound = std::sqrt((float)hd) * 1.4426950408889634f * mq * mk * 1.02f;  // 2 % slack: bf16 rounding of q, k
      }
      TRY(make_tmap(&bw.tm_w1, bw.w1, 3 * H + M, H, bn1));
      TRY(make_tmap(&bw.tm_w2, bw.w2, H, H + M, bn2));
      TRY(make_tmap(&bw.tm_w1_h, bw.w1, 3 * H + M, H, bn1 / 2));
      TRY(make_tmap(&bw.tm_w2_h, bw.w2, H, H + M, bn2 / 2));
      TRY(make_tmap(&bw.tm_w1_u, bw.w1, 3 * H + M, H, 64));
      TRY(make_tmap(&bw.tm_w2_u, bw.w2, H, H + M, fused_mlp_out_unit(H) / 2));
    }
  }
  {
    GET(w, "linear.weight", D, H);
    GET(b, "linear.bias", D);
    // split head weight [W_hi | W_hi | W_lo] (see ln_modulate_kernel<.., SPLIT3>)
    std::vector<__nv_bfloat16> w3((size_t)D * 3 * H);
    for (int r = 0; r < D; ++r)
      for (int k = 0; k < H; ++k) {
        float v = w->data[(size_t)r * H + k];
        __nv_bfloat16 hi = __float2bfloat16_rn(v);
        __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        w3[(size_t)r * 3 * H + k] = hi;
        w3[(size_t)r * 3 * H + H + k] = hi;
        w3[(size_t)r * 3 * H + 2 * H + k] = lo;
      }
    TRY(A.upload(w3.data(), w3.size() * sizeof(__nv_bfloat16), (void**)&bb->w_out));
    TRY(A.upload_f32(b->data, D, &bb->b_out));
    TRY(make_tmap(&bb->tm_wout, bb->w_out, D, 3 * H, bn_out));
  }
#undef GET
  *out = guard.release();
  return 0;
}

extern "C" void lamslide_backbone_destroy(lamslide_backbone* h) { delete h; }

extern "C" size_t lamslide_backbone_workspace_bytes(const lamslide_backbone* h, int32_t B, int32_t T, int32_t L) {
  if (!h || B <= 0 || T <= 0 || L <= 0) return 0;
  return plan_workspace(h, nullptr, B, T, L).bytes;
}

// evaluations per pass of the batched vector path of lamslide_ode_sample: all of them when that costs little memory (4AA, B = 64:
// 4.3 MB per evaluation), at most 16, and no more than 256 MB of modulation vectors
static int ode_vec_chunk(const lamslide_backbone* bb, int B, int n_evals) {
  const size_t per_eval = (size_t)B * bb->mod_width * 4;
  const int by_mem = (int)std::max<size_t>(1, ((size_t)256 << 20) / per_eval);
  return std::max(1, std::min(std::min(n_evals, 16), by_mem));
}

extern "C" size_t lamslide_ode_workspace_bytes(const lamslide_backbone* h, int32_t B, int32_t T, int32_t L, int32_t num_steps) {
  if (!h || B <= 0 || T <= 0 || L <= 0 || num_steps < 2) return 0;
  return plan_workspace(h, nullptr, B, T, L, ode_vec_chunk(h, B, num_steps - 1)).bytes;
}

template <int HD>
static int launch_linear1(const lamslide_backbone* bb, const CUtensorMap& ta, const BlockWeights& bw, int rows,
                          const typename EpiLinear1<HD>::Params& ep, cudaStream_t st) {
  const int N = 3 * bb->H + bb->M, K = bb->H;
  if constexpr (HD == 24) {
    if (bb->bn1 == 192) return launch_gemm<192, EpiLinear1<24>>(ta, bw.tm_w1, rows, N, K, ep, st);
    return launch_gemm<96, EpiLinear1<24>>(ta, bw.tm_w1, rows, N, K, ep, st);
  } else {
    if (bb->bn1 == 128) return launch_gemm<128, EpiLinear1<HD>>(ta, bw.tm_w1, rows, N, K, ep, st);
    return launch_gemm<64, EpiLinear1<HD>>(ta, bw.tm_w1, rows, N, K, ep, st);
  }
}

static int launch_linear2(const lamslide_backbone* bb, const CUtensorMap& ta, const BlockWeights& bw, int rows,
                          const EpiLinear2::Params& ep, cudaStream_t st) {
  const int N = bb->H, K = bb->H + bb->M;
  switch (bb->bn2) {
    case 192: return launch_gemm<192, EpiLinear2>(ta, bw.tm_w2, rows, N, K, ep, st);
    case 128: return launch_gemm<128, EpiLinear2>(ta, bw.tm_w2, rows, N, K, ep, st);
    case 96: return launch_gemm<96, EpiLinear2>(ta, bw.tm_w2, rows, N, K, ep, st);
    default: return launch_gemm<64, EpiLinear2>(ta, bw.tm_w2, rows, N, K, ep, st);
  }
}

// linear1 (q | k | v columns only) with the spatial attention fused into its epilogue (gemm_ws.cuh: EpiLinear1Ws<HD, AL>): sequences of
// AL consecutive rows.  Returns 1 when the shape is not covered (caller: plain linear1 + attention kernel).
template <int HD, int AL>
static int launch_linear1_attn(const lamslide_backbone* bb, const CUtensorMap& ta, const BlockWeights& bw, const CUtensorMap& qkv_st,
                               const CUtensorMap& act_st, int rows, const typename EpiLinear1Ws<HD, AL>::Params& ep, cudaStream_t st) {
  const int N = 3 * bb->H, K = bb->H;
  if (ep.M != 0 || bb->H % bb->bn1 != 0 || rows % AL != 0) return 1;
  if constexpr (HD == 24) {
    if (bb->bn1 == 192) return launch_gemm_ws<192, EpiLinear1Ws<24, AL>>(ta, bw.tm_w1, &bw.tm_w1_h, qkv_st, act_st, rows, N, K, ep, st);
  } else {
    if (bb->bn1 == 128) return launch_gemm_ws<128, EpiLinear1Ws<HD, AL>>(ta, bw.tm_w1, &bw.tm_w1_h, qkv_st, act_st, rows, N, K, ep, st);
  }
  return 1;
}

// warp-specialised variants; return 1 when the tiling is not covered (caller falls back to the kernels above)
template <int HD>
static int launch_linear1_ws(const lamslide_backbone* bb, const CUtensorMap& ta, const BlockWeights& bw, const CUtensorMap& qkv_st,
                             const CUtensorMap& act_st, int rows, const typename EpiLinear1Ws<HD>::Params& ep, cudaStream_t st) {
  const int N = 3 * bb->H + ep.M, K = bb->H;  // ep.M = 0: q | k | v only (the MLP half runs in the fused kernel)
  if constexpr (HD == 24) {
    if (bb->bn1 == 192) return launch_gemm_ws<192, EpiLinear1Ws<24>>(ta, bw.tm_w1, &bw.tm_w1_h, qkv_st, act_st, rows, N, K, ep, st);
  } else {
    if (bb->bn1 == 128) return launch_gemm_ws<128, EpiLinear1Ws<HD>>(ta, bw.tm_w1, &bw.tm_w1_h, qkv_st, act_st, rows, N, K, ep, st);
    if constexpr (HD == 16) {
      if (bb->bn1 == 64) return launch_gemm_ws<64, EpiLinear1Ws<16>>(ta, bw.tm_w1, &bw.tm_w1_h, qkv_st, act_st, rows, N, K, ep, st);
    }
  }
  return 1;
}

static int launch_linear2_ws(const lamslide_backbone* bb, const CUtensorMap& ta, const BlockWeights& bw, const CUtensorMap& h_red,
                             int rows, const EpiLinear2Ws::Params& ep, cudaStream_t st) {
  const int N = bb->H, K = bb->H + bb->M;
  switch (bb->bn2) {
    case 192: return launch_gemm_ws<192, EpiLinear2Ws>(ta, bw.tm_w2, &bw.tm_w2_h, h_red, h_red, rows, N, K, ep, st);
    case 128: return launch_gemm_ws<128, EpiLinear2Ws>(ta, bw.tm_w2, &bw.tm_w2_h, h_red, h_red, rows, N, K, ep, st);
    case 64: return launch_gemm_ws<64, EpiLinear2Ws>(ta, bw.tm_w2, &bw.tm_w2_h, h_red, h_red, rows, N, K, ep, st);
    default: return 1;
  }
}

// fused MLP half of linear1 + GELU + linear2 + gated residual (mlp_fused.cuh).  Returns 1 when the shape is not covered.
// ring depths of the fused MLP kernel (mlp_fused.cuh): ring 1 (16 KB stages) up to 4 deep, ring 2 what is left (3..6)
static bool fused_mlp_stages(int H, int M, int* s1, int* s2) {
  if (H % 128 != 0 || H > 384 || M % 128 != 0 || M <= 0) return false;
  for (int a = 4; a >= 2; --a)
    for (int b = 6; b >= 3; --b) {
      const FusedMlpSmem plan = fused_mlp_smem(H, M, a, b);
      if (plan.total <= 232448) {
        *s1 = a, *s2 = b;
        return true;
      }
    }
  return false;
}
static bool fused_mlp_ok(int H, int M) {
  int a, b;
  return fused_mlp_stages(H, M, &a, &b);
}
static int launch_fused_mlp(const CUtensorMap& tm_u, const CUtensorMap& tm_attn, const CUtensorMap& tm_w1u, const CUtensorMap& tm_w2u,
                            const CUtensorMap& tm_h, const CUtensorMap& tm_u_st, int rows, const FusedMlpParams& p_in, cudaStream_t st) {
  int s1 = 0, s2 = 0;
  if (!fused_mlp_stages(p_in.H, p_in.M, &s1, &s2)) return 1;
  FusedMlpParams p = p_in;
#ifdef LAMSLIDE_DEBUG_KNOBS
  if (const char* e = getenv("LAMSLIDE_FUSED_DEBUG")) p.debug = atoi(e);  // profiling aids (mlp_fused.cuh)
  if (const char* e = getenv("LAMSLIDE_FUSED_TRACE")) p.trace = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));  // device buffer, 3 x 4096 x 8 B
  if (const char* e = getenv("LAMSLIDE_FUSED_STAGES")) s1 = std::max(2, std::min(s1, atoi(e))), s2 = std::max(2, std::min(s2, atoi(e)));
#endif
  const FusedMlpSmem plan = fused_mlp_smem(p.H, p.M, s1, s2);
  TRY(ensure_dynamic_smem((const void*)mlp_fused_kernel, 232448));
  const int mblocks = cdiv(rows, kBlockM);
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  int grid_cap = num_sms();
  grid_cap = std::max(2, std::min(grid_cap, env_int("LAMSLIDE_FUSED_GRID", grid_cap)));  // profiling aid
  cfg.gridDim = dim3(std::min(grid_cap / 2 * 2, cdiv(mblocks, 2) * 2));
  cfg.blockDim = dim3(kFusedThreads);
  cfg.dynamicSmemBytes = plan.total;
  cfg.stream = st;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  CUDA_TRY(cudaLaunchKernelEx(&cfg, mlp_fused_kernel, tm_u, tm_attn, tm_w1u, tm_w2u, tm_h, tm_u_st, mblocks, s1, s2, p));
  LAUNCH_CHECK();
  return 0;
}

static int launch_plain(int bn, const CUtensorMap& ta, const CUtensorMap& tb, int rows, int N, int K, const EpiPlain::Params& ep,
                        cudaStream_t st) {
  switch (bn) {
    case 16: return launch_gemm<16, EpiPlain>(ta, tb, rows, N, K, ep, st);
    case 32: return launch_gemm<32, EpiPlain>(ta, tb, rows, N, K, ep, st);
    case 48: return launch_gemm<48, EpiPlain>(ta, tb, rows, N, K, ep, st);
    case 64: return launch_gemm<64, EpiPlain>(ta, tb, rows, N, K, ep, st);
    case 96: return launch_gemm<96, EpiPlain>(ta, tb, rows, N, K, ep, st);
    case 128: return launch_gemm<128, EpiPlain>(ta, tb, rows, N, K, ep, st);
    case 192: return launch_gemm<192, EpiPlain>(ta, tb, rows, N, K, ep, st);
    case 256: return launch_gemm<256, EpiPlain>(ta, tb, rows, N, K, ep, st);
    default: return fail(LAMSLIDE_ERR_INVALID, "unsupported GEMM tile width %d", bn);
  }
}

// mode: 0 = auto, 1 = force the streaming (online-max) flash kernel, 2 = force the mma.sync whole-sequence kernel, 3 + 4 v = force the
// tcgen05 kernel (v = 0: shipped exponential mix; v = 1, 2, 3: 0, 2 or 4 of every 8 exponential pairs on the FMA-pipe polynomial; v = 4: profiling
// aid, no exponentials — wrong results; v = 5 / 6: v = 0 / 4 with the cycle trace of lamslide_debug_attention_trace).
// logit_bound: upper bound of |q.k| in the exp2 domain (<= 0: unknown) — the tcgen05 and whole-sequence kernels skip the running maximum.
constexpr float kSeqKernelMaxLogit = 64.f;
static long long* g_atc_trace = nullptr;  // profiling aid (lamslide_debug_attention_trace): device buffer of 32 x 8 bytes per CTA
constexpr int kAtcPolyDefault = 3;  // measured (scripts/attn_ubench.cu): 3 of 8 pairs on the polynomial keeps MUFU and the FMA pipe level
template <int HD>
static int launch_attention(const __nv_bfloat16* qkv, __nv_bfloat16* out, int H, int ldo, int heads, const SeqMap& sm, int n_seq,
                            int mode, float logit_bound, cudaStream_t st) {
  const bool force_flash = mode == 1;
  const bool bounded = logit_bound > 0.f && logit_bound <= kSeqKernelMaxLogit;
  const size_t seq_smem = AttnSeqCfg<HD>::smem_bytes(sm.S);
  static const bool legacy_attn = env_flag("LAMSLIDE_LEGACY_ATTN");   // debug builds: streaming flash kernel everywhere
  static const bool no_tc = env_flag("LAMSLIDE_ATTN_NO_TC");          // debug builds: mma.sync whole-sequence kernel instead of tcgen05
  const bool seq_ok = seq_smem <= 232448 - 1024 && (mode == 2 || (!legacy_attn && bounded));
  if (mode == 2 && !seq_ok) return fail(LAMSLIDE_ERR_INVALID, "sequence length %d too long for the whole-sequence attention kernel", sm.S);
  // tcgen05 kernel (attn_tc.cuh): long sequences with bounded logits whose double-buffered K / V images fit in shared memory
  const size_t tc_smem = AtcCfg<HD>::smem_bytes(sm.S);
  const bool tc_forced = (mode & 3) == 3;
  const bool tc_ok = tc_smem <= 232448 && (tc_forced || (mode == 0 && !legacy_attn && !no_tc && sm.S >= 384 && bounded));
  if (tc_forced && !tc_ok) return fail(LAMSLIDE_ERR_INVALID, "sequence length %d too long for the tcgen05 attention kernel", sm.S);
  if (tc_ok) {
    const int variant = tc_forced ? (mode >> 2) : env_int("LAMSLIDE_ATTN_TC_VARIANT", 0);
    if (variant >= 7 && variant <= 10) {  // three tile groups, P in place (attn_tc3.cuh): 7 shipped mix, 8 no exponentials, 9 / 10: 4 / 2 of 8 pairs polynomial
      // (experiment kernel: validated for >= 3 query tiles per sequence; with two tiles and several items per CTA it faults)
      if (sm.S <= 256) return fail(LAMSLIDE_ERR_INVALID, "the three-group attention variant needs sequences longer than 256");
      void (*k3)(const __nv_bfloat16*, __nv_bfloat16*, int, int, SeqMap, int, int) =
          variant == 8 ? attn_tc3_kernel<HD, -1> : variant == 9 ? attn_tc3_kernel<HD, 4> : variant == 10 ? attn_tc3_kernel<HD, 2> : attn_tc3_kernel<HD, kAtcPolyDefault>;
      TRY(ensure_dynamic_smem((const void*)k3, 232448, true));
      const int n_items = n_seq * heads;
      k3<<<(unsigned)std::min(num_sms(), n_items), kAtc3Threads, Atc3Cfg<HD>::smem_bytes(sm.S), st>>>(qkv, out, H, ldo, sm, heads, n_items);
      COUNT_KERNEL("attn_tc3");
      LAUNCH_CHECK();
      return 0;
    }
    void (*kern)(const __nv_bfloat16*, __nv_bfloat16*, int, int, SeqMap, int, int, long long*) =
        variant == 1 ? attn_tc_kernel<HD, 0> : variant == 2 ? attn_tc_kernel<HD, 2> : variant == 3 ? attn_tc_kernel<HD, 4>
        : variant == 4 ? attn_tc_kernel<HD, -1> : variant == 5 ? attn_tc_kernel<HD, kAtcPolyDefault, true>
        : variant == 6 ? attn_tc_kernel<HD, -1, true> : attn_tc_kernel<HD, kAtcPolyDefault>;
    TRY(ensure_dynamic_smem((const void*)kern, 232448, true));
    const int n_items = n_seq * heads;
    kern<<<(unsigned)std::min(num_sms(), n_items), kAtcThreads, tc_smem, st>>>(qkv, out, H, ldo, sm, heads, n_items, g_atc_trace);
    COUNT_KERNEL("attn_tc");
  } else if ((sm.S > 32 && seq_ok && !force_flash) || mode == 2) {
    // mma.sync whole-sequence kernel: shapes the tcgen05 kernel does not take (short sequences: MD17 L = 192; K / V images too large
    // for double buffering).  16 query rows per warp and 16 warps per CTA.  Measured on B200 (4AA temporal attention): 661 us; it is
    // bound by the 14 HMMA + 16 MUFU.EX2 per warp and 16-key block (scripts/hmma_bench.cu), exponentials on the FMA pipe made it slower.
    void (*kern)(const __nv_bfloat16*, __nv_bfloat16*, int, int, SeqMap, int) = attn_seq_kernel<HD, 0x0000u, 1>;
    TRY(ensure_dynamic_smem((const void*)kern, 232448 - 1024));
    kern<<<(unsigned)(n_seq * heads), 512, seq_smem, st>>>(qkv, out, H, ldo, sm, heads);
    COUNT_KERNEL("attn_seq");
  } else if (sm.S <= 32 && !force_flash) {
    // contiguous sequences whose q|k|v rows fit 8 warps x 6 KB of shared memory: warp per sequence (attn_rows_kernel)
    static const bool legacy_small = env_flag("LAMSLIDE_LEGACY_SMALL_ATTN");
    static const bool no_short_mma = env_flag("LAMSLIDE_NO_SHORT_MMA_ATTN");
    const bool contiguous = sm.seq_stride == 1 && sm.inner == 1 && sm.outer_stride == sm.S;
    const size_t rows_smem = (size_t)8 * sm.S * 3 * H * 2;
    if (!legacy_small && contiguous && rows_smem <= 48 * 1024 && ldo % 8 == 0 && H % 8 == 0) {
      attn_rows_kernel<HD><<<(unsigned)cdiv((long long)n_seq, 8), 256, rows_smem, st>>>(qkv, out, H, ldo, heads, sm.S, n_seq);
      COUNT_KERNEL("attn_rows");
    } else if (!legacy_small && !no_short_mma && ldo % 8 == 0 && H % 8 == 0) {
      // any token stride: warp per (sequence, head) on the warp-level tensor path (attn_short_mma_kernel)
      constexpr int kPitch = ((HD == 24 ? 32 : HD) + 8);
      const size_t smem = (size_t)8 * 3 * 32 * kPitch * 2;
      const long long n_work = (long long)n_seq * heads;
      TRY(ensure_dynamic_smem((const void*)attn_short_mma_kernel<HD>, 64 * 1024));
      attn_short_mma_kernel<HD><<<(unsigned)cdiv(n_work, 8), 256, smem, st>>>(qkv, out, H, ldo, heads, sm, n_work);
      COUNT_KERNEL("attn_short");
    } else {
      long long items = (long long)n_seq * sm.S * heads;
      attn_small_kernel<HD><<<cdiv(items, 256), 256, 0, st>>>(qkv, out, H, ldo, heads, sm, items);
      COUNT_KERNEL("attn_small");
    }
  } else {
    int nqt = cdiv(sm.S, 128);
    dim3 grid((unsigned)(n_seq * nqt), heads);
    attn_flash_kernel<HD><<<grid, 256, 0, st>>>(qkv, out, H, ldo, sm, nqt);
    COUNT_KERNEL("attn_flash");
  }
  LAUNCH_CHECK();
  return 0;
}

static int attention_dispatch(const __nv_bfloat16* qkv, __nv_bfloat16* out, int H, int ldo, int heads, int hd, const SeqMap& sm,
                              int n_seq, int mode, float logit_bound, cudaStream_t st) {
  switch (hd) {
    case 16: return launch_attention<16>(qkv, out, H, ldo, heads, sm, n_seq, mode, logit_bound, st);
    case 24: return launch_attention<24>(qkv, out, H, ldo, heads, sm, n_seq, mode, logit_bound, st);
    case 32: return launch_attention<32>(qkv, out, H, ldo, heads, sm, n_seq, mode, logit_bound, st);
    default: return fail(LAMSLIDE_ERR_INVALID, "head_dim %d unsupported", hd);
  }
}

static int ln_modulate_split3(const lamslide_backbone* bb, const float* h, __nv_bfloat16* u3, const float* shift, const float* scale,
                              int rows, int rows_per_sample, cudaStream_t st) {
  dim3 grid(cdiv(rows, 8));
  switch (bb->H / 128) {
    case 1: ln_modulate_kernel<1, true><<<grid, 256, 0, st>>>(h, u3, shift, scale, bb->mod_width, rows, rows_per_sample); break;
    case 2: ln_modulate_kernel<2, true><<<grid, 256, 0, st>>>(h, u3, shift, scale, bb->mod_width, rows, rows_per_sample); break;
    case 3: ln_modulate_kernel<3, true><<<grid, 256, 0, st>>>(h, u3, shift, scale, bb->mod_width, rows, rows_per_sample); break;
    case 4: ln_modulate_kernel<4, true><<<grid, 256, 0, st>>>(h, u3, shift, scale, bb->mod_width, rows, rows_per_sample); break;
    default: return fail(LAMSLIDE_ERR_INVALID, "hidden_size %d unsupported", bb->H);
  }
  LAUNCH_CHECK();
  return 0;
}

static int ln_modulate(const lamslide_backbone* bb, const float* h, __nv_bfloat16* u, const float* shift, const float* scale, int rows,
                       int rows_per_sample, cudaStream_t st) {
  dim3 grid(cdiv(rows, 8));
  switch (bb->H / 128) {
    case 1: ln_modulate_kernel<1><<<grid, 256, 0, st>>>(h, u, shift, scale, bb->mod_width, rows, rows_per_sample); break;
    case 2: ln_modulate_kernel<2><<<grid, 256, 0, st>>>(h, u, shift, scale, bb->mod_width, rows, rows_per_sample); break;
    case 3: ln_modulate_kernel<3><<<grid, 256, 0, st>>>(h, u, shift, scale, bb->mod_width, rows, rows_per_sample); break;
    case 4: ln_modulate_kernel<4><<<grid, 256, 0, st>>>(h, u, shift, scale, bb->mod_width, rows, rows_per_sample); break;
    default: return fail(LAMSLIDE_ERR_INVALID, "hidden_size %d unsupported", bb->H);
  }
  LAUNCH_CHECK();
  return 0;
}

// ---- per-sample vector path (mmdit.py:93-124, 184-197; latent_si_v31.py:176-186) for n evaluations at once: rows r = (evaluation,
// sample), t_all [n * B].  e = timestep_embedding(t);  hid = silu(time_in.in_layer(e));  svec = silu(time_in.out_layer(hid) [+ vec_in(y)
// of the row's sample]);  mod = [every modulation.lin | adaLN_modulation.1](svec)  ->  ws.mod [n * B, depth * 6H + 2H].
// fp32 through the first stage's linear kernels (3xTF32 tensor-core GEMM when the shapes allow: fp32-accurate).
struct LinW;
struct FsLN;
static int fs_linear(const LinW& L, const float* X, int ldx, float* Y, int ldy, long long rows, int act, const float* res, int ldr,
                     const float* rowadd, int period, int ldra, cudaStream_t st, const FsLN* ln = nullptr,
                     const long long* res_idx = nullptr);
struct LinW {
  float *w = nullptr, *b = nullptr;
  int out = 0, in = 0;
  // tcgen05 3xTF32 path (linear_tc5.cuh): TF32 head / remainder of the weight and their TMA descriptors; bn = 0: not prepared
  float *w_hi = nullptr, *w_lo = nullptr;
  CUtensorMap tm_hi, tm_lo;
  int bn = 0;
};
static int vector_path(const lamslide_backbone* bb, const BackboneWorkspace& w, const float* t_all, const float* y, int B, int n,
                       cudaStream_t st) {
  ProfScope ps(PC_VEC, st);
  const int H = bb->H, R = B * n;
  if (n > w.n_vec) return fail(LAMSLIDE_ERR_WORKSPACE, "vector path: %d evaluations, workspace planned for %d", n, w.n_vec);
  timestep_embed_kernel<<<cdiv(R * 128, 256), 256, 0, st>>>(t_all, w.e, R);
  LAUNCH_CHECK();
  LinW l;
  const float* yv = nullptr;
  if (y) {
    const int V = bb->cfg.vec_in_dim;
    l.w = bb->vec_w1, l.b = bb->vec_b1, l.out = H, l.in = V;
    TRY(fs_linear(l, y, V, w.yhid, H, B, 2, nullptr, 0, nullptr, 0, 0, st));
    l.w = bb->vec_w2, l.b = bb->vec_b2, l.out = H, l.in = H;
    TRY(fs_linear(l, w.yhid, H, w.yvec, H, B, 0, nullptr, 0, nullptr, 0, 0, st));
    yv = w.yvec;
  }
  l.w = bb->time_w1, l.b = bb->time_b1, l.out = H, l.in = 256;
  TRY(fs_linear(l, w.e, 256, w.hid, H, R, 2, nullptr, 0, nullptr, 0, 0, st));
  l.w = bb->time_w2, l.b = bb->time_b2, l.out = H, l.in = H;
  TRY(fs_linear(l, w.hid, H, w.svec, H, R, 2, nullptr, 0, yv, B, H, st));
  l.w = bb->mod_w, l.b = bb->mod_b, l.out = bb->mod_width, l.in = H;
  TRY(fs_linear(l, w.svec, H, w.mod, bb->mod_width, R, 0, nullptr, 0, nullptr, 0, 0, st));
  return 0;
}

struct ForwardCtx {
  BackboneWorkspace ws;
  CUtensorMap tm_u, tm_act, tm_u3;
  CUtensorMap tm_u_st;                          // bf16 store of the fused MLP drain's LN + modulate output (16-column x 32-row boxes)
  CUtensorMap tm_h_red;                         // f32 reduce-add of the linear2 epilogue (16-column x 32-row boxes)
  CUtensorMap tm_qkv_st, tm_act_st;             // linear1 epilogue: dense {2 hd, 32} bf16 boxes (hd = 24: paired heads, 96-byte rows)
  CUtensorMap tm_emb_a;                         // [n, 6D] split input-embedding operand (lives in the act buffer)
};

static int forward_prepare(lamslide_backbone* bb, ForwardCtx& fc, int B, int T, int L, void* workspace, size_t workspace_bytes,
                           cudaStream_t st, int n_vec = 1) {
  if (B <= 0 || T <= 0 || L <= 0) return fail(LAMSLIDE_ERR_INVALID, "B, T, L must be positive");
  fc.ws = plan_workspace(bb, workspace, B, T, L, n_vec);
  if (!workspace || workspace_bytes < fc.ws.bytes)
    return fail(LAMSLIDE_ERR_WORKSPACE, "workspace too small: %zu < %zu bytes", workspace_bytes, fc.ws.bytes);
  if (((uintptr_t)workspace & 1023) != 0) return fail(LAMSLIDE_ERR_INVALID, "workspace must be 1024-byte aligned");
  const long long n = (long long)B * T * L;
  if (n > 0x7fffffffLL / 4) return fail(LAMSLIDE_ERR_INVALID, "too many tokens (%lld)", n);
  TRY(make_tmap(&fc.tm_u, fc.ws.u, (uint64_t)n, bb->H, kBlockM));
  TRY(make_tmap(&fc.tm_act, fc.ws.act, (uint64_t)n, bb->H + bb->M, kBlockM));
  TRY(make_tmap(&fc.tm_u3, fc.ws.qkv, (uint64_t)n, 3 * bb->H, kBlockM));  // head input [hi | lo | hi] reuses the qkv buffer
  TRY(make_tmap_ex(&fc.tm_qkv_st, fc.ws.qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)n, 3 * bb->H, 2 * bb->hd, 32, CU_TENSOR_MAP_SWIZZLE_NONE));
  TRY(make_tmap_ex(&fc.tm_act_st, fc.ws.act, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)n, bb->H + bb->M, 2 * bb->hd, 32,
                   CU_TENSOR_MAP_SWIZZLE_NONE));
  if (bb->w_emb) TRY(make_tmap(&fc.tm_emb_a, fc.ws.act, (uint64_t)n, 6 * bb->D, kBlockM));
  TRY(make_tmap_ex(&fc.tm_h_red, fc.ws.h, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (uint64_t)n, bb->H, 16, 32, CU_TENSOR_MAP_SWIZZLE_64B));
  TRY(make_tmap_ex(&fc.tm_u_st, fc.ws.u, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)n, bb->H, 16, 32, CU_TENSOR_MAP_SWIZZLE_32B));
  const int half = bb->hd / 2;
  rope_table_kernel<<<cdiv(L * half, 256), 256, 0, st>>>(fc.ws.cos_s, fc.ws.sin_s, L, half, (double)bb->cfg.theta);
  LAUNCH_CHECK();
  rope_table_kernel<<<cdiv(T * half, 256), 256, 0, st>>>(fc.ws.cos_t, fc.ws.sin_t, T, half, (double)bb->cfg.theta);
  LAUNCH_CHECK();
  return 0;
}

// LatentSIV3.forward on prepared workspace; result (net output) in ws.m_out, or `out` when given.
static int forward_run(lamslide_backbone* bb, ForwardCtx& fc, const float* x, const float* t, const float* x_cond,
                       const int64_t* mask, const float* y, float* out, int B, int T, int L, cudaStream_t st,
                       const float* mod_ready = nullptr) {
  const int H = bb->H, M = bb->M, D = bb->D, hd = bb->hd, heads = bb->heads;
  const int n = B * T * L;
  BackboneWorkspace& w = fc.ws;
  if (y && bb->cfg.vec_in_dim <= 0) return fail(LAMSLIDE_ERR_INVALID, "y given but the model has no vec_in");

  // 1. h = x_in(x) + cond_to_emb(x_cond) + mask_to_emb(mask) [+ layer_norm]
  {
    ProfScope ps(PC_EMBED, st);
    static const bool legacy_embed = env_flag("LAMSLIDE_LEGACY_EMBED");
    int re = 1;
    if (bb->w_emb && !legacy_embed) {
      const long long nthr = (long long)n * 2 * (D / 4);
      split3_embed_kernel<<<cdiv(nthr, 256), 256, 0, st>>>((const float4*)x, (const float4*)x_cond, w.act, n, D);
      LAUNCH_CHECK();
      EpiEmbedWs::Params ee{bb->b_in, bb->emask, (const long long*)mask, w.h, H, n};
      switch (bb->bn_emb) {
        case 192: re = launch_gemm_ws<192, EpiEmbedWs>(fc.tm_emb_a, bb->tm_wemb, &bb->tm_wemb_h, fc.tm_emb_a, fc.tm_emb_a, n, H, 6 * D, ee, st); break;
        case 128: re = launch_gemm_ws<128, EpiEmbedWs>(fc.tm_emb_a, bb->tm_wemb, &bb->tm_wemb_h, fc.tm_emb_a, fc.tm_emb_a, n, H, 6 * D, ee, st); break;
        case 64: re = launch_gemm_ws<64, EpiEmbedWs>(fc.tm_emb_a, bb->tm_wemb, &bb->tm_wemb_h, fc.tm_emb_a, fc.tm_emb_a, n, H, 6 * D, ee, st); break;
        default: break;
      }
      if (re < 0) return re;
    }
    size_t smem = (size_t)2 * D * 36 * sizeof(float);
    if (re == 0) {
    } else if (H <= 256) {
      embed_in_kernel<1><<<cdiv(n, 32), 256, smem, st>>>(x, x_cond, (const long long*)mask, bb->wt_in, bb->b_in, bb->emask, w.h, n, D, H);
    } else {
      embed_in_kernel<2><<<cdiv(n, 32), 256, smem, st>>>(x, x_cond, (const long long*)mask, bb->wt_in, bb->b_in, bb->emask, w.h, n, D, H);
    }
    LAUNCH_CHECK();
    if (bb->cfg.normalize) {
      layernorm_rows_kernel<<<cdiv(n, 8), 256, 0, st>>>(w.h, n, H, 1e-5f);
      LAUNCH_CHECK();
    }
  }
  // 2. per-sample vectors: all modulations of all layers (vector_path), unless the caller computed them for a whole time grid
  const float* mod = mod_ready;
  if (!mod) {
    TRY(vector_path(bb, w, t, y, B, 1, st));
    mod = w.mod;
  }
  // 3. layers
  static const bool legacy_gemm = env_flag("LAMSLIDE_LEGACY_GEMM");  // A/B switch: one-tile-per-CTA GEMM kernels
  static const bool no_fused = env_flag("LAMSLIDE_NO_FUSED_MLP");    // A/B switch: separate linear1 (full) + linear2 kernels
  // fused MLP path: linear1 computes q | k | v only; the MLP half, the GELU and linear2 run in mlp_fused_kernel
  const bool fused = !legacy_gemm && !no_fused && fused_mlp_ok(H, M) && (3 * H) % bb->bn1 == 0 &&
                     (bb->bn1 == 192 || bb->bn1 == 128 || (bb->bn1 == 64 && hd == 16));
  const float q_premul = (float)(1.4426950408889634 / std::sqrt((double)hd));
  // spatial attention over L <= 8 latents inside the linear1 epilogue (needs the fused-MLP path: linear1 computes q | k | v only)
  static const bool no_fuse_spatial = env_flag("LAMSLIDE_NO_FUSED_SPATIAL_ATTN");
  const bool fuse_spatial = fused && !no_fuse_spatial;
  static const bool no_fuse_ln = env_flag("LAMSLIDE_NO_FUSED_LN");
  const bool fuse_ln = fused && !no_fuse_ln;
  bool u_ready = false;  // w.u already holds LN + modulate of the current h for this block (written by the previous block's drain)
  for (int i = 0; i < bb->depth; ++i) {
    for (int s = 0; s < 2; ++s) {
      const BlockWeights& bw = bb->blocks[2 * i + s];
      const float* modl = mod + (size_t)i * 6 * H + (size_t)s * 3 * H;  // shift | scale | gate
      if (!u_ready) {
        ProfScope ps(PC_LNMOD, st);
        TRY(ln_modulate(bb, w.h, w.u, modl, modl + H, n, T * L, st));
      }
      u_ready = false;
      SeqMap sm;
      int n_seq;
      const float *cs, *sn;
      int pos_div, pos_mod;
      if (s == 0) {  // spatial: sequences (b, t) over l
        sm = SeqMap{L, 1, L, 0, 1};
        n_seq = B * T;
        cs = w.cos_s, sn = w.sin_s, pos_div = 1, pos_mod = L;
      } else {  // temporal: sequences (b, l) over t
        sm = SeqMap{T, L, T * L, 1, L};
        n_seq = B * L;
        cs = w.cos_t, sn = w.sin_t, pos_div = L, pos_mod = T;
      }
#define L1_ATTN(HD_, AL_)                                                                                              \
  {                                                                                                                      \
    typename EpiLinear1Ws<HD_, AL_>::Params epa{bw.b1, {}, cs, sn, w.qkv, w.act, H, 0, n, pos_div, pos_mod, 0, H + M};  \
    for (int j = 0; j < HD_; ++j) epa.gam[0][j] = bw.gq_h[j] * q_premul, epa.gam[1][j] = bw.gk_h[j];                     \
    r1 = launch_linear1_attn<HD_, AL_>(bb, fc.tm_u, bw, fc.tm_qkv_st, fc.tm_act_st, n, epa, st);                         \
    if (r1 < 0) return r1;                                                                                               \
    attn_done = r1 == 0;                                                                                                 \
  }
#define L1_PARAMS(HD_)                                                                                                 \
  ProfScope ps(PC_LINEAR1, st);                                                                                          \
  int r1 = 1;                                                                                                            \
  if (fuse_spatial && s == 0 && L == 2) L1_ATTN(HD_, 2)                                                                  \
  else if (fuse_spatial && s == 0 && L == 4) L1_ATTN(HD_, 4)                                                             \
  else if (fuse_spatial && s == 0 && L == 8 && HD_ >= 16) L1_ATTN(HD_, 8)                                                \
  if (r1 == 1 && !legacy_gemm) {                                                                                         \
    typename EpiLinear1Ws<HD_>::Params epw{bw.b1, {}, cs, sn, w.qkv, w.act, H, fused ? 0 : M, n, pos_div, pos_mod, 0, H + M};             \
    for (int j = 0; j < HD_; ++j) epw.gam[0][j] = bw.gq_h[j] * q_premul, epw.gam[1][j] = bw.gk_h[j];                                         \
    r1 = launch_linear1_ws<HD_>(bb, fc.tm_u, bw, fc.tm_qkv_st, fc.tm_act_st, n, epw, st);                                \
    if (r1 < 0) return r1;                                                                                               \
  }                                                                                                                      \
  if (r1 == 1) {                                                                                                         \
    typename EpiLinear1<HD_>::Params ep{bw.b1, bw.gq, bw.gk, cs, sn, w.qkv, w.act, H, M, n, pos_div, pos_mod, q_premul}; \
    TRY(launch_linear1<HD_>(bb, fc.tm_u, bw, n, ep, st));                                                                \
  }
      bool attn_done = false;  // the spatial attention ran inside the linear1 epilogue
      if (hd == 16) {
        L1_PARAMS(16)
      } else if (hd == 24) {
        L1_PARAMS(24)
      } else {
        L1_PARAMS(32)
      }
#undef L1_PARAMS
#undef L1_ATTN
      if (!attn_done) {
        ProfScope ps(s == 0 ? PC_ATTN_SPATIAL : PC_ATTN_TEMPORAL, st);
        TRY(attention_dispatch(w.qkv, w.act, H, H + M, heads, hd, sm, n_seq, 0, bw.logit_bound, st));
      }
      {
        ProfScope ps(PC_LINEAR2, st);
        int r2 = 1;
        if (fused) {
          FusedMlpParams fp{bw.b1 + 3 * H, bw.b2, modl + 2 * H, bb->mod_width, T * L, H, M, n, nullptr, nullptr, nullptr, nullptr, nullptr, 0};
          const bool last_block = i == bb->depth - 1 && s == 1;
          if (fuse_ln && !last_block) {  // the drain also writes the next block's LN + modulate input
            const float* modn = s == 0 ? modl + 3 * H : mod + (size_t)(i + 1) * 6 * H;
            fp.h = w.h, fp.u_out = w.u, fp.ln_shift = modn, fp.ln_scale = modn + H;
          }
          r2 = launch_fused_mlp(fc.tm_u, fc.tm_act, bw.tm_w1_u, bw.tm_w2_u, fc.tm_h_red, fc.tm_u_st, n, fp, st);
          if (r2 < 0) return r2;
          u_ready = r2 == 0 && fp.ln_scale != nullptr;
        } else if (!legacy_gemm) {
          EpiLinear2Ws::Params e2w{bw.b2, modl + 2 * H, bb->mod_width, T * L, H, n};
          r2 = launch_linear2_ws(bb, fc.tm_act, bw, fc.tm_h_red, n, e2w, st);
          if (r2 < 0) return r2;
        }
        if (r2 == 1) {
          EpiLinear2::Params e2{w.h, bw.b2, modl + 2 * H, bb->mod_width, T * L, H, n};
          TRY(launch_linear2(bb, fc.tm_act, bw, n, e2, st));
        }
      }
    }
  }
  // 4. final adaLN + linear
  {
    ProfScope ps(PC_HEAD, st);
    const float* ada = mod + (size_t)bb->depth * 6 * H;  // shift | scale
    TRY(ln_modulate_split3(bb, w.h, w.qkv, ada, ada + H, n, T * L, st));
    EpiPlain::Params ep{out ? out : w.m_out, bb->b_out, D, n};
    TRY(launch_plain(bb->bn_out, fc.tm_u3, bb->tm_wout, n, D, 3 * H, ep, st));
  }
  return 0;
}

static int check_device(const lamslide_backbone* h) {
  int dev = -1;
  CUDA_TRY(cudaGetDevice(&dev));
  if (dev != h->device) return fail(LAMSLIDE_ERR_INVALID, "handle was created on device %d, current device is %d", h->device, dev);
  return 0;
}

extern "C" int lamslide_backbone_forward(lamslide_backbone* h, const float* x, const float* t, const float* x_cond,
                                         const int64_t* x_cond_mask, const float* y, float* out, int32_t B, int32_t T, int32_t L,
                                         void* workspace, size_t workspace_bytes, void* stream) {
  if (!h || !x || !t || !x_cond || !x_cond_mask || !out) return fail(LAMSLIDE_ERR_INVALID, "null argument");
  TRY(check_device(h));
  cudaStream_t st = (cudaStream_t)stream;
  ForwardCtx fc;
  TRY(forward_prepare(h, fc, B, T, L, workspace, workspace_bytes, st));
  return forward_run(h, fc, x, t, x_cond, x_cond_mask, y, out, B, T, L, st);
}

// t_all[s * B + b] = times.t[s]: the time vectors of up to 16 consecutive ODE steps (integrators.py:107-112: th.ones(B) * t)
struct StepTimes {
  float t[16];
};
__global__ void fill_steps_kernel(float* p, StepTimes times, int B, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * n) p[i] = times.t[i / B];
}
__global__ void copy_f4_kernel(const float4* __restrict__ src, float4* __restrict__ dst, long long n4) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) dst[i] = src[i];
}

// torch.linspace(start, end, steps) in fp32 (symmetric evaluation, as ATen's CPU/CUDA kernels do)
static float linspace_f32(float start, float end, int steps, int i) {
  float step = (end - start) / (float)(steps - 1);
  int halfway = steps / 2;
  return i < halfway ? start + step * (float)i : end - step * (float)(steps - i - 1);
}

// v = cm * m + cx * x for every (path, prediction) of Transport.get_drift (transport.py:158-202; path.py ICPlan / GVPCPlan)
static int drift_coefficients(int path_type, int prediction, double t, double* cm, double* cx) {
  if (prediction == 0) {
    *cm = 1.0, *cx = 0.0;
    return 0;
  }
  const double PI = 3.14159265358979323846;
  double alpha, sigma, d_sigma, ratio;
  if (path_type == 1) {  // GVP
    alpha = std::sin(t * PI / 2), sigma = std::cos(t * PI / 2), d_sigma = -PI / 2 * std::sin(t * PI / 2);
    ratio = PI / (2 * std::tan(t * PI / 2));
  } else if (path_type == 0) {  // Linear
    alpha = t, sigma = 1 - t, d_sigma = -1, ratio = 1 / t;
  } else {
    return fail(LAMSLIDE_ERR_INVALID, "path_type %d unsupported (0 Linear, 1 GVP)", path_type);
  }
  const double drift_var = ratio * sigma * sigma - sigma * d_sigma;  // -drift_mean = ratio * x
  switch (prediction) {
    case 1:  // data: score = -(x - alpha m) / sigma^2
      *cm = drift_var * alpha / (sigma * sigma);
      *cx = ratio - drift_var / (sigma * sigma);
      return 0;
    case 2:  // noise: score = -m / sigma
      *cm = -drift_var / sigma;
      *cx = ratio;
      return 0;
    case 3:  // score
      *cm = drift_var;
      *cx = ratio;
      return 0;
    default: return fail(LAMSLIDE_ERR_INVALID, "prediction %d unsupported", prediction);
  }
}

extern "C" int lamslide_ode_sample(lamslide_backbone* h, float* x, const float* x_cond, const int64_t* x_cond_mask, const float* y,
                                   int32_t path_type, int32_t prediction, int32_t num_steps, float* states_out,
                                   float* velocities_out, int32_t B, int32_t T, int32_t L, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  if (!h || !x || !x_cond || !x_cond_mask) return fail(LAMSLIDE_ERR_INVALID, "null argument");
  if (num_steps < 2) return fail(LAMSLIDE_ERR_INVALID, "num_steps must be >= 2");
  TRY(check_device(h));
  cudaStream_t st = (cudaStream_t)stream;
  // the per-sample vector path (timestep embedding -> time_in -> every layer's modulation) depends on the time grid only, which is
  // known in advance: it runs once per chunk of up to 16 steps instead of once per step
  const int chunk = ode_vec_chunk(h, B, num_steps - 1);
  ForwardCtx fc;
  TRY(forward_prepare(h, fc, B, T, L, workspace, workspace_bytes, st, chunk));
  // Transport.check_interval (transport.py:69-101) for the ODE sampler
  float t0 = 0.f, t1 = 1.f;
  if (prediction != 0) t0 = 1e-3f, t1 = 1.f - 1e-3f;
  const long long nel = (long long)B * T * L * h->D;
  const long long n4 = nel / 4;  // D % 16 == 0
  if (states_out) {
    copy_f4_kernel<<<cdiv(n4, 256), 256, 0, st>>>((const float4*)x, (float4*)states_out, n4);
    LAUNCH_CHECK();
  }
  for (int i0 = 0; i0 < num_steps - 1; i0 += chunk) {
    const int n = std::min(chunk, num_steps - 1 - i0);
    StepTimes times;
    for (int s = 0; s < n; ++s) times.t[s] = linspace_f32(t0, t1, num_steps, i0 + s);
    fill_steps_kernel<<<cdiv(B * n, 256), 256, 0, st>>>(fc.ws.tvec, times, B, n);
    LAUNCH_CHECK();
    TRY(vector_path(h, fc.ws, fc.ws.tvec, y, B, n, st));
    for (int s = 0; s < n; ++s) {
      const int i = i0 + s;
      const float ti = times.t[s];
      const float tn = linspace_f32(t0, t1, num_steps, i + 1);
      TRY(forward_run(h, fc, x, nullptr, x_cond, x_cond_mask, y, nullptr, B, T, L, st, fc.ws.mod + (size_t)s * B * h->mod_width));
      double cm, cx;
      TRY(drift_coefficients(path_type, prediction, (double)ti, &cm, &cx));
      ProfScope ps(PC_EULER, st);
      drift_euler_kernel<<<cdiv(n4, 256), 256, 0, st>>>(
          (const float4*)fc.ws.m_out, (float4*)x, velocities_out ? (float4*)(velocities_out + (size_t)i * nel) : nullptr,
          states_out ? (float4*)(states_out + (size_t)(i + 1) * nel) : nullptr, (float)cm, (float)cx, tn - ti, n4);
      LAUNCH_CHECK();
    }
  }
  return 0;
}

extern "C" int lamslide_euler_step(float* x, const float* net_out, int32_t path_type, int32_t prediction, float t, float t_next,
                                   float* velocity_out, int64_t numel, void* stream) {
  if (!x || !net_out) return fail(LAMSLIDE_ERR_INVALID, "null argument");
  if (numel % 4 != 0) return fail(LAMSLIDE_ERR_INVALID, "numel must be a multiple of 4");
  double cm, cx;
  TRY(drift_coefficients(path_type, prediction, (double)t, &cm, &cx));
  drift_euler_kernel<<<cdiv(numel / 4, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)net_out, (float4*)x, (float4*)velocity_out, nullptr,
                                                                            (float)cm, (float)cx, t_next - t, numel / 4);
  LAUNCH_CHECK();
  return 0;
}

extern "C" int lamslide_setup_conditioning(const float* latents, float* x_cond, int64_t* x_cond_mask, int32_t B, int32_t T, int32_t L,
                                           int32_t D, int32_t cond_begin, int32_t cond_end, int32_t mask_cond_mean, void* stream) {
  if (!latents || !x_cond || !x_cond_mask) return fail(LAMSLIDE_ERR_INVALID, "null argument");
  if (cond_end > T) cond_end = T;  // python slice semantics of latents[:, c0:c1]
  if (cond_begin < 0 || cond_begin >= cond_end) return fail(LAMSLIDE_ERR_INVALID, "bad cond_idx [%d, %d)", cond_begin, cond_end);
  ProfScope ps(PC_COND, (cudaStream_t)stream);
  long long tot = (long long)B * L * D;
  conditioning_kernel<<<cdiv(tot, 256), 256, 0, (cudaStream_t)stream>>>(latents, x_cond, (long long*)x_cond_mask, B, T, L, D,
                                                                        cond_begin, cond_end, mask_cond_mean);
  LAUNCH_CHECK();
  return 0;
}

// out = px * x + pm * m + pw * w (m, w nullable; out may alias x): one step piece of the SDE sampler (integrators.py:29-52).
extern "C" int lamslide_lincomb3(float* out, const float* x, const float* m, const float* w, float px, float pm, float pw, int64_t numel,
                                 void* stream) {
  if (!out || !x) return fail(LAMSLIDE_ERR_INVALID, "null argument");
  if (numel <= 0 || numel % 4) return fail(LAMSLIDE_ERR_INVALID, "numel %lld must be a positive multiple of 4", (long long)numel);
  if (((uintptr_t)out | (uintptr_t)x | (uintptr_t)m | (uintptr_t)w) & 15) return fail(LAMSLIDE_ERR_INVALID, "operands must be 16-byte aligned");
  const long long n4 = numel / 4;
  lincomb3_kernel<<<cdiv(n4, 256), 256, 0, (cudaStream_t)stream>>>((float4*)out, (const float4*)x, (const float4*)m, (const float4*)w, px,
                                                                   pm, pw, n4);
  LAUNCH_CHECK();
  return 0;
}

static int fill_lincomb(LincombN* a, const float* const* srcs, const float* coefs, int32_t n) {
  if (!srcs || !coefs || n < 1 || n > 8) return fail(LAMSLIDE_ERR_INVALID, "between 1 and 8 terms (got %d)", n);
  a->n = n;
  for (int j = 0; j < 8; ++j) {
    a->src[j] = j < n ? reinterpret_cast<const float4*>(srcs[j]) : nullptr;
    a->coef[j] = j < n ? coefs[j] : 0.f;
    if (j < n && (!srcs[j] || ((uintptr_t)srcs[j] & 15))) return fail(LAMSLIDE_ERR_INVALID, "term %d: null or not 16-byte aligned", j);
  }
  return 0;
}

// out = sum_j coefs[j] * srcs[j]: the stage / solution / interpolant combinations of the Runge-Kutta ODE integrators.
extern "C" int lamslide_lincomb_n(float* out, const float* const* srcs, const float* coefs, int32_t n, int64_t numel, void* stream) {
  if (!out || ((uintptr_t)out & 15)) return fail(LAMSLIDE_ERR_INVALID, "out: null or not 16-byte aligned");
  if (numel <= 0 || numel % 4) return fail(LAMSLIDE_ERR_INVALID, "numel %lld must be a positive multiple of 4", (long long)numel);
  LincombN a;
  TRY(fill_lincomb(&a, srcs, coefs, n));
  const long long n4 = numel / 4;
  lincomb_n_kernel<<<cdiv(n4, 256), 256, 0, (cudaStream_t)stream>>>((float4*)out, a, n4);
  LAUNCH_CHECK();
  return 0;
}

// *acc (device fp64) += sum_i ((sum_j coefs[j] srcs[j][i]) / (atol + rtol max(|a_i|, |b_i|)))^2
extern "C" int lamslide_rk_error_sumsq(const float* const* srcs, const float* coefs, int32_t n, const float* a, const float* b, double rtol,
                                       double atol, int64_t numel, double* acc, void* stream) {
  if (!a || !b || !acc || (((uintptr_t)a | (uintptr_t)b) & 15)) return fail(LAMSLIDE_ERR_INVALID, "a / b / acc: null or not 16-byte aligned");
  if (numel <= 0 || numel % 4) return fail(LAMSLIDE_ERR_INVALID, "numel %lld must be a positive multiple of 4", (long long)numel);
  LincombN e;
  TRY(fill_lincomb(&e, srcs, coefs, n));
  const long long n4 = numel / 4;
  const int grid = (int)std::min<long long>(cdiv(n4, 256), 4LL * num_sms());
  rk_error_sumsq_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(e, (const float4*)a, (const float4*)b, rtol, atol, n4, acc);
  LAUNCH_CHECK();
  return 0;
}

// K-sample evaluation metrics (second_stage/nba.py:161-238, pedestrian.py:149-226: mode 0; md17.py:139-171: mode 1).
extern "C" int lamslide_ksample_errors(const float* preds, const float* target, float* ades, float* fdes, int32_t K, int32_t num_runs,
                                       int32_t B, int32_t T, int32_t A, int32_t D, int32_t mode, void* stream) {
  if (!preds || !target || !ades || !fdes) return fail(LAMSLIDE_ERR_INVALID, "null argument");
  if (K <= 0 || B <= 0 || T <= 0 || A <= 0 || D <= 0 || (mode != 0 && mode != 1))
    return fail(LAMSLIDE_ERR_INVALID, "bad K-sample shape K %d B %d T %d A %d D %d mode %d", K, B, T, A, D, mode);
  if (mode == 0 && (num_runs <= 0 || num_runs > K)) return fail(LAMSLIDE_ERR_INVALID, "num_runs %d outside [1, K = %d]", num_runs, K);
  const int n_out = mode == 0 ? B * A : B;
  ksample_errors_kernel<<<cdiv(n_out, 128), 128, 0, (cudaStream_t)stream>>>(preds, target, ades, fdes, K, num_runs, B, T, A, D, mode);
  LAUNCH_CHECK();
  return 0;
}

// ================================================================================================ first stage
struct LNW {
  float *w = nullptr, *b = nullptr;
};
struct AttnBlockW {  // CrossAttentionBlock / SelfAttentionBlock (torch_modules.py:189-273)
  bool cross = false;
  int dim = 0, ctx_dim = 0, heads = 0, dh = 0;
  LNW norm, norm_ctx, ff_norm;
  LinW to_q, to_kv, to_qkv, to_out, ff0, ff1;
  float *gq = nullptr, *gk = nullptr;
};

struct lamslide_first_stage {
  lamslide_first_stage_config cfg;
  std::vector<std::string> out_names;
  int device = 0;
  Arena arena;
  int feat_dim = 0;   // columns of the feature matrix = K of net_merge.0 as packed (rounded up to a multiple of 4, zero padded)
  int feat_used = 0;  // columns that carry features
  float* ent_table = nullptr;                           // [num_entities, entity_dim], max_norm pre-applied
  float *tab0 = nullptr, *tab1 = nullptr;               // embedding_res / embed_atom / embed_team ; embed_group
  int tab0_dim = 0, tab1_dim = 0;
  float* sincos = nullptr;                              // [max_res, dim_input]
  float* point_basis = nullptr;                         // [3, nb]
  int point_nb = 0;
  LinW point_mlp, merge0, merge2, enc_mlp0, enc_mlp2, quant, post_quant, query_mlp, extender;
  float* latents = nullptr;                             // [L, D]
  // folded at create (host fp64): what depends only on the learned latents / on the entity id
  float* enc_q0 = nullptr;      // [L, inner]            to_q(LN(latents)) of the first encoder cross block
  float* dec_q_tab = nullptr;   // [num_entities, dq]    query_mlp(entity_embedding)
  float* dec_qq_tab = nullptr;  // [num_entities, inner] to_q(LN(dec_q_tab)) of the decoder output block
  std::vector<AttnBlockW> enc_cross, enc_self, dec_self, dec_cross;
  AttnBlockW out_block;
  std::vector<LinW> head0, head2;
};

// tile width of the tcgen05 linear kernel for an N-wide layer (a function of the layer shape only)
static int tc5_bn(int N) {
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  if (N <= 96) return 96;
  if (N <= 128) return 128;
  if (N <= 192 || N % 192 == 0) return 192;
  return 256;
}
// split a host weight [out, in] into its TF32 head and remainder, upload both and describe them to the TMA
static int prepare_tc5(Arena& A, const float* host_w, LinW* L) {
  if (L->in % 4 != 0 || L->in < 4) return 0;  // TMA row pitch must be a multiple of 16 bytes: such layers stay on the legacy kernels
  const size_t n = (size_t)L->out * L->in;
  std::vector<float> hi(n), lo(n);
  for (size_t i = 0; i < n; ++i) {
    hi[i] = tf32_head(host_w[i]);
    lo[i] = host_w[i] - hi[i];
  }
  TRY(A.upload_f32(hi.data(), n, &L->w_hi));
  TRY(A.upload_f32(lo.data(), n, &L->w_lo));
  L->bn = tc5_bn(L->out);
  TRY(make_tmap_f32(&L->tm_hi, L->w_hi, L->out, L->in, L->in, L->bn));
  TRY(make_tmap_f32(&L->tm_lo, L->w_lo, L->out, L->in, L->in, L->bn));
  return 0;
}

struct FsLoader {
  StateDict sd;
  Arena& A;
  FsLoader(const lamslide_tensor* t, int n, Arena& a) : sd(t, n), A(a) {}
  int lin(const std::string& p, int out, int in, bool bias, LinW* L) {
    const lamslide_tensor* w = sd.get(p + ".weight", {out, in});
    if (!w) return LAMSLIDE_ERR_MISSING;
    TRY(A.upload_f32(w->data, (size_t)out * in, &L->w));
    if (bias) {
      const lamslide_tensor* b = sd.get(p + ".bias", {out});
      if (!b) return LAMSLIDE_ERR_MISSING;
      TRY(A.upload_f32(b->data, out, &L->b));
    }
    L->out = out, L->in = in;
    return prepare_tc5(A, w->data, L);
  }
  int ln(const std::string& p, int dim, LNW* n) {
    const lamslide_tensor* w = sd.get(p + ".weight", {dim});
    const lamslide_tensor* b = sd.get(p + ".bias", {dim});
    if (!w || !b) return LAMSLIDE_ERR_MISSING;
    TRY(A.upload_f32(w->data, dim, &n->w));
    TRY(A.upload_f32(b->data, dim, &n->b));
    return 0;
  }
  int vec(const std::string& name, int n, float** out) {
    const lamslide_tensor* t = sd.get(name, {n});
    if (!t) return LAMSLIDE_ERR_MISSING;
    return A.upload_f32(t->data, n, out);
  }
  // nn.Embedding(max_norm=1): rows with norm > 1 are rescaled to 1/(norm + 1e-7) on lookup (idempotent => apply once)
  int table(const std::string& name, int rows, int dim, bool max_norm, float** out) {
    const lamslide_tensor* t = sd.get(name, {rows, dim});
    if (!t) return LAMSLIDE_ERR_MISSING;
    std::vector<float> tmp(t->data, t->data + (size_t)rows * dim);
    if (max_norm) {
      for (int r = 0; r < rows; ++r) {
        float ss = 0.f;  // fp32 like torch's embedding_renorm_
        for (int j = 0; j < dim; ++j) ss += tmp[(size_t)r * dim + j] * tmp[(size_t)r * dim + j];
        float nrm = std::sqrt(ss);
        if (nrm > 1.0f) {
          float sc = 1.0f / (nrm + 1e-7f);
          for (int j = 0; j < dim; ++j) tmp[(size_t)r * dim + j] *= sc;
        }
      }
    }
    return A.upload_f32(tmp.data(), tmp.size(), out);
  }
  int attn_block(const std::string& p, bool cross, int dim, int ctx_dim, int heads, int dh, bool qk_norm, AttnBlockW* b) {
    b->cross = cross, b->dim = dim, b->ctx_dim = ctx_dim, b->heads = heads, b->dh = dh;
    const int inner = heads * dh;
    TRY(ln(p + "attn.norm", dim, &b->norm));
    if (cross) {
      TRY(ln(p + "attn.norm_context", ctx_dim, &b->norm_ctx));
      TRY(lin(p + "attn.fn.to_q", inner, dim, false, &b->to_q));
      TRY(lin(p + "attn.fn.to_kv", 2 * inner, ctx_dim, false, &b->to_kv));
    } else {
      TRY(lin(p + "attn.fn.to_qkv", 3 * inner, dim, false, &b->to_qkv));
    }
    TRY(lin(p + "attn.fn.to_out", dim, inner, true, &b->to_out));
    if (qk_norm) {
      TRY(vec(p + "attn.fn.norm.query_norm.scale", dh, &b->gq));
      TRY(vec(p + "attn.fn.norm.key_norm.scale", dh, &b->gk));
    }
    TRY(ln(p + "ff.norm", dim, &b->ff_norm));
    TRY(lin(p + "ff.fn.net.0.0", dim, dim, true, &b->ff0));
    TRY(lin(p + "ff.fn.net.1", dim, dim, true, &b->ff1));
    return 0;
  }
};

extern "C" int lamslide_first_stage_create(const lamslide_first_stage_config* cfg, const lamslide_tensor* tensors, int32_t n_tensors,
                                           lamslide_first_stage** out) {
  if (!cfg || !tensors || !out) return fail(LAMSLIDE_ERR_INVALID, "null argument");
  *out = nullptr;
  if (cfg->enc_dim_head_cross != 16 || cfg->enc_dim_head_latent != 16 || cfg->dec_dim_head_cross != 16 || cfg->dec_dim_head_latent != 16)
    return fail(LAMSLIDE_ERR_INVALID, "first-stage attention supports dim_head = 16 only (all shipped configs)");
  if (cfg->n_outputs < 1 || cfg->n_outputs > LAMSLIDE_MAX_OUTPUTS) return fail(LAMSLIDE_ERR_INVALID, "bad n_outputs");
  if (cfg->dim_latent % 4 || cfg->dim_input % 4 || cfg->entity_dim % 4 || cfg->dec_dim_query % 4)
    return fail(LAMSLIDE_ERR_INVALID, "first-stage widths must be multiples of 4");
  auto* fs = new lamslide_first_stage();
  std::unique_ptr<lamslide_first_stage> guard(fs);
  fs->cfg = *cfg;
  CUDA_TRY(cudaGetDevice(&fs->device));
  for (int i = 0; i < cfg->n_outputs; ++i) fs->out_names.push_back(cfg->output_names[i]);
  FsLoader ld(tensors, n_tensors, fs->arena);
  const int Din = cfg->dim_input, D = cfg->dim_latent, E = cfg->entity_dim, C = Din + E, dq = cfg->dec_dim_query;
  const bool qkn = cfg->qk_norm != 0;
  TRY(ld.table("encoder.entity_embedding.embedding.weight", cfg->num_entities, E, true, &fs->ent_table));
  switch (cfg->kind) {
    case LAMSLIDE_FS_PEPTIDE: {
      TRY(ld.table("embedding_res.weight", 20, 64, true, &fs->tab0));
      fs->tab0_dim = 64;
      fs->feat_used = 64 + 42;
      const lamslide_tensor* sc = ld.sd.get("embed_res_pos.embeddings", {cfg->max_res, Din});
      if (!sc) return LAMSLIDE_ERR_MISSING;
      TRY(fs->arena.upload_f32(sc->data, (size_t)cfg->max_res * Din, &fs->sincos));
      break;
    }
    case LAMSLIDE_FS_MD17: {
      TRY(ld.table("embed_atom.weight", cfg->n_atom_types, 64, true, &fs->tab0));
      fs->tab0_dim = 64;
      fs->point_nb = 63;
      const lamslide_tensor* pb = ld.sd.get("embed_pos.basis", {3, 63});
      if (!pb) return LAMSLIDE_ERR_MISSING;
      TRY(fs->arena.upload_f32(pb->data, 3 * 63, &fs->point_basis));
      TRY(ld.lin("embed_pos.mlp", 128, 129, true, &fs->point_mlp));
      fs->feat_used = 64 + 128;
      break;
    }
    case LAMSLIDE_FS_NBA:
      TRY(ld.table("embed_team.weight", 3, 32, false, &fs->tab0));
      TRY(ld.table("embed_group.weight", 2, 32, false, &fs->tab1));
      fs->tab0_dim = fs->tab1_dim = 32;
      fs->feat_used = 2 + 32 + 32;
      break;
    case LAMSLIDE_FS_PEDESTRIAN: fs->feat_used = 2; break;
    default: return fail(LAMSLIDE_ERR_INVALID, "unknown first-stage kind %d", cfg->kind);
  }
  // feature widths that are not a multiple of 4 (peptide 106, nba 66) are zero padded — in the feature matrix and in net_merge.0's K —
  // so that layer takes the vectorised / 3xTF32 kernels (the scalar fallback cost 552 us per 4AA encode); the 2-wide pedestrian
  // input is used in place
  fs->feat_dim = fs->feat_used >= 4 ? (fs->feat_used + 3) / 4 * 4 : fs->feat_used;
  {
    const lamslide_tensor* w = ld.sd.get("net_merge.0.weight", {Din, fs->feat_used});
    const lamslide_tensor* b = ld.sd.get("net_merge.0.bias", {Din});
    if (!w || !b) return LAMSLIDE_ERR_MISSING;
    std::vector<float> wp((size_t)Din * fs->feat_dim, 0.f);
    for (int r = 0; r < Din; ++r) memcpy(&wp[(size_t)r * fs->feat_dim], &w->data[(size_t)r * fs->feat_used], (size_t)fs->feat_used * 4);
    TRY(fs->arena.upload_f32(wp.data(), wp.size(), &fs->merge0.w));
    TRY(fs->arena.upload_f32(b->data, Din, &fs->merge0.b));
    fs->merge0.out = Din, fs->merge0.in = fs->feat_dim;
    TRY(prepare_tc5(fs->arena, wp.data(), &fs->merge0));
  }
  TRY(ld.lin("net_merge.2", Din, Din, true, &fs->merge2));
  TRY(ld.lin("encoder.mlp.0", D, C, true, &fs->enc_mlp0));
  TRY(ld.lin("encoder.mlp.2", C, D, true, &fs->enc_mlp2));
  {
    const lamslide_tensor* lt = ld.sd.get("encoder.latents", {cfg->enc_num_latents, D});
    if (!lt) return LAMSLIDE_ERR_MISSING;
    TRY(fs->arena.upload_f32(lt->data, (size_t)cfg->enc_num_latents * D, &fs->latents));
  }
  fs->enc_cross.resize(cfg->enc_blocks_cross);
  for (int i = 0; i < cfg->enc_blocks_cross; ++i)
    TRY(ld.attn_block("encoder.cross_attn_blocks." + std::to_string(i) + ".", true, D, C, cfg->enc_heads_cross, 16, qkn, &fs->enc_cross[i]));
  fs->enc_self.resize(cfg->enc_blocks_attn);
  for (int i = 0; i < cfg->enc_blocks_attn; ++i)
    TRY(ld.attn_block("encoder.blocks_attn." + std::to_string(i) + ".", false, D, D, cfg->enc_heads_latent, 16, qkn, &fs->enc_self[i]));
  TRY(ld.lin("quant.0", D, D, true, &fs->quant));
  TRY(ld.lin("post_quant.1", D, D, true, &fs->post_quant));
  TRY(ld.lin("decoder.query_mlp.1", dq, E, true, &fs->query_mlp));
  fs->dec_self.resize(cfg->dec_blocks_attn);
  for (int i = 0; i < cfg->dec_blocks_attn; ++i)
    TRY(ld.attn_block("decoder.self_attn_blocks." + std::to_string(i) + ".", false, D, D, cfg->dec_heads_latent, 16, qkn, &fs->dec_self[i]));
  fs->dec_cross.resize(cfg->dec_blocks_cross);
  for (int i = 0; i < cfg->dec_blocks_cross; ++i)
    TRY(ld.attn_block("decoder.cross_attn_blocks." + std::to_string(i) + ".", true, D, dq, cfg->dec_heads_cross, 16, qkn, &fs->dec_cross[i]));
  TRY(ld.attn_block("decoder.output_block.", true, dq, D, cfg->dec_heads_cross, 16, qkn, &fs->out_block));
  fs->head0.resize(cfg->n_outputs);
  fs->head2.resize(cfg->n_outputs);
  for (int i = 0; i < cfg->n_outputs; ++i) {
    std::string p = "decoder.output_layers." + fs->out_names[i];
    TRY(ld.lin(p + ".0", dq, dq, true, &fs->head0[i]));
    TRY(ld.lin(p + ".2", cfg->output_dims[i], dq, true, &fs->head2[i]));
  }
  if (cfg->dec_query_splitter) {
    // Conv1d(D -> D*n, k=1) then "B (D N) L -> B (L N) D": permute output channels (d*n + j) -> (j*D + d) at pack time so the
    // GEMM output [F*L, n*D] IS the [F, L*n, D] key/value token matrix (decoder.py:385-389).
    const int n = cfg->dec_num_split;
    const lamslide_tensor* w = ld.sd.get("decoder.extender.1.weight", {(int64_t)D * n, D});
    const lamslide_tensor* b = ld.sd.get("decoder.extender.1.bias", {(int64_t)D * n});
    if (!w || !b) return LAMSLIDE_ERR_MISSING;
    std::vector<float> wp((size_t)D * n * D), bp((size_t)D * n);
    for (int d = 0; d < D; ++d)
      for (int j = 0; j < n; ++j) {
        memcpy(&wp[((size_t)j * D + d) * D], &w->data[((size_t)d * n + j) * D], (size_t)D * 4);
        bp[(size_t)j * D + d] = b->data[(size_t)d * n + j];
      }
    TRY(fs->arena.upload_f32(wp.data(), wp.size(), &fs->extender.w));
    TRY(fs->arena.upload_f32(bp.data(), bp.size(), &fs->extender.b));
    fs->extender.out = D * n, fs->extender.in = D;
    TRY(prepare_tc5(fs->arena, wp.data(), &fs->extender));
  }
  // ---- constant folding (host, fp64): queries that depend only on the learned latents (encoder.py:39: the latents are the same
  // for every frame) or only on the entity id (decoder.py:83-86) become small tables, so that the per-frame work starts at the
  // attention itself
  static const bool no_fold = env_flag("LAMSLIDE_FS_NO_FOLD");
  if (!no_fold) {
    auto layer_norm = [](std::vector<double>& x, int rows, int dim, const float* w, const float* b) {
      for (int r = 0; r < rows; ++r) {
        double m = 0, v = 0;
        for (int j = 0; j < dim; ++j) m += x[(size_t)r * dim + j];
        m /= dim;
        for (int j = 0; j < dim; ++j) v += (x[(size_t)r * dim + j] - m) * (x[(size_t)r * dim + j] - m);
        const double rstd = 1.0 / std::sqrt(v / dim + 1e-5);
        for (int j = 0; j < dim; ++j) x[(size_t)r * dim + j] = (x[(size_t)r * dim + j] - m) * rstd * w[j] + b[j];
      }
    };
    auto linear = [](const std::vector<double>& x, int rows, int in, const float* w, const float* b, int out) {
      std::vector<double> y((size_t)rows * out);
      for (int r = 0; r < rows; ++r)
        for (int o = 0; o < out; ++o) {
          double acc = b ? b[o] : 0.0;
          for (int k = 0; k < in; ++k) acc += x[(size_t)r * in + k] * w[(size_t)o * in + k];
          y[(size_t)r * out + o] = acc;
        }
      return y;
    };
    auto upload = [&](const std::vector<double>& x, float** dst) {
      std::vector<float> f(x.begin(), x.end());
      return fs->arena.upload_f32(f.data(), f.size(), dst);
    };
    if (cfg->enc_blocks_cross > 0) {
      const int Lq = cfg->enc_num_latents, inner = cfg->enc_heads_cross * 16;
      const std::string p = "encoder.cross_attn_blocks.0.attn.";
      const lamslide_tensor *lt = ld.sd.get("encoder.latents", {Lq, D}), *nw = ld.sd.get(p + "norm.weight", {D}),
                            *nb = ld.sd.get(p + "norm.bias", {D}), *wq = ld.sd.get(p + "fn.to_q.weight", {inner, D});
      if (!lt || !nw || !nb || !wq) return LAMSLIDE_ERR_MISSING;
      std::vector<double> x(lt->data, lt->data + (size_t)Lq * D);
      layer_norm(x, Lq, D, nw->data, nb->data);
      TRY(upload(linear(x, Lq, D, wq->data, nullptr, inner), &fs->enc_q0));
    }
    {
      const int ne = cfg->num_entities, inner = cfg->dec_heads_cross * 16;
      const std::string p = "decoder.output_block.attn.";
      const lamslide_tensor *et = ld.sd.get("encoder.entity_embedding.embedding.weight", {ne, E}),
                            *qw = ld.sd.get("decoder.query_mlp.1.weight", {dq, E}), *qb = ld.sd.get("decoder.query_mlp.1.bias", {dq}),
                            *nw = ld.sd.get(p + "norm.weight", {dq}), *nb = ld.sd.get(p + "norm.bias", {dq}),
                            *wq = ld.sd.get(p + "fn.to_q.weight", {inner, dq});
      if (!et || !qw || !qb || !nw || !nb || !wq) return LAMSLIDE_ERR_MISSING;
      std::vector<double> e((size_t)ne * E);
      for (int r = 0; r < ne; ++r) {  // nn.Embedding(max_norm=1), as FsLoader::table
        float ss = 0.f;
        for (int j = 0; j < E; ++j) ss += et->data[(size_t)r * E + j] * et->data[(size_t)r * E + j];
        const float nrm = std::sqrt(ss), sc = nrm > 1.0f ? 1.0f / (nrm + 1e-7f) : 1.0f;
        for (int j = 0; j < E; ++j) e[(size_t)r * E + j] = nrm > 1.0f ? et->data[(size_t)r * E + j] * sc : et->data[(size_t)r * E + j];
      }
      std::vector<double> q = linear(e, ne, E, qw->data, qb->data, dq);
      TRY(upload(q, &fs->dec_q_tab));
      layer_norm(q, ne, dq, nw->data, nb->data);
      TRY(upload(linear(q, ne, dq, wq->data, nullptr, inner), &fs->dec_qq_tab));
    }
  }
  *out = guard.release();
  return 0;
}

extern "C" void lamslide_first_stage_destroy(lamslide_first_stage* h) { delete h; }

// ---- bump allocator over the caller's workspace (dry run when base == nullptr)
struct Bump {
  uint8_t* base;
  size_t off = 0;
  explicit Bump(void* b) : base((uint8_t*)b) {}
  float* f(size_t n) {
    float* p = base ? (float*)(base + off) : nullptr;
    off += align_up(n * 4, 256);
    return p;
  }
};

static int launch_linear_tc5(const LinW& L, const LinearArgs& a, cudaStream_t st);
static int fs_layernorm(const float* x, int ldx, int period, float* y, int ldy, const LNW* n, long long rows, int dim, float eps,
                        cudaStream_t st);
// LayerNorm to apply to the result of a linear layer (the next sub-layer's pre-norm): out = LN(y) [* n->w + n->b] over groups of
// `group` columns (0: the whole row).  y_needed = false: the un-normalised result itself is not used by anybody.
struct FsLN {
  float* out = nullptr;
  int ld = 0;
  const LNW* n = nullptr;
  int group = 0;
  bool y_needed = true;
};
// the same LayerNorm as a separate launch (layers the tcgen05 kernel does not take, or a group it cannot hold)
static int fs_ln_after(const FsLN& ln, const float* Y, int ldy, long long rows, int N, int group, cudaStream_t st) {
  if (group == N) return fs_layernorm(Y, ldy, 0, ln.out, ln.ld, ln.n, rows, N, 1e-5f, st);
  if (ldy != N || ln.ld != N || N % group != 0) return fail(LAMSLIDE_ERR_INVALID, "grouped LayerNorm needs dense rows");
  return fs_layernorm(Y, group, 0, ln.out, group, ln.n, rows * (N / group), group, 1e-5f, st);
}
// act: 0 none, 1 erf-GELU (before the adds), 2 SiLU of the sum (after the adds)
static int fs_linear(const LinW& L, const float* X, int ldx, float* Y, int ldy, long long rows, int act, const float* res, int ldr,
                     const float* rowadd, int period, int ldra, cudaStream_t st, const FsLN* ln, const long long* res_idx) {
  LinearArgs a;
  a.X = X, a.ldx = ldx, a.W = L.w, a.bias = L.b, a.Y = Y, a.ldy = ldy, a.res = res, a.ldr = ldr, a.res_idx = res_idx;
  a.rowadd = rowadd, a.rowadd_period = period > 0 ? period : 1, a.ldra = ldra;
  a.rows = (int)rows, a.N = L.out, a.K = L.in, a.gelu = act;
  a.debug = env_int("LAMSLIDE_L5_DEBUG", 0);
  static const bool legacy_fs = env_flag("LAMSLIDE_LEGACY_FS_LINEAR");
  const bool vec_ok = !legacy_fs && L.in % 4 == 0 && ldx % 4 == 0 && ((uintptr_t)X & 15) == 0 && ((uintptr_t)L.w & 15) == 0;
  static const bool no_tc5 = env_flag("LAMSLIDE_FS_NO_TCGEN05");
  static const bool no_ln_epi = env_flag("LAMSLIDE_FS_NO_LN_EPILOGUE");
  const int group = ln ? (ln->group > 0 ? ln->group : L.out) : 0;
  if (L.bn && !legacy_fs && !no_tc5 && ldx % 4 == 0 && ((uintptr_t)X & 15) == 0) {
    // the LayerNorm runs in the epilogue when a thread pair (group = the row, one n-tile) or a thread (group = half a tile) holds a group
    const bool row_ln = ln && group == L.out && L.out <= L.bn;
    const bool half_ln = ln && !row_ln && group * 2 == L.bn && L.out % L.bn == 0 && ln->ld == L.out;
    const bool fuse_ln = !no_ln_epi && (row_ln || half_ln);
    if (fuse_ln) {
      a.ln_out = ln->out, a.ld_ln = ln->ld, a.ln_group = row_ln ? 0 : group;
      a.ln_w = ln->n ? ln->n->w : nullptr, a.ln_b = ln->n ? ln->n->b : nullptr;
      if (!ln->y_needed) a.Y = nullptr;
    }
    TRY(launch_linear_tc5(L, a, st));
    if (ln && !fuse_ln) TRY(fs_ln_after(*ln, Y, ldy, rows, L.out, group, st));
    return 0;
  }
  static const bool no_tc = env_flag("LAMSLIDE_FS_NO_TF32X3");
  // the kernel is chosen by the layer's shape only, never by the row count: a trajectory must decode to the same bits whether it is
  // sampled alone, in a batch of 64 or as a shard of a multi-GPU run
  const bool tc_ok = vec_ok && !no_tc && L.out % 2 == 0 && ldy % 2 == 0 && ((uintptr_t)Y & 7) == 0 && (!res || (ldr % 2 == 0 && ((uintptr_t)res & 7) == 0));
  if (tc_ok) {
    dim3 grid(cdiv(L.out, 64), cdiv(rows, 128));
    linear_f32_tc_kernel<<<grid, 256, 0, st>>>(a);
  } else if (vec_ok) {
    dim3 grid(cdiv(L.out, 64), cdiv(rows, 128));
    linear_f32_v2_kernel<<<grid, 256, 0, st>>>(a);
  } else {
    dim3 grid(cdiv(L.out, 64), cdiv(rows, 64));
    linear_f32_kernel<<<grid, 256, 0, st>>>(a);
  }
  LAUNCH_CHECK();
  if (ln) TRY(fs_ln_after(*ln, Y, ldy, rows, L.out, group, st));
  return 0;
}
template <int BN>
static int launch_linear_tc5_bn(const LinW& L, const LinearArgs& a, cudaStream_t st) {
  using C = L5Cfg<BN>;
  auto kern = linear_tc5_kernel<BN>;
  TRY(ensure_dynamic_smem((const void*)kern, C::kSmem));
  CUtensorMap ta;
  TRY(make_tmap_f32(&ta, a.X, (uint64_t)a.rows, (uint64_t)a.K, (uint64_t)a.ldx, 128));
  const int m_tiles = cdiv(a.rows, 128), n_tiles = cdiv(a.N, BN);
  const int grid = std::min(m_tiles * n_tiles, num_sms());
  kern<<<grid, kL5Threads, C::kSmem, st>>>(ta, L.tm_hi, L.tm_lo, a, m_tiles, n_tiles);
  LAUNCH_CHECK();
  COUNT_KERNEL("linear_tc5");
  return 0;
}
static int launch_linear_tc5(const LinW& L, const LinearArgs& a, cudaStream_t st) {
  switch (L.bn) {
    case 32: return launch_linear_tc5_bn<32>(L, a, st);
    case 64: return launch_linear_tc5_bn<64>(L, a, st);
    case 96: return launch_linear_tc5_bn<96>(L, a, st);
    case 128: return launch_linear_tc5_bn<128>(L, a, st);
    case 192: return launch_linear_tc5_bn<192>(L, a, st);
    case 256: return launch_linear_tc5_bn<256>(L, a, st);
  }
  return fail(LAMSLIDE_ERR_INVALID, "linear_tc5: no kernel for tile width %d", L.bn);
}
static int fs_layernorm(const float* x, int ldx, int period, float* y, int ldy, const LNW* n, long long rows, int dim, float eps,
                        cudaStream_t st) {
  layernorm_f32_kernel<<<cdiv(rows, 8), 256, 0, st>>>(x, ldx, period, y, ldy, n ? n->w : nullptr, n ? n->b : nullptr, (int)rows, dim, eps);
  LAUNCH_CHECK();
  return 0;
}
static int fs_attn(const SmallAttnArgs& a, cudaStream_t st) {
  long long tot = (long long)a.frames * a.heads * a.Sq;
  small_attn_f32_kernel<<<cdiv(tot, 128), 128, 0, st>>>(a);
  LAUNCH_CHECK();
  return 0;
}
// dst[r, col0 : col0 + width] = table[idx[r]]
static int fs_gather(float* dst, int ldd, int col0, const float* table, int width, const long long* idx, long long rows, cudaStream_t st) {
  if (width % 4 == 0 && col0 % 4 == 0 && ldd % 4 == 0 && ((uintptr_t)dst & 15) == 0 && ((uintptr_t)table & 15) == 0)
    gather_cols4_kernel<<<cdiv(rows * (width / 4), 256), 256, 0, st>>>(dst, ldd, col0, table, width, idx, rows);
  else
    gather_cols_kernel<<<cdiv(rows * width, 256), 256, 0, st>>>(dst, ldd, col0, table, width, idx, rows);
  LAUNCH_CHECK();
  return 0;
}

struct BlockScratch {
  float *xn, *cn, *q, *kv, *att, *f1;
};
static BlockScratch block_scratch(Bump& bp, long long rows_x, int dim, long long rows_ctx, int ctx_dim, int inner) {
  BlockScratch s;
  s.xn = bp.f((size_t)rows_x * dim);
  s.cn = bp.f((size_t)rows_ctx * ctx_dim);
  s.q = bp.f((size_t)rows_x * inner);
  s.kv = bp.f((size_t)std::max(rows_ctx * 2, rows_x * 3) * inner);
  s.att = bp.f((size_t)rows_x * inner);
  s.f1 = bp.f((size_t)rows_x * dim);
  return s;
}

// What a block may find already done by its predecessor, and what it should leave for its successor.  Every LayerNorm whose input is
// the output of a linear layer runs in that layer's epilogue (FsLN); queries that depend only on the entity id or on the learned
// latents come from tables folded at *_create (FsFolded).
struct BlockIo {
  bool xn_ready = false;            // s.xn already holds LN(x; b.norm)
  bool cn_ready = false;            // s.cn already holds LN(ctx; b.norm_ctx)
  const float* q_table = nullptr;   // to_q(LN(x; b.norm)) as a table: rows q_index[row] (or, q_index null, row % Sx: the same for all frames)
  const long long* q_index = nullptr;
  const float* x_table = nullptr;   // the block input x itself as a table (residual of to_out): rows x_index[row] or row % Sx
  const long long* x_index = nullptr;
  const FsLN* after = nullptr;      // LayerNorm of the block's output (the successor's pre-norm)
};

// x [F*Sx, dim] (updated in place; written, not read, when io.x_table is given)  <-  CrossAttentionBlock(x, context = ctx [F*Sc, ctx_dim], mask)
static int run_cross_block(const AttnBlockW& b, float* x, int F, int Sx, const float* ctx, int Sc, const uint8_t* mask,
                           const BlockScratch& s, cudaStream_t st, const BlockIo& io = BlockIo()) {
  const long long rx = (long long)F * Sx, rc = (long long)F * Sc;
  const int inner = b.heads * b.dh;
  SmallAttnArgs a;
  if (io.q_table) {
    a.q = io.q_table, a.q_frame_stride = 0, a.ldq = inner, a.q_index = io.q_index;
  } else {
    if (!io.xn_ready) TRY(fs_layernorm(x, b.dim, 0, s.xn, b.dim, &b.norm, rx, b.dim, 1e-5f, st));
    TRY(fs_linear(b.to_q, s.xn, b.dim, s.q, inner, rx, false, nullptr, 0, nullptr, 0, 0, st));
    a.q = s.q, a.q_frame_stride = (long long)Sx * inner, a.ldq = inner;
  }
  if (!io.cn_ready) TRY(fs_layernorm(ctx, b.ctx_dim, 0, s.cn, b.ctx_dim, &b.norm_ctx, rc, b.ctx_dim, 1e-5f, st));
  TRY(fs_linear(b.to_kv, s.cn, b.ctx_dim, s.kv, 2 * inner, rc, false, nullptr, 0, nullptr, 0, 0, st));
  a.k = s.kv, a.v = s.kv + inner, a.kv_frame_stride = (long long)Sc * 2 * inner, a.ldkv = 2 * inner;  // k first, then v
  a.gq = b.gq, a.gk = b.gk, a.mask = mask, a.out = s.att, a.ldo = inner;
  a.frames = F, a.Sq = Sx, a.Sk = Sc, a.heads = b.heads, a.scale = 1.0f / std::sqrt((float)b.dh);
  TRY(fs_attn(a, st));
  const FsLN ff{s.xn, b.dim, &b.ff_norm, 0, true};
  if (io.x_table && io.x_index)
    TRY(fs_linear(b.to_out, s.att, inner, x, b.dim, rx, false, io.x_table, b.dim, nullptr, 0, 0, st, &ff, io.x_index));
  else if (io.x_table)
    TRY(fs_linear(b.to_out, s.att, inner, x, b.dim, rx, false, nullptr, 0, io.x_table, Sx, b.dim, st, &ff));
  else
    TRY(fs_linear(b.to_out, s.att, inner, x, b.dim, rx, false, x, b.dim, nullptr, 0, 0, st, &ff));
  TRY(fs_linear(b.ff0, s.xn, b.dim, s.f1, b.dim, rx, true, nullptr, 0, nullptr, 0, 0, st));
  TRY(fs_linear(b.ff1, s.f1, b.dim, x, b.dim, rx, false, x, b.dim, nullptr, 0, 0, st, io.after));
  return 0;
}

static int run_self_block(const AttnBlockW& b, float* x, int F, int S, const BlockScratch& s, cudaStream_t st, const BlockIo& io = BlockIo()) {
  const long long rx = (long long)F * S;
  const int inner = b.heads * b.dh;
  if (!io.xn_ready) TRY(fs_layernorm(x, b.dim, 0, s.xn, b.dim, &b.norm, rx, b.dim, 1e-5f, st));
  TRY(fs_linear(b.to_qkv, s.xn, b.dim, s.kv, 3 * inner, rx, false, nullptr, 0, nullptr, 0, 0, st));
  SmallAttnArgs a;
  a.q = s.kv, a.q_frame_stride = (long long)S * 3 * inner, a.ldq = 3 * inner;
  a.k = s.kv + inner, a.v = s.kv + 2 * inner, a.kv_frame_stride = (long long)S * 3 * inner, a.ldkv = 3 * inner;
  a.gq = b.gq, a.gk = b.gk, a.mask = nullptr, a.out = s.att, a.ldo = inner;
  a.frames = F, a.Sq = S, a.Sk = S, a.heads = b.heads, a.scale = 1.0f / std::sqrt((float)b.dh);
  TRY(fs_attn(a, st));
  const FsLN ff{s.xn, b.dim, &b.ff_norm, 0, true};
  TRY(fs_linear(b.to_out, s.att, inner, x, b.dim, rx, false, x, b.dim, nullptr, 0, 0, st, &ff));
  TRY(fs_linear(b.ff0, s.xn, b.dim, s.f1, b.dim, rx, true, nullptr, 0, nullptr, 0, 0, st));
  TRY(fs_linear(b.ff1, s.f1, b.dim, x, b.dim, rx, false, x, b.dim, nullptr, 0, 0, st, io.after));
  return 0;
}

static int encode_impl(lamslide_first_stage* fs, const lamslide_frame_inputs* in, float* latents_out, int F, int N, void* workspace,
                       size_t* bytes, cudaStream_t st) {
  const lamslide_first_stage_config& c = fs->cfg;
  const int Din = c.dim_input, D = c.dim_latent, E = c.entity_dim, C = Din + E, L = c.enc_num_latents;
  const long long R = (long long)F * N, FL = (long long)F * L;
  Bump bp(workspace);
  float* feat = bp.f((size_t)R * fs->feat_dim);
  float* pfeat = c.kind == LAMSLIDE_FS_MD17 ? bp.f((size_t)R * 129) : nullptr;
  float* hid1 = bp.f((size_t)R * Din);
  float* ctx_in = bp.f((size_t)R * C);
  float* hid2 = bp.f((size_t)R * D);
  float* ctx = bp.f((size_t)R * C);
  float* z = bp.f((size_t)FL * D);
  int inner = std::max(c.enc_heads_cross, c.enc_heads_latent) * 16;
  BlockScratch s = block_scratch(bp, FL, D, R, C, inner);
  if (bytes) {
    *bytes = bp.off;
    return 0;
  }
  const int TB = 256;
  // 1. Backbone.prepare_inputs features
  const float* feat_in = feat;
  switch (c.kind) {
    case LAMSLIDE_FS_PEPTIDE:
      if (!in->index0) return fail(LAMSLIDE_ERR_INVALID, "peptide encode needs aatype (index0)");
      if (N > c.max_res) return fail(LAMSLIDE_ERR_INVALID, "N = %d exceeds max_res = %d", N, c.max_res);
      TRY(fs_gather(feat, fs->feat_dim, 0, fs->tab0, 64, (const long long*)in->index0, R, st));
      copy_cols_kernel<<<cdiv(R * (fs->feat_dim - 64), TB), TB, 0, st>>>(feat, fs->feat_dim, 64, in->pos, 42, fs->feat_dim - fs->feat_used, R);
      LAUNCH_CHECK();
      break;
    case LAMSLIDE_FS_MD17:
      if (!in->index0) return fail(LAMSLIDE_ERR_INVALID, "md17 encode needs atom (index0)");
      TRY(fs_gather(feat, fs->feat_dim, 0, fs->tab0, 64, (const long long*)in->index0, R, st));
      point_feats_kernel<<<cdiv(R * 129, TB), TB, 0, st>>>(in->pos, fs->point_basis, pfeat, fs->point_nb, R);
      LAUNCH_CHECK();
      TRY(fs_linear(fs->point_mlp, pfeat, 129, feat + 64, fs->feat_dim, R, false, nullptr, 0, nullptr, 0, 0, st));
      break;
    case LAMSLIDE_FS_NBA:
      if (!in->index0 || !in->index1) return fail(LAMSLIDE_ERR_INVALID, "nba encode needs team (index0) and group (index1)");
      copy_cols_kernel<<<cdiv(R * 2, TB), TB, 0, st>>>(feat, fs->feat_dim, 0, in->pos, 2, 0, R);
      LAUNCH_CHECK();
      gather_cols_kernel<<<cdiv(R * 32, TB), TB, 0, st>>>(feat, fs->feat_dim, 2, fs->tab0, 32, (const long long*)in->index0, R);
      LAUNCH_CHECK();
      gather_cols_kernel<<<cdiv(R * (fs->feat_dim - 34), TB), TB, 0, st>>>(feat, fs->feat_dim, 34, fs->tab1, 32, (const long long*)in->index1, R,
                                                                           fs->feat_dim - fs->feat_used);
      LAUNCH_CHECK();
      break;
    default: feat_in = in->pos; break;
  }
  // net_merge (+ residue-index sin/cos embedding for peptides) written straight into the context matrix [x | E_ent]
  TRY(fs_linear(fs->merge0, feat_in, fs->feat_dim, hid1, Din, R, true, nullptr, 0, nullptr, 0, 0, st));
  TRY(fs_linear(fs->merge2, hid1, Din, ctx_in, C, R, false, nullptr, 0, fs->sincos, N, Din, st));
  TRY(fs_gather(ctx_in, C, Din, fs->ent_table, E, (const long long*)in->entities, R, st));
  // EncoderBase.prepare_inputs MLP (encoder.py:35-41)
  TRY(fs_linear(fs->enc_mlp0, ctx_in, C, hid2, D, R, true, nullptr, 0, nullptr, 0, 0, st));
  TRY(fs_linear(fs->enc_mlp2, hid2, D, ctx, C, R, false, nullptr, 0, nullptr, 0, 0, st));
  // the blocks: each leaves LN(z) with its successor's pre-norm in s.xn (z itself only where a residual needs it)
  const int ncross = (int)fs->enc_cross.size(), nself = (int)fs->enc_self.size();
  const bool fold0 = fs->enc_q0 != nullptr && ncross > 0;  // first cross block: its queries come from the learned latents alone
  if (!fold0) {
    bcast_rows_kernel<<<cdiv(FL * D, TB), TB, 0, st>>>(z, fs->latents, D, L, FL);
    LAUNCH_CHECK();
  }
  bool xn_ready = false;
  for (int i = 0; i < ncross + nself; ++i) {
    const bool cross = i < ncross;
    const AttnBlockW& b = cross ? fs->enc_cross[i] : fs->enc_self[i - ncross];
    const AttnBlockW* nb = i + 1 < ncross ? &fs->enc_cross[i + 1] : i + 1 < ncross + nself ? &fs->enc_self[i + 1 - ncross] : nullptr;
    FsLN after{s.xn, D, nb ? &nb->norm : nullptr, 0, true};
    BlockIo io;
    io.xn_ready = xn_ready;
    io.after = nb ? &after : nullptr;
    if (cross) {
      if (i == 0 && fold0) io.q_table = fs->enc_q0, io.x_table = fs->latents;
      TRY(run_cross_block(b, z, F, L, ctx, N, in->mask, s, st, io));
    } else {
      TRY(run_self_block(b, z, F, L, s, st, io));
    }
    xn_ready = nb != nullptr;
  }
  // quant: Linear -> LayerNorm(no affine) (lightning_base.py:24-27)
  const FsLN qln{latents_out, D, nullptr, 0, false};
  TRY(fs_linear(fs->quant, z, D, s.xn, D, FL, false, nullptr, 0, nullptr, 0, 0, st, &qln));
  return 0;
}

static int decode_impl(lamslide_first_stage* fs, const float* latents, const int64_t* entities, float* const* outs, int F, int N,
                       void* workspace, size_t* bytes, cudaStream_t st) {
  const lamslide_first_stage_config& c = fs->cfg;
  const int D = c.dim_latent, E = c.entity_dim, L = c.enc_num_latents, dq = c.dec_dim_query;
  const int Lk = c.dec_query_splitter ? L * c.dec_num_split : L;
  const long long R = (long long)F * N, FL = (long long)F * L, FLk = (long long)F * Lk;
  Bump bp(workspace);
  float* zl = bp.f((size_t)FL * D);
  float* z = bp.f((size_t)FL * D);
  float* ent = bp.f((size_t)R * E);
  float* Q = bp.f((size_t)R * dq);
  float* zk = c.dec_query_splitter ? bp.f((size_t)FLk * D) : nullptr;
  float* t1 = bp.f((size_t)R * dq);
  BlockScratch s_self = block_scratch(bp, FL, D, FL, D, c.dec_heads_latent * 16);
  BlockScratch s_cross = block_scratch(bp, FL, D, R, dq, c.dec_heads_cross * 16);
  BlockScratch s_out = block_scratch(bp, R, dq, FLk, D, c.dec_heads_cross * 16);
  if (bytes) {
    *bytes = bp.off;
    return 0;
  }
  const int nself = (int)fs->dec_self.size(), ncross = (int)fs->dec_cross.size();
  // queries = query_mlp(entity_embedding(entities)) (decoder.py:83-86; Dropout inactive in eval) depend on the entity id only: with
  // the folded tables (q_tab, its LayerNorm and to_q of the output block) the per-row matrix is built only as the context of
  // decoder cross blocks, which no shipped config has
  const bool folded = fs->dec_q_tab != nullptr;
  if (!folded || ncross > 0) {
    if (folded) {
      TRY(fs_gather(Q, dq, 0, fs->dec_q_tab, dq, (const long long*)entities, R, st));
    } else {
      TRY(fs_gather(ent, E, 0, fs->ent_table, E, (const long long*)entities, R, st));
      TRY(fs_linear(fs->query_mlp, ent, E, Q, dq, R, false, nullptr, 0, nullptr, 0, 0, st));
    }
  }
  // post_quant: LayerNorm(no affine) -> Linear (lightning_base.py:28-31), leaving the first block's pre-norm in its xn
  TRY(fs_layernorm(latents, D, 0, zl, D, nullptr, FL, D, 1e-5f, st));
  {
    const AttnBlockW* nb = nself > 0 ? &fs->dec_self[0] : ncross > 0 ? &fs->dec_cross[0] : nullptr;
    const FsLN after{nself > 0 ? s_self.xn : s_cross.xn, D, nb ? &nb->norm : nullptr, 0, true};
    TRY(fs_linear(fs->post_quant, zl, D, z, D, FL, false, nullptr, 0, nullptr, 0, 0, st, nb ? &after : nullptr));
  }
  for (int i = 0; i < nself + ncross; ++i) {
    const bool self = i < nself;
    const AttnBlockW& b = self ? fs->dec_self[i] : fs->dec_cross[i - nself];
    const AttnBlockW* nb = i + 1 < nself ? &fs->dec_self[i + 1] : i + 1 < nself + ncross ? &fs->dec_cross[i + 1 - nself] : nullptr;
    const bool next_self = i + 1 < nself;
    FsLN after{next_self ? s_self.xn : s_cross.xn, D, nb ? &nb->norm : nullptr, 0, true};
    BlockIo io;
    io.xn_ready = true;
    io.after = nb ? &after : nullptr;
    if (self) TRY(run_self_block(b, z, F, L, s_self, st, io));
    else TRY(run_cross_block(b, z, F, L, Q, N, nullptr, s_cross, st, io));
  }
  // output block: queries attend to the (extended) latents.  The LayerNorm of its context runs in the epilogue of the layer that
  // produces the context when that is the extender (per 'D'-wide token of the [F*L, n*D] row = the [F, L*n, D] token matrix).
  BlockIo io;
  const float* kvsrc = z;
  if (c.dec_query_splitter) {
    const FsLN kln{s_out.cn, D * c.dec_num_split, &fs->out_block.norm_ctx, D, false};
    TRY(fs_linear(fs->extender, z, D, zk, D * c.dec_num_split, FL, false, nullptr, 0, nullptr, 0, 0, st, &kln));
    kvsrc = zk;
    io.cn_ready = true;
  }
  if (folded) io.q_table = fs->dec_qq_tab, io.q_index = (const long long*)entities, io.x_table = fs->dec_q_tab, io.x_index = (const long long*)entities;
  TRY(run_cross_block(fs->out_block, Q, F, N, kvsrc, Lk, nullptr, s_out, st, io));
  for (int i = 0; i < c.n_outputs; ++i) {
    if (!outs[i]) continue;
    TRY(fs_linear(fs->head0[i], Q, dq, t1, dq, R, true, nullptr, 0, nullptr, 0, 0, st));
    TRY(fs_linear(fs->head2[i], t1, dq, outs[i], c.output_dims[i], R, false, nullptr, 0, nullptr, 0, 0, st));
  }
  return 0;
}

extern "C" size_t lamslide_first_stage_workspace_bytes(const lamslide_first_stage* h, int32_t frames, int32_t N) {
  if (!h || frames <= 0 || N <= 0) return 0;
  size_t a = 0, b = 0;
  encode_impl(const_cast<lamslide_first_stage*>(h), nullptr, nullptr, frames, N, nullptr, &a, nullptr);
  decode_impl(const_cast<lamslide_first_stage*>(h), nullptr, nullptr, nullptr, frames, N, nullptr, &b, nullptr);
  return std::max(a, b);
}

static int fs_check(lamslide_first_stage* h, int frames, int N, void* ws, size_t ws_bytes) {
  if (!h) return fail(LAMSLIDE_ERR_INVALID, "null handle");
  if (frames <= 0 || N <= 0) return fail(LAMSLIDE_ERR_INVALID, "frames and N must be positive");
  int dev = -1;
  CUDA_TRY(cudaGetDevice(&dev));
  if (dev != h->device) return fail(LAMSLIDE_ERR_INVALID, "handle was created on device %d, current device is %d", h->device, dev);
  size_t need = lamslide_first_stage_workspace_bytes(h, frames, N);
  if (!ws || ws_bytes < need) return fail(LAMSLIDE_ERR_WORKSPACE, "workspace too small: %zu < %zu bytes", ws_bytes, need);
  if (((uintptr_t)ws & 255) != 0) return fail(LAMSLIDE_ERR_INVALID, "workspace must be 256-byte aligned");
  return 0;
}

extern "C" int lamslide_encode(lamslide_first_stage* h, const lamslide_frame_inputs* in, float* latents_out, int32_t frames, int32_t N,
                               void* workspace, size_t workspace_bytes, void* stream) {
  if (!in || !in->pos || !in->entities || !latents_out) return fail(LAMSLIDE_ERR_INVALID, "null argument");
  TRY(fs_check(h, frames, N, workspace, workspace_bytes));
  ProfScope ps(PC_ENCODE, (cudaStream_t)stream);
  return encode_impl(h, in, latents_out, frames, N, workspace, nullptr, (cudaStream_t)stream);
}

extern "C" int lamslide_decode(lamslide_first_stage* h, const float* latents, const int64_t* entities, float* const* outs,
                               int32_t frames, int32_t N, void* workspace, size_t workspace_bytes, void* stream) {
  if (!latents || !entities || !outs) return fail(LAMSLIDE_ERR_INVALID, "null argument");
  TRY(fs_check(h, frames, N, workspace, workspace_bytes));
  ProfScope ps(PC_DECODE, (cudaStream_t)stream);
  return decode_impl(h, latents, entities, outs, frames, N, workspace, nullptr, (cudaStream_t)stream);
}

extern "C" int lamslide_debug_fs_linear(const float* x, const float* w_host, const float* bias_host, float* y, int32_t rows, int32_t N,
                                        int32_t K, int32_t ldx, int32_t ldy, int32_t act, const float* res, int32_t ldr,
                                        const float* rowadd, int32_t period, int32_t ldra, int32_t path, void* stream) {
  if (!x || !w_host || !y || rows <= 0 || N <= 0 || K <= 0) return fail(LAMSLIDE_ERR_INVALID, "bad argument");
  Arena arena;
  LinW L;
  TRY(arena.upload_f32(w_host, (size_t)N * K, &L.w));
  if (bias_host) TRY(arena.upload_f32(bias_host, N, &L.b));
  L.out = N, L.in = K;
  if (path == 0) TRY(prepare_tc5(arena, w_host, &L));
  TRY(fs_linear(L, x, ldx, y, ldy, rows, act, res, ldr, rowadd, period, ldra, (cudaStream_t)stream));
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));  // the weights die with `arena`
  return 0;
}

// the same layer followed by the LayerNorm of its result (the successor's pre-norm; in the kernel's epilogue where it can hold a group):
// ln_out [rows, N] = LN over groups of `group` columns (0: the row) [* ln_w + ln_b]; y is skipped when y_needed == 0 and the LN is fused
extern "C" int lamslide_debug_fs_linear_ln(const float* x, const float* w_host, const float* bias_host, float* y, float* ln_out,
                                           const float* ln_w, const float* ln_b, int32_t group, int32_t rows, int32_t N, int32_t K,
                                           const float* res, int32_t y_needed, void* stream) {
  if (!x || !w_host || !y || !ln_out || rows <= 0 || N <= 0 || K <= 0) return fail(LAMSLIDE_ERR_INVALID, "bad argument");
  Arena arena;
  LinW L;
  TRY(arena.upload_f32(w_host, (size_t)N * K, &L.w));
  if (bias_host) TRY(arena.upload_f32(bias_host, N, &L.b));
  L.out = N, L.in = K;
  TRY(prepare_tc5(arena, w_host, &L));
  LNW n;
  n.w = const_cast<float*>(ln_w), n.b = const_cast<float*>(ln_b);
  const FsLN ln{ln_out, N, ln_w ? &n : nullptr, group, y_needed != 0};
  TRY(fs_linear(L, x, K, y, N, rows, 0, res, N, nullptr, 0, 0, (cudaStream_t)stream, &ln));
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}

// ================================================================================================ misc + test hooks
extern "C" int lamslide_abi_version(void) { return LAMSLIDE_ABI_VERSION; }
extern "C" const char* lamslide_last_error(void) { return g_err.c_str(); }
extern "C" int64_t lamslide_launch_count(int32_t reset) {
  int64_t v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

extern "C" void lamslide_debug_attention_trace(void* device_buffer) { g_atc_trace = (long long*)device_buffer; }

extern "C" int64_t lamslide_debug_kernel_count(const char* name, int32_t reset) {
  if (!name) return -1;
  auto it = g_named_launches.find(name);
  const int64_t v = it == g_named_launches.end() ? 0 : it->second;
  if (reset) g_named_launches.clear();
  return v;
}

extern "C" int lamslide_profile_begin(void) {
  for (auto& r : g_prof) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_prof.clear();
  g_prof_on = true;
  return 0;
}

extern "C" int lamslide_profile_end(char* json_out, size_t json_bytes) {
  g_prof_on = false;
  CUDA_TRY(cudaDeviceSynchronize());
  double ms[PC_COUNT] = {0};
  long long cnt[PC_COUNT] = {0};
  for (auto& r : g_prof) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) ms[r.cat] += t, cnt[r.cat]++;
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_prof.clear();
  std::string js = "{";
  for (int i = 0; i < PC_COUNT; ++i) {
    char buf[160];
    snprintf(buf, sizeof(buf), "%s\"%s\": {\"ms\": %.6f, \"launch_groups\": %lld}", i ? ", " : "", kProfNames[i], ms[i], cnt[i]);
    js += buf;
  }
  js += "}";
  if (!json_out || json_bytes < js.size() + 1) return fail(LAMSLIDE_ERR_INVALID, "profile buffer too small (%zu needed)", js.size() + 1);
  memcpy(json_out, js.c_str(), js.size() + 1);
  return 0;
}

extern "C" int lamslide_debug_gemm(const void* a_bf16, const void* b_bf16, const float* bias, float* c, int32_t M, int32_t N, int32_t K,
                                   int32_t block_n, void* stream) {
  if (!a_bf16 || !b_bf16 || !c) return fail(LAMSLIDE_ERR_INVALID, "null argument");
  int bn = block_n > 0 ? block_n : plain_bn_for(N);
  if (bn <= 0 || N % bn != 0) return fail(LAMSLIDE_ERR_INVALID, "N = %d is not a multiple of the tile width %d", N, bn);
  CUtensorMap ta, tb;
  TRY(make_tmap(&ta, a_bf16, M, K, kBlockM));
  TRY(make_tmap(&tb, b_bf16, N, K, bn));
  EpiPlain::Params ep{c, bias, N, M};
  return launch_plain(bn, ta, tb, M, N, K, ep, (cudaStream_t)stream);
}

extern "C" int lamslide_debug_attention(const void* qkv_bf16, void* out_bf16, int32_t B, int32_t T, int32_t L, int32_t H, int32_t heads,
                                        int32_t ldo, int32_t temporal, int32_t force_flash, void* stream) {
  if (!qkv_bf16 || !out_bf16 || heads <= 0 || H % heads) return fail(LAMSLIDE_ERR_INVALID, "bad argument");
  SeqMap sm = temporal ? SeqMap{T, L, T * L, 1, L} : SeqMap{L, 1, L, 0, 1};
  int n_seq = temporal ? B * L : B * T;
  return attention_dispatch((const __nv_bfloat16*)qkv_bf16, (__nv_bfloat16*)out_bf16, H, ldo, heads, H / heads, sm, n_seq,
                            force_flash, 0.f, (cudaStream_t)stream);
}

// ---- linear1 / linear2 with their fused epilogues in isolation (tests/test_gpu_kernels.py).  Debug only: allocates and syncs.
template <int HD>
static int debug_linear1_impl(const void* u, const void* w1, const float* bias, const float* gq, const float* gk, void* qkv, void* act,
                              int rows, int H, int M, int pos_div, int pos_mod, float theta, int legacy, cudaStream_t st) {
  const int N = 3 * H + M, half = HD / 2;
  const int bn = HD == 24 ? pick_bn({192, 96}, H, M, HD) : pick_bn({128, 64}, H, M, HD);
  if (!bn) return fail(LAMSLIDE_ERR_INVALID, "no linear1 tiling for H %d M %d hd %d", H, M, HD);
  static float *cs = nullptr, *sn = nullptr;  // debug hook: cached RoPE tables (not thread-safe)
  static size_t cap = 0;
  static int have_pos = 0, have_half = 0;
  static float have_theta = 0.f;
  const size_t need = (size_t)pos_mod * half * 4;
  if (need > cap) {
    cudaFree(cs), cudaFree(sn);
    CUDA_TRY(cudaMalloc(&cs, need));
    CUDA_TRY(cudaMalloc(&sn, need));
    cap = need, have_pos = 0;
  }
  if (have_pos != pos_mod || have_half != half || have_theta != theta) {
    rope_table_kernel<<<cdiv(pos_mod * half, 256), 256, 0, st>>>(cs, sn, pos_mod, half, (double)theta);
    have_pos = pos_mod, have_half = half, have_theta = theta;
  }
  CUtensorMap ta, tb, tbh_map, tq, tact;
  TRY(make_tmap(&ta, u, rows, H, kBlockM));
  TRY(make_tmap(&tb, w1, N, H, bn));
  TRY(make_tmap(&tbh_map, w1, N, H, bn / 2));
  TRY(make_tmap_ex(&tq, qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, rows, 3 * H, 2 * HD, 32, CU_TENSOR_MAP_SWIZZLE_NONE));
  TRY(make_tmap_ex(&tact, act, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, rows, H + M, 2 * HD, 32, CU_TENSOR_MAP_SWIZZLE_NONE));
  const CUtensorMap* tbh = (legacy & 16) ? nullptr : &tbh_map;  // +16: force the one-CTA (no multicast) variant
  const bool flags32 = (legacy & 32) != 0;
  const bool fused_attn = (legacy & 64) != 0;  // +64: spatial attention in the epilogue, sequences of pos_mod consecutive rows
  legacy &= 15;
  const float q_premul = (float)(1.4426950408889634 / std::sqrt((double)HD));
  int rc;
  if (fused_attn) {
    if (pos_div != 1) return fail(LAMSLIDE_ERR_INVALID, "fused spatial attention needs pos_div = 1");
    float hq[32], hk[32];
    CUDA_TRY(cudaMemcpy(hq, gq, HD * 4, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(hk, gk, HD * 4, cudaMemcpyDeviceToHost));
    lamslide_backbone fake;  // only H and bn1 are read by launch_linear1_attn
    fake.H = H, fake.bn1 = bn;
    BlockWeights bw;
    bw.tm_w1 = tb, bw.tm_w1_h = tbh_map;
#define DBG_ATTN(AL_)                                                                                                              \
  {                                                                                                                                \
    typename EpiLinear1Ws<HD, AL_>::Params ep{bias, {}, cs, sn, (__nv_bfloat16*)qkv, (__nv_bfloat16*)act, H, 0, rows, 1, pos_mod, 0, H + M}; \
    for (int j = 0; j < HD; ++j) ep.gam[0][j] = hq[j] * q_premul, ep.gam[1][j] = hk[j];                                            \
    rc = launch_linear1_attn<HD, AL_>(&fake, ta, bw, tq, tact, rows, ep, st);                                                      \
  }
    if (pos_mod == 2) DBG_ATTN(2)
    else if (pos_mod == 4) DBG_ATTN(4)
    else if (pos_mod == 8) DBG_ATTN(8)
    else rc = 1;
#undef DBG_ATTN
    if (rc == 1) rc = fail(LAMSLIDE_ERR_INVALID, "fused spatial attention does not cover H %d hd %d L %d rows %d", H, HD, pos_mod, rows);
    return rc;
  }
  if (legacy != 1) {
    typename EpiLinear1Ws<HD>::Params ep{bias, {}, cs, sn, (__nv_bfloat16*)qkv, (__nv_bfloat16*)act, H, M, rows, pos_div, pos_mod,
                                         legacy == 2 ? 1 : legacy == 3 ? 2 : legacy == 4 ? 3 : legacy == 5 ? 4 : 0, H + M};
    {  // debug hook: the scales arrive as device pointers; fetch them once per distinct pointer pair (not thread-safe)
      static const float *last_q = nullptr, *last_k = nullptr;
      static float hq[32], hk[32];
      if (!(flags32) || last_q != gq || last_k != gk) {  // +32: the scales are unchanged since the last call (timing loops)
        CUDA_TRY(cudaMemcpy(hq, gq, HD * 4, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(hk, gk, HD * 4, cudaMemcpyDeviceToHost));
        last_q = gq, last_k = gk;
      }
      for (int j = 0; j < HD; ++j) ep.gam[0][j] = hq[j] * q_premul, ep.gam[1][j] = hk[j];
    }
    if constexpr (HD == 24) {
      rc = bn == 192 ? launch_gemm_ws<192, EpiLinear1Ws<24>>(ta, tb, tbh, tq, tact, rows, N, H, ep, st) : 1;
    } else {
      if (bn == 128) rc = launch_gemm_ws<128, EpiLinear1Ws<HD>>(ta, tb, tbh, tq, tact, rows, N, H, ep, st);
      else if constexpr (HD == 16) rc = launch_gemm_ws<64, EpiLinear1Ws<16>>(ta, tb, tbh, tq, tact, rows, N, H, ep, st);
      else rc = 1;
    }
    if (rc == 1) rc = fail(LAMSLIDE_ERR_INVALID, "persistent linear1 kernel does not cover H %d M %d hd %d", H, M, HD);
  } else {
    typename EpiLinear1<HD>::Params ep{bias, gq, gk, cs, sn, (__nv_bfloat16*)qkv, (__nv_bfloat16*)act, H, M, rows, pos_div, pos_mod, q_premul};
    if constexpr (HD == 24) {
      rc = bn == 192 ? launch_gemm<192, EpiLinear1<24>>(ta, tb, rows, N, H, ep, st) : launch_gemm<96, EpiLinear1<24>>(ta, tb, rows, N, H, ep, st);
    } else {
      rc = bn == 128 ? launch_gemm<128, EpiLinear1<HD>>(ta, tb, rows, N, H, ep, st) : launch_gemm<64, EpiLinear1<HD>>(ta, tb, rows, N, H, ep, st);
    }
  }
  return rc;
}

extern "C" int lamslide_debug_linear1(const void* u_bf16, const void* w1_bf16, const float* bias, const float* q_scale,
                                      const float* k_scale, void* qkv_bf16, void* act_bf16, int32_t rows, int32_t H, int32_t M,
                                      int32_t heads, int32_t pos_div, int32_t pos_mod, float theta, int32_t legacy, void* stream) {
  if (!u_bf16 || !w1_bf16 || !bias || !q_scale || !k_scale || !qkv_bf16 || !act_bf16 || heads <= 0 || H % heads)
    return fail(LAMSLIDE_ERR_INVALID, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  switch (H / heads) {
    case 16: return debug_linear1_impl<16>(u_bf16, w1_bf16, bias, q_scale, k_scale, qkv_bf16, act_bf16, rows, H, M, pos_div, pos_mod, theta, legacy, st);
    case 24: return debug_linear1_impl<24>(u_bf16, w1_bf16, bias, q_scale, k_scale, qkv_bf16, act_bf16, rows, H, M, pos_div, pos_mod, theta, legacy, st);
    case 32: return debug_linear1_impl<32>(u_bf16, w1_bf16, bias, q_scale, k_scale, qkv_bf16, act_bf16, rows, H, M, pos_div, pos_mod, theta, legacy, st);
    default: return fail(LAMSLIDE_ERR_INVALID, "head_dim %d unsupported", H / heads);
  }
}

extern "C" int lamslide_debug_linear2(const void* act_bf16, const void* w2_bf16, const float* bias, const float* gate, float* h,
                                      int32_t rows, int32_t H, int32_t M, int32_t rows_per_sample, int32_t legacy, void* stream) {
  if (!act_bf16 || !w2_bf16 || !bias || !gate || !h) return fail(LAMSLIDE_ERR_INVALID, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int bn = pick_bn({192, 128, 96, 64}, H, 0, 0);
  if (!bn) return fail(LAMSLIDE_ERR_INVALID, "no linear2 tiling for H %d", H);
  CUtensorMap ta, tb, tbh_map, th;
  TRY(make_tmap(&ta, act_bf16, rows, H + M, kBlockM));
  TRY(make_tmap(&tb, w2_bf16, H, H + M, bn));
  TRY(make_tmap(&tbh_map, w2_bf16, H, H + M, bn / 2));
  const CUtensorMap* tbh = (legacy & 16) ? nullptr : &tbh_map;
  legacy &= 15;
  TRY(make_tmap_ex(&th, h, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, rows, H, 16, 32, CU_TENSOR_MAP_SWIZZLE_64B));
  int rc;
  if (!legacy) {
    EpiLinear2Ws::Params ep{bias, gate, H, rows_per_sample, H, rows};
    switch (bn) {
      case 192: rc = launch_gemm_ws<192, EpiLinear2Ws>(ta, tb, tbh, th, th, rows, H, H + M, ep, st); break;
      case 128: rc = launch_gemm_ws<128, EpiLinear2Ws>(ta, tb, tbh, th, th, rows, H, H + M, ep, st); break;
      case 64: rc = launch_gemm_ws<64, EpiLinear2Ws>(ta, tb, tbh, th, th, rows, H, H + M, ep, st); break;
      default: rc = 1; break;
    }
    if (rc == 1) rc = fail(LAMSLIDE_ERR_INVALID, "persistent linear2 kernel does not cover H %d M %d", H, M);
  } else {
    EpiLinear2::Params ep{h, bias, gate, H, rows_per_sample, H, rows};
    switch (bn) {
      case 192: rc = launch_gemm<192, EpiLinear2>(ta, tb, rows, H, H + M, ep, st); break;
      case 128: rc = launch_gemm<128, EpiLinear2>(ta, tb, rows, H, H + M, ep, st); break;
      case 96: rc = launch_gemm<96, EpiLinear2>(ta, tb, rows, H, H + M, ep, st); break;
      default: rc = launch_gemm<64, EpiLinear2>(ta, tb, rows, H, H + M, ep, st); break;
    }
  }
  return rc;
}

// the TMA + tcgen05 main loop of the persistent kernel alone (accumulators are drained and dropped): what the loads and
// the tensor pipe can sustain for a shape, independent of any epilogue.  block_n in {192, 128, 64}.
extern "C" int lamslide_debug_gemm_mainloop(const void* a_bf16, const void* b_bf16, int32_t rows, int32_t N, int32_t K, int32_t block_n,
                                            void* stream) {
  if (!a_bf16 || !b_bf16 || block_n == 0 || N % block_n) return fail(LAMSLIDE_ERR_INVALID, "bad argument");
  const bool one_cta = block_n < 0;  // negative block_n: force the one-CTA (no multicast) variant
  if (one_cta) block_n = -block_n;
  CUtensorMap ta, tb, tbh_map;
  TRY(make_tmap(&ta, a_bf16, rows, K, kBlockM));
  TRY(make_tmap(&tb, b_bf16, N, K, block_n));
  TRY(make_tmap(&tbh_map, b_bf16, N, K, block_n / 2));
  const CUtensorMap* tbh = one_cta ? nullptr : &tbh_map;
  EpiNullWs::Params ep{0};
  int rc;
  switch (block_n) {
    case 192: rc = launch_gemm_ws<192, EpiNullWs>(ta, tb, tbh, ta, ta, rows, N, K, ep, (cudaStream_t)stream); break;
    case 128: rc = launch_gemm_ws<128, EpiNullWs>(ta, tb, tbh, ta, ta, rows, N, K, ep, (cudaStream_t)stream); break;
    case 64: rc = launch_gemm_ws<64, EpiNullWs>(ta, tb, tbh, ta, ta, rows, N, K, ep, (cudaStream_t)stream); break;
    default: return fail(LAMSLIDE_ERR_INVALID, "block_n %d unsupported", block_n);
  }
  if (rc == 1) return fail(LAMSLIDE_ERR_INVALID, "shape does not fit the persistent kernel");
  return rc;
}

// fused MLP half + linear2 + gated residual in isolation (tests/test_gpu_kernels.py):  h += gate[b] * ([attn | gelu(u W1m^T + b1m)] W2^T + b2)
// u [rows,H] bf16, act [rows,H+M] bf16 (only the attention half [:, :H] is read), w1 [3H+M,H] bf16, w2 [H,H+M] bf16, b1 [3H+M], b2 [H].
static int debug_fused_mlp(const void* u_bf16, const void* act_bf16, const void* w1_bf16, const void* w2_bf16, const float* b1,
                           const float* b2, const float* gate, float* h, int32_t rows, int32_t H, int32_t M, int32_t rows_per_sample,
                           const float* ln_shift, const float* ln_scale, void* u_out, void* stream) {
  if (!u_bf16 || !act_bf16 || !w1_bf16 || !w2_bf16 || !b1 || !b2 || !gate || !h) return fail(LAMSLIDE_ERR_INVALID, "bad argument");
  CUtensorMap tu, ta, tw1, tw2, th, tus;
  TRY(make_tmap_ex(&tus, u_out ? u_out : u_bf16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, rows, H, 16, 32, CU_TENSOR_MAP_SWIZZLE_32B));
  TRY(make_tmap(&tu, u_bf16, rows, H, kBlockM));
  TRY(make_tmap(&ta, act_bf16, rows, H + M, kBlockM));
  TRY(make_tmap(&tw1, w1_bf16, 3 * H + M, H, 64));
  TRY(make_tmap(&tw2, w2_bf16, H, H + M, fused_mlp_out_unit(H) / 2));
  TRY(make_tmap_ex(&th, h, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, rows, H, 16, 32, CU_TENSOR_MAP_SWIZZLE_64B));
  FusedMlpParams fp{b1 + 3 * H, b2, gate, H, rows_per_sample, H, M, rows, nullptr, nullptr, nullptr, nullptr, nullptr, 0};
  if (ln_scale) fp.h = h, fp.u_out = (__nv_bfloat16*)u_out, fp.ln_shift = ln_shift, fp.ln_scale = ln_scale;
  int rc = launch_fused_mlp(tu, ta, tw1, tw2, th, tus, rows, fp, (cudaStream_t)stream);
  if (rc == 1) return fail(LAMSLIDE_ERR_INVALID, "fused MLP kernel does not cover H %d M %d", H, M);
  return rc;
}
extern "C" int lamslide_debug_fused_mlp(const void* u_bf16, const void* act_bf16, const void* w1_bf16, const void* w2_bf16, const float* b1,
                                        const float* b2, const float* gate, float* h, int32_t rows, int32_t H, int32_t M,
                                        int32_t rows_per_sample, void* stream) {
  return debug_fused_mlp(u_bf16, act_bf16, w1_bf16, w2_bf16, b1, b2, gate, h, rows, H, M, rows_per_sample, nullptr, nullptr, nullptr, stream);
}
// same with the drain that also writes the next block's LN + modulate input: u_out [rows, H] bf16 = LN(h_new) * (1 + scale[b]) + shift[b]
// (shift, scale: [n_samples, H] device fp32; u_out may alias u_bf16)
extern "C" int lamslide_debug_fused_mlp_ln(const void* u_bf16, const void* act_bf16, const void* w1_bf16, const void* w2_bf16,
                                           const float* b1, const float* b2, const float* gate, float* h, int32_t rows, int32_t H,
                                           int32_t M, int32_t rows_per_sample, const float* ln_shift, const float* ln_scale, void* u_out,
                                           void* stream) {
  if (!ln_shift || !ln_scale || !u_out) return fail(LAMSLIDE_ERR_INVALID, "bad argument");
  return debug_fused_mlp(u_bf16, act_bf16, w1_bf16, w2_bf16, b1, b2, gate, h, rows, H, M, rows_per_sample, ln_shift, ln_scale, u_out, stream);
}

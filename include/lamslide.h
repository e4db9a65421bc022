/* lamslide.h — C ABI of the B200-native LaM-SLidE sampling hot path (liblamslide.so).
 *
 * The reference (ml-jku/LaM-SLidE) is pure Python and has no FFI of its own: its "operator API" for this path is the
 * nn.Module signatures listed below.  Each entry point replaces one of them; a maintainer binds them with ctypes
 * (see INTEGRATION.md — the stub is what lam_slide_b200/_lib.py does).
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch types.  `stream` is a cudaStream_t passed as void*.
 *   - every function returns 0 on success, a negative lamslide_status otherwise; lamslide_last_error() gives the text.
 *     (Reference behaviour: Python exceptions — ValueError for hidden_size % num_heads, latent_si_v31.py:92-95;
 *     shape asserts transport.py:197-199.)
 *   - weights are passed ONCE at create time as named host fp32 tensors using the reference's state-dict keys
 *     (SURVEY.md §8(b)); the library packs them (bf16 K-major for tensor-core operands, fp32 for vectors) into device
 *     memory it owns.  All activations / workspaces are caller-owned DEVICE buffers; no allocation, no host sync and no
 *     default-stream work happens inside forward / sample / encode / decode, so calls can be captured in a CUDA graph.
 *   - a handle is bound to the device that was current at create time; not thread-safe per handle.
 *   - tensors are contiguous, row-major, in the reference's layouts:  x, x_cond, out: [B,T,L,D] fp32;
 *     x_cond_mask: [B,T,L] int64;  t: [B] fp32;  y: [B,vec_in_dim] fp32;  entities / aatype / ...: int64.
 */
#ifndef LAMSLIDE_H_
#define LAMSLIDE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LAMSLIDE_ABI_VERSION 1

typedef enum {
  LAMSLIDE_OK = 0,
  LAMSLIDE_ERR_INVALID = -1,     /* bad argument / unsupported shape (reference: ValueError / assert) */
  LAMSLIDE_ERR_MISSING = -2,     /* a state-dict key is missing or has the wrong shape */
  LAMSLIDE_ERR_CUDA = -3,        /* CUDA runtime / driver error */
  LAMSLIDE_ERR_WORKSPACE = -4    /* workspace too small */
} lamslide_status;

/* One named host tensor of a state dict (fp32, contiguous). */
typedef struct {
  const char* name;
  const float* data;
  int32_t ndim;
  int64_t shape[4];
} lamslide_tensor;

/* ---- second stage: LatentSIV3 (src/models/components/latent/latent_si_v31.py:66-188) ------------------------------ */
typedef struct {
  int32_t depth, in_dim, hidden_size, num_heads;
  int32_t mlp_hidden;   /* int(hidden_size * mlp_ratio) */
  int32_t vec_in_dim;   /* 0 = no vec_in (y must be NULL) */
  int32_t normalize;    /* F.layer_norm after the input embedding (latent_si_v31.py:173-174) */
  float theta;          /* RoPE base (10000) */
} lamslide_backbone_config;

typedef struct lamslide_backbone lamslide_backbone;

/* LatentSIV3.__init__ + load_state_dict: packs the weights.  Keys: x_in.*, cond_to_emb.*, mask_to_emb.weight,
 * time_in.{in,out}_layer.*, [vec_in.*], blocks.{i}.modulation.lin.*, blocks.{i}.{spatial,temporal}_block.{linear1,linear2}.*,
 * blocks.{i}.*.norm.{query,key}_norm.scale, adaLN_modulation.1.*, linear.*  */
int lamslide_backbone_create(const lamslide_backbone_config* cfg, const lamslide_tensor* tensors, int32_t n_tensors,
                             lamslide_backbone** out);
void lamslide_backbone_destroy(lamslide_backbone* h);
size_t lamslide_backbone_workspace_bytes(const lamslide_backbone* h, int32_t B, int32_t T, int32_t L);

/* LatentSIV3.forward(x, t, x_cond, x_cond_mask, y=None) -> [B,T,L,in_dim]  (latent_si_v31.py:168-188). */
int lamslide_backbone_forward(lamslide_backbone* h, const float* x, const float* t, const float* x_cond,
                              const int64_t* x_cond_mask, const float* y, float* out, int32_t B, int32_t T, int32_t L,
                              void* workspace, size_t workspace_bytes, void* stream);

/* Sampler(transport).get_sample_fn("ODE", {"sampling_method": "euler", "num_steps": n})(init, model, **kwargs)
 * (transport.py:365-411, 475-503; integrators.py:84-120; torchdiffeq fixed-grid Euler) fused with Transport.get_drift
 * (transport.py:158-202).  path_type: 0 Linear, 1 GVP.  prediction: 0 velocity, 1 data, 2 noise, 3 score.
 * The time interval follows Transport.check_interval (transport.py:69-101): [0,1] for velocity models, [1e-3, 1-1e-3] otherwise.
 * x holds the initial noise on entry and the final state on exit.  states_out (nullable): [num_steps,B,T,L,D] like the
 * reference's return value; velocities_out (nullable): [num_steps-1,B,T,L,D] drift evaluations (parity tests). */
/* workspace of lamslide_ode_sample for `num_steps` grid points (>= lamslide_backbone_workspace_bytes: the per-sample vector path of
 * up to 16 evaluations is computed in one pass, since the time grid is known in advance). */
size_t lamslide_ode_workspace_bytes(const lamslide_backbone* h, int32_t B, int32_t T, int32_t L, int32_t num_steps);
int lamslide_ode_sample(lamslide_backbone* h, float* x, const float* x_cond, const int64_t* x_cond_mask, const float* y,
                        int32_t path_type, int32_t prediction, int32_t num_steps, float* states_out, float* velocities_out,
                        int32_t B, int32_t T, int32_t L, void* workspace, size_t workspace_bytes, void* stream);

/* One explicit Euler step of the probability-flow ODE for an arbitrary model callable (the generic path of
 * Sampler.sample_ode when the model is not a lamslide_backbone): v = drift(net_out, x, t) (transport.py:158-202),
 * x <- x + (t_next - t) * v (torchdiffeq fixed-grid Euler).  velocity_out nullable.  numel = B*T*L*D (multiple of 4). */
int lamslide_euler_step(float* x, const float* net_out, int32_t path_type, int32_t prediction, float t, float t_next,
                        float* velocity_out, int64_t numel, void* stream);

/* SecondStageCondLightningBase.setup_conditioning (lightning_base.py:240-263). */
int lamslide_setup_conditioning(const float* latents, float* x_cond, int64_t* x_cond_mask, int32_t B, int32_t T, int32_t L,
                                int32_t D, int32_t cond_begin, int32_t cond_end, int32_t mask_cond_mean, void* stream);

/* out = px * x + pm * m + pw * w over numel floats (m, w nullable; out may alias x; numel % 4 == 0, 16-byte aligned): the building
 * block of the SDE sampler's steps — sde.__Euler_Maruyama_step / __Heun_step (integrators.py:29-52) and the last step of
 * Sampler.sample_sde (transport.py:266-299) are linear in (state, network output, noise) with time-only coefficients. */
int lamslide_lincomb3(float* out, const float* x, const float* m, const float* w, float px, float pm, float pw, int64_t numel,
                      void* stream);

/* Building blocks of the ODE integrators other than fixed-grid Euler — Sampler.sample_ode hands its drift to torchdiffeq.odeint with
 * sampling_method "dopri5" by default (transport.py:365-411, integrators.py:103-120, configs/eval_peptide.yaml:21-23); the host
 * side (lam_slide_b200/odeint.py) restates torchdiffeq's adaptive Runge-Kutta controller around these:
 *   lincomb_n:       out = sum_j coefs[j] * srcs[j]  (1 <= n <= 8; srcs: HOST array of device pointers, coefs: HOST array; out may be
 *                    one of the sources) — stage combinations y0 + sum_j (dt beta_ij) k_j, solution, mid-point, interpolant;
 *   rk_error_sumsq:  *acc (device double, caller zeroes it) += sum_i ((sum_j coefs[j] srcs[j][i]) / (atol + rtol max(|a_i|, |b_i|)))^2
 *                    — the squared RMS error norm of an embedded step (torchdiffeq rk_common.py: _compute_error_ratio), fused. */
int lamslide_lincomb_n(float* out, const float* const* srcs, const float* coefs, int32_t n, int64_t numel, void* stream);
int lamslide_rk_error_sumsq(const float* const* srcs, const float* coefs, int32_t n, const float* a, const float* b, double rtol,
                            double atol, int64_t numel, double* acc, void* stream);

/* K-sample evaluation metrics on the device.  preds [K, B, T, A, D] = K batched sample() results restricted
 * to the frames after the conditioning window, target [B, T, A, D]; err = L2 norm over D.
 *   mode 0 - Wrapper.test_step / _compute_errors of second_stage/nba.py:161-238 and pedestrian.py:149-226 (unclustered metric):
 *            ades, fdes [B * A]: min over the first num_runs samples of the time-averaged / final-frame error of every agent
 *            (the caller drops padded agents with batch["attention_mask"][:, -1], as the reference does before stacking);
 *   mode 1 - Wrapper.test_step of second_stage/md17.py:139-171: ades, fdes [B]: mean over the K samples of the error averaged
 *            over frames and atoms / over the atoms of the final frame. */
int lamslide_ksample_errors(const float* preds, const float* target, float* ades, float* fdes, int32_t K, int32_t num_runs,
                            int32_t B, int32_t T, int32_t A, int32_t D, int32_t mode, void* stream);

/* ---- first stage: BackboneBase + Encoder / Decoder / DecoderQuerySplitter ------------------------------------------- */
typedef enum { LAMSLIDE_FS_PEPTIDE = 0, LAMSLIDE_FS_MD17 = 1, LAMSLIDE_FS_NBA = 2, LAMSLIDE_FS_PEDESTRIAN = 3 } lamslide_fs_kind;

#define LAMSLIDE_MAX_OUTPUTS 4
typedef struct {
  int32_t kind;              /* lamslide_fs_kind: selects Backbone.prepare_inputs (first_stage/{peptide,md17,nba,pedestrian}.py) */
  int32_t dim_input, dim_latent, num_entities, entity_dim;
  int32_t qk_norm;
  /* Encoder (encoder.py:44-103) */
  int32_t enc_num_latents, enc_heads_cross, enc_dim_head_cross, enc_heads_latent, enc_dim_head_latent;
  int32_t enc_blocks_cross, enc_blocks_attn;
  /* Decoder / DecoderQuerySplitter (decoder.py:12-102, 313-411) */
  int32_t dec_query_splitter, dec_num_split, dec_dim_query;
  int32_t dec_heads_cross, dec_dim_head_cross, dec_heads_latent, dec_dim_head_latent;
  int32_t dec_blocks_cross, dec_blocks_attn;
  int32_t n_outputs;
  const char* output_names[LAMSLIDE_MAX_OUTPUTS];
  int32_t output_dims[LAMSLIDE_MAX_OUTPUTS];
  /* dataset specifics */
  int32_t max_res;       /* peptide: SinCosPositionalEmbedding1D positions */
  int32_t n_atom_types;  /* md17 */
} lamslide_first_stage_config;

typedef struct lamslide_first_stage lamslide_first_stage;

/* Per-frame inputs of Backbone.encode (lightning_base.py:37-40); frames = B*T, N entities per frame (zero padded).
 *   pos      : peptide atom14_pos [F,N,42]; md17 pos [F,N,3]; nba / pedestrian pos [F,N,2]   (fp32)
 *   index0   : peptide aatype; md17 atom; nba team; pedestrian NULL                          (int64 [F,N])
 *   index1   : nba group; otherwise NULL                                                      (int64 [F,N])
 *   entities : entity ids                                                                     (int64 [F,N])
 *   mask     : attention_mask, 1 = valid key (bool / uint8 [F,N]) or NULL (peptide passes None — peptide.py:79) */
typedef struct {
  const float* pos;
  const int64_t* index0;
  const int64_t* index1;
  const int64_t* entities;
  const uint8_t* mask;
} lamslide_frame_inputs;

int lamslide_first_stage_create(const lamslide_first_stage_config* cfg, const lamslide_tensor* tensors, int32_t n_tensors,
                                lamslide_first_stage** out);
void lamslide_first_stage_destroy(lamslide_first_stage* h);
size_t lamslide_first_stage_workspace_bytes(const lamslide_first_stage* h, int32_t frames, int32_t N);

/* FirstStageLightningBase.encode(batch) -> latents [F, num_latents, dim_latent] (lightning_base.py:155-156). */
int lamslide_encode(lamslide_first_stage* h, const lamslide_frame_inputs* in, float* latents_out, int32_t frames, int32_t N,
                    void* workspace, size_t workspace_bytes, void* stream);
/* FirstStageLightningBase.decode(latents, entities) -> {name: [F, N, out_dim]} (lightning_base.py:149-153);
 * outs[i] receives output i in the order of cfg.output_names (NULL entries are skipped). */
int lamslide_decode(lamslide_first_stage* h, const float* latents, const int64_t* entities, float* const* outs, int32_t frames,
                    int32_t N, void* workspace, size_t workspace_bytes, void* stream);

/* ---- misc ------------------------------------------------------------------------------------------------------------ */
int lamslide_abi_version(void);
const char* lamslide_last_error(void);
/* number of kernels this library launched on the calling thread since the last reset (bench.py's gpu_launches). */
int64_t lamslide_launch_count(int32_t reset);

/* test hook: launches of one named kernel family on the calling thread since the last reset ("attn_tc", "attn_seq", "attn_flash",
 * "attn_rows", "attn_small", ...); reset != 0 clears ALL named counters after reading. */
int64_t lamslide_debug_kernel_count(const char* name, int32_t reset);

/* profiling aid: device buffer (148 x 32 x 8 bytes) that the tcgen05 attention kernel's trace variants (lamslide_debug_attention mode
 * 3 + 4 * 5 / 3 + 4 * 6) fill with the cycles its MMA warp and one softmax warp spend per kind of wait / work.  NULL turns it off. */
void lamslide_debug_attention_trace(void* device_buffer);

/* Optional per-kernel-class timing with CUDA events on the launching stream (bench.py's roofline and time shares).
 * begin() arms it for the calling thread; end() synchronises the device and writes a JSON object
 * {"gemm_linear1": {"ms": .., "launch_groups": ..}, ...} into json_out.  Not for use under CUDA-graph capture. */
int lamslide_profile_begin(void);
int lamslide_profile_end(char* json_out, size_t json_bytes);

/* test hooks (tests/ only): the tcgen05 GEMM in isolation, C[M,N] = A[M,K] B[N,K]^T (+ bias); A, B device bf16, C device fp32. */
int lamslide_debug_gemm(const void* a_bf16, const void* b_bf16, const float* bias, float* c, int32_t M, int32_t N, int32_t K,
                        int32_t block_n, void* stream);
/* the attention kernels in isolation on a token-major qkv buffer [tokens, 3H] bf16 -> out [tokens, ldo] bf16.
 * temporal != 0: sequences over T (stride L); else over L.  mode: 0 = automatic choice, 1 = streaming flash kernel
 * (running maximum), 2 = whole-sequence kernel (K/V resident in shared memory, no running maximum), 3 = tcgen05 kernel
 * (S = Q K^T and O = P V on the 5th-gen tensor cores, accumulators in TMEM; 3 + 4 v selects variant v: 1 - 4 exponential mixes of
 * attn_tc.cuh (all MUFU, 2 / 8, 4 / 8 polynomial, none), 5 / 6 its cycle trace, 7 - 10 the three-group kernel of attn_tc3.cuh). */
int lamslide_debug_attention(const void* qkv_bf16, void* out_bf16, int32_t B, int32_t T, int32_t L, int32_t H, int32_t heads,
                             int32_t ldo, int32_t temporal, int32_t mode, void* stream);
/* linear1 of one ParallelMLPAttentionV2 block with its fused epilogue (mmdit.py:241-247: bias, QK-RMSNorm, RoPE, q pre-scale,
 * erf-GELU) in isolation: u [rows,H] bf16, w1 [3H+M,H] bf16 -> qkv [rows,3H] bf16, act[:, H:] of [rows,H+M] bf16.
 * The rope position of a row is (row / pos_div) % pos_mod.  legacy: 0 persistent kernel, 1 one-tile-per-CTA kernel,
 * 2 / 3 profiling aids of the persistent kernel (2: epilogue math without the global stores, 3: stores without the math);
 * + 16: run the persistent kernel as single CTAs instead of 2-CTA clusters that multicast the weight tiles;
 * + 64: the spatial attention fused into the epilogue (sequences of pos_mod consecutive rows, pos_div = 1): only the q | k | v columns
 *       are computed, qkv is not written, act[:, :H] receives softmax(q k^T / sqrt(hd)) v (mmdit.py:42-55, 240-249). */
int lamslide_debug_linear1(const void* u_bf16, const void* w1_bf16, const float* bias, const float* q_scale, const float* k_scale,
                           void* qkv_bf16, void* act_bf16, int32_t rows, int32_t H, int32_t M, int32_t heads, int32_t pos_div,
                           int32_t pos_mod, float theta, int32_t legacy, void* stream);
/* linear2 + gated residual (mmdit.py:248; latent_si_v31.py:54,61) in isolation: h [rows,H] fp32 += gate[row / rows_per_sample]
 * * (act [rows,H+M] bf16 . w2 [H,H+M]^T + bias).  gate: [n_samples, H] fp32. */
int lamslide_debug_linear2(const void* act_bf16, const void* w2_bf16, const float* bias, const float* gate, float* h, int32_t rows,
                           int32_t H, int32_t M, int32_t rows_per_sample, int32_t legacy, void* stream);

/* the TMA + tcgen05 main loop of the persistent GEMM alone (no epilogue work, nothing stored): profiling aid that
 * separates "operand feed + tensor pipe" from "epilogue" for a shape.  |block_n| in {192, 128, 64}; negative: single CTAs
 * instead of 2-CTA clusters with weight-tile multicast. */
int lamslide_debug_gemm_mainloop(const void* a_bf16, const void* b_bf16, int32_t rows, int32_t N, int32_t K, int32_t block_n,
                                 void* stream);

/* the fused "MLP half of linear1 + GELU + linear2 + gated residual" kernel in isolation (mmdit.py:241-248, latent_si_v31.py:54):
 * h [rows,H] fp32 += gate[row / rows_per_sample] * ([act[:, :H] | gelu(u w1[3H:]^T + b1[3H:])] w2^T + b2);
 * u [rows,H], act [rows,H+M], w1 [3H+M,H], w2 [H,H+M] device bf16; b1 [3H+M], b2 [H], gate [n_samples,H] device fp32. */
int lamslide_debug_fused_mlp(const void* u_bf16, const void* act_bf16, const void* w1_bf16, const void* w2_bf16, const float* b1,
                             const float* b2, const float* gate, float* h, int32_t rows, int32_t H, int32_t M,
                             int32_t rows_per_sample, void* stream);

/* the same kernel with the drain that also applies the NEXT block's pre-norm + modulate (latent_si_v31.py:50,57) to the rows it has
 * just updated: u_out [rows,H] bf16 = LayerNorm(h_new) * (1 + ln_scale[b]) + ln_shift[b]  (no affine, eps 1e-6; ln_shift / ln_scale
 * [n_samples,H] device fp32; u_out may be the buffer u_bf16 points to). */
int lamslide_debug_fused_mlp_ln(const void* u_bf16, const void* act_bf16, const void* w1_bf16, const void* w2_bf16, const float* b1,
                                const float* b2, const float* gate, float* h, int32_t rows, int32_t H, int32_t M,
                                int32_t rows_per_sample, const float* ln_shift, const float* ln_scale, void* u_out, void* stream);

/* one linear layer of the first stage in isolation: y [rows, N] (pitch ldy) = epi(x [rows, K] (pitch ldx) . w^T + bias), with the
 * epilogue options of the first-stage layers (act: 0 none, 1 erf-GELU, 2 SiLU of the sum; + rowadd[row % period] ; + res).
 * w [N, K] and bias [N] are HOST fp32 (packed as lamslide_first_stage_create packs them), the rest device fp32.
 * path: 0 = the kernel the product picks for this shape (tcgen05 3xTF32 when K % 4 == 0), 1 = the mma.sync / FMA kernels. */
int lamslide_debug_fs_linear(const float* x, const float* w_host, const float* bias_host, float* y, int32_t rows, int32_t N, int32_t K,
                             int32_t ldx, int32_t ldy, int32_t act, const float* res, int32_t ldr, const float* rowadd, int32_t period,
                             int32_t ldra, int32_t path, void* stream);

/* the same layer followed by the LayerNorm of its result (eps 1e-5; the successor's pre-norm, run in the layer's epilogue where a
 * thread or thread pair holds a group): ln_out [rows,N] = LN over groups of `group` columns (0: the whole row) [* ln_w + ln_b];
 * res [rows,N] optional residual; y_needed == 0: y need not be written. */
int lamslide_debug_fs_linear_ln(const float* x, const float* w_host, const float* bias_host, float* y, float* ln_out, const float* ln_w,
                                const float* ln_b, int32_t group, int32_t rows, int32_t N, int32_t K, const float* res,
                                int32_t y_needed, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LAMSLIDE_H_ */

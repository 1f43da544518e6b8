/* vaw_b200.h — C ABI of libvaw_b200.so: the B200 (sm_100a) kernels behind the diffusion training step of
 * LilYau350/Variance-Aware-Weight (GaussianDiffusion.training_losses + timestep importance sampling + DiT/U-ViT
 * denoiser forward/backward).
 *
 * Conventions
 *   - every entry point returns 0 (VAW_OK) or a negative code; vaw_last_error() returns the message (thread-local)
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless the comment says "host"
 *   - no allocation inside the library: the caller owns every buffer (the Python host side uses torch's allocator)
 *   - every kernel-launching call takes a cudaStream_t (passed as void*) and is asynchronous
 *   - the library holds no mutable global state besides caches of immutable device properties
 *
 * Each entry cites the reference interface (file:line under /root/reference) that it replaces.
 */
#ifndef VAW_B200_H_
#define VAW_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VAW_OK 0
#define VAW_ERR_INVALID (-1)
#define VAW_ERR_CUDA (-2)
#define VAW_ERR_UNSUPPORTED (-3)

typedef void* vaw_stream_t; /* cudaStream_t */

/* ModelMeanType codes = the reference enum values (tools/gaussian_diffusion.py:21-32, enum.auto() from 1) */
#define VAW_MEAN_PREVIOUS_X 1
#define VAW_MEAN_START_X 2
#define VAW_MEAN_EPSILON 3
#define VAW_MEAN_VELOCITY 4
#define VAW_MEAN_VECTOR 5
#define VAW_MEAN_SCORE 6

/* weight_type families of compute_mse_loss_weight (tools/gaussian_diffusion.py:1092-1148) */
#define VAW_W_CONSTANT 0
#define VAW_W_LAMBDA 1
#define VAW_W_MIN_SNR 2 /* min_snr_<k> */
#define VAW_W_MAX_SNR 3 /* max_snr_<k> */
#define VAW_W_DEBIAS 4
#define VAW_W_MIN_DEBIAS 5
#define VAW_W_MAX_DEBIAS 6
#define VAW_W_P2 7
#define VAW_W_TRUNC_SNR 8
#define VAW_W_SNR 9
#define VAW_W_INV_SNR 10

/* dtype codes */
#define VAW_F32 0
#define VAW_BF16 1

/* ---- library -------------------------------------------------------------------------------------------- */
const char* vaw_last_error(void);
int vaw_version(void);
int vaw_device_check(void); /* 0 iff the current device is compute capability 10.x */

/* ---- K1: fused q_sample + target -------------------------------------------------------------------------
 * Replaces GaussianDiffusion.q_sample (tools/gaussian_diffusion.py:234-252), compute_target (:818-832) and the
 * _extract_into_tensor gathers (:1059-1072).  x_t = fl(fl(a x0) + fl(s eps)), bit-exact with the reference fp32
 * path.  t != NULL: tab_* are [T] fp32 tables gathered by t[n] (int64).  t == NULL: tab_* are per-sample [N]
 * arrays (FlowMatching.q_sample, :1273-1277).  tab_c0/tab_c1: posterior_mean_coef1/2 (PREVIOUS_X) or
 * d_alpha/d_sigma (VECTOR); may be NULL otherwise.  target may be NULL (EPSILON / START_X need no tensor). */
/* sample_from_latent (tools/trainer.py:21-25), the step before the path: latent [N, 2C, H, W] holds
 * (mean | std) along channels; out[N, C, H, W] = (mean + std * eps) * scale, bit-exact with the reference's fp32 ops. */
int vaw_sample_from_latent(const float* latent, const float* eps, float* out, long long N, long long chw, float scale,
                           vaw_stream_t stream);
int vaw_qsample_target(const float* x0, const float* noise, const long long* t, const float* tab_alpha,
                       const float* tab_sigma, const float* tab_c0, const float* tab_c1, float* x_t, float* target,
                       int mean_type, long long N, long long chw, vaw_stream_t stream);

/* ---- K1 / K2 with the noise drawn in-kernel -------------------------------------------------------------
 * Bit-compatible with the tensors torch.randn_like / torch.randint produce from the same CUDA generator state (seed,
 * offset): the kernels walk ATen's Philox4x32-10 subsequence / element mapping (DistributionTemplates.h), so the noise
 * of training_losses (tools/gaussian_diffusion.py:849-850), the timesteps of sample_t (:810-816) and the draw of
 * sample_from_latent (tools/trainer.py:21-25) never exist as tensors in HBM.  The caller advances the generator by
 * vaw_philox_offset_increment(numel) per draw, exactly what ATen's kernels would have consumed.
 * vaw_qsample_philox: x0 [N, chw], or latent [N, 2 chw] (mean | std) from which x0 = (mean + std * eps1) * latent_scale
 * is rebuilt with the draw at offset_latent; noise at offset_noise; x_start_out / noise_out / target nullable. */
int vaw_philox_offset_increment(long long numel, unsigned long long* increment);
int vaw_qsample_philox(const float* x0, const float* latent, float latent_scale, unsigned long long seed,
                       unsigned long long offset_latent, unsigned long long offset_noise, const long long* t,
                       const float* tab_alpha, const float* tab_sigma, const float* tab_c0, const float* tab_c1,
                       float* x_start_out, float* noise_out, float* x_t, float* target, int mean_type, long long N,
                       long long chw, vaw_stream_t stream);
/* out[i] = low + (curand4 value % (high - low)), the stream of torch.randint(low, high, (n,), device=cuda) */
int vaw_randint_philox(unsigned long long seed, unsigned long long offset, long long low, long long high, long long* out,
                       long long n, vaw_stream_t stream);
/* K2 (vaw_wmse_fwd_bwd_strided) that re-draws each noise element instead of reading a noise tensor; x0 may be NULL for
 * the EPSILON / SCORE targets. */
int vaw_wmse_fwd_bwd_philox(const void* out, int out_dtype, long long out_stride, const float* x0, unsigned long long seed,
                            unsigned long long offset_noise, const long long* t, const float* tab_alpha,
                            const float* tab_sigma, const float* tab_c0, const float* tab_c1, const float* w_tab,
                            float* mse, void* grad_out, long long grad_stride, float gscale, int mean_type, long long N,
                            long long chw, vaw_stream_t stream);

/* ---- K8: fused reverse-process step ------------------------------------------------------------
 * Replaces the elementwise tail of GaussianDiffusion.p_mean_variance (tools/gaussian_diffusion.py:278-384) followed
 * by p_sample (:455-506), ddim_sample (:603-651) or ddim_reverse_sample (:653-689), and the 8 per-call table uploads
 * of _extract_into_tensor (:1059-1072).  model_out is the denoiser output, fp32 or bf16, out_stride values per sample
 * (chw, or 2*chw with the variance channels behind the mean channels for the LEARNED* variance types, :312-314).
 * tab is a [VAW_RT_ROWS][T] fp32 table (each float64 schedule table rounded once to fp32, as the reference does).
 * Outputs are fp32 [N, chw]; any of them may be NULL (at least one must not be).  mode VAW_RS_MOMENTS only
 * evaluates p_mean_variance.  mean_type / var_type use the reference's enum values (:21-45).  Each fp32 operation
 * is rounded separately in the reference's order: results are bit-identical to the eager path except where exp()
 * is involved (device vs host libm, <= 1 ulp).  VELOCITY uses the per-sample coefficient the reference intends
 * (its :392-397 broadcasts over the wrong axis and only runs for N == 1 or N == W).  t outside [0, T) is clamped
 * (the reference's gather would device-assert); no other validation of device data is done. */
enum { VAW_VT_LEARNED = 1, VAW_VT_FIXED_SMALL = 2, VAW_VT_FIXED_LARGE = 3, VAW_VT_LEARNED_RANGE = 4 };
enum { VAW_RS_DDPM = 0, VAW_RS_DDIM = 1, VAW_RS_DDIM_REVERSE = 2, VAW_RS_MOMENTS = 3 };
enum { VAW_RT_SQRT_RECIP_AC = 0, VAW_RT_SQRT_RECIPM1_AC, VAW_RT_SQRT_AC, VAW_RT_SQRT_1MAC, VAW_RT_INV_COEF1,
       VAW_RT_COEF2_OVER_COEF1, VAW_RT_COEF1, VAW_RT_COEF2, VAW_RT_LOGVAR /* LEARNED_RANGE: min_log */,
       VAW_RT_MAX_LOG, VAW_RT_VARIANCE, VAW_RT_AC, VAW_RT_AC_PREV, VAW_RT_AC_NEXT,
       VAW_RT_TRUE_LOGVAR /* posterior_log_variance_clipped whatever the variance type (vaw_vb_terms) */, VAW_RT_ROWS };
int vaw_reverse_step(const void* model_out, int out_dtype, long long out_stride, const float* x, const float* noise,
                     const long long* t, const float* tab, int T, float* sample, float* pred_xstart, float* mean,
                     float* log_variance, float* variance, int mean_type, int var_type, int mode, float eta, int clip,
                     long long N, long long chw, vaw_stream_t stream);
/* IntervalCFG combine (tools/sampler.py:46-48): both = [cond | uncond] halves of a doubled batch (half values each),
 * y = uncond + scale * (cond - uncond), every op rounded in the tensor's dtype like the eager expression. */
int vaw_cfg_combine(const void* both, void* y, int dtype, float scale, long long half, vaw_stream_t stream);

/* ---- EDM sampler (tools/cfg_edm.py; the sampler of the shipped recipe, run.sh `--solver heun`) -------------------------
 * The state is float64 [n] like the reference's; all coefficients are per-step scalars evaluated on the host.
 * vaw_edm_pre (ablation_sampler :193-194 + Net.forward :52,58): x_hat = a x_cur + c noise (noise nullable: x_hat = a x_cur),
 * x_in = c_in * float32(x_hat / s_hat) is the denoiser input.  x_hat may be NULL (only x_in wanted). */
int vaw_edm_pre(const double* x_cur, const double* noise, double a, double c, double s_hat, float c_in, double* x_hat,
                float* x_in, long long n, vaw_stream_t stream);
/* vaw_edm_post (Net.forward :62-77 + ablation_sampler :197-207): out = denoiser output (fp32 / bf16, out_stride values per
 * sample, the first chw are used), x_src = the float64 state the denoiser saw (x_hat or x_prime), s_src = s(t) of it.
 *   denoised = c_skip x + c_out out (float32; pred VAW_MEAN_START_X: denoised = out)   d = A x_src - Bc denoised (float64)
 *   mode -1: x_in_next = denoised (Net.forward alone)      mode 0 (Euler / last step): x_out = x_hat + h d
 *   mode 1 (Heun predictor): d_out = d, x_out = x_prime = x_hat + h d (h = alpha h), x_in_next = c_in_next *
 *           float32(x_prime / s_next) - the next denoiser input
 *   mode 2 (Heun corrector): x_out = x_hat + h (c1 d_prev + c2 d)
 * Every product / sum / quotient is rounded separately in the reference's order. */
int vaw_edm_post(const void* out, int out_dtype, long long out_stride, const double* x_src, double s_src, float c_skip,
                 float c_out, int pred, double A, double Bc, int mode, double h, const double* x_hat, const double* d_prev,
                 double c1, double c2, double* x_out, double* d_out, double s_next, float c_in_next, float* x_in_next,
                 long long N, long long chw, vaw_stream_t stream);
/* ---- flow-matching SDE sampler (tools/gaussian_diffusion.py:1206-1257 conversions, :1366-1408 sde_sample) ---------------
 * coef: HOST array {alpha_t, sigma_t, d_alpha_t, d_sigma_t, 2 sigma_t d_sigma_t} (float32) of the step's time.
 * drift = vector - 0.5 diffusion score from the model output `out` evaluated at x_eval (mean_type START_X .. VECTOR).
 *   noise term = noise_scale * noise * sqrt_abs_step, noise_scale = sqrt(diffusion) of the step's CURRENT time
 *   mode 0: x_out = x_base + drift step (+ noise term when noise != NULL)                           [Euler / last step]
 *   mode 1: same and drift_out = drift                                                              [Heun predictor]
 *   mode 2: x_out = x_base + 0.5 (drift_prev + drift) step + noise term                             [Heun corrector]   */
int vaw_flow_sde_step(const void* out, int out_dtype, const float* x_eval, const float* coef, int mean_type, int mode,
                      const float* x_base, const float* drift_prev, const float* noise, float step, float noise_scale,
                      float sqrt_abs_step, float* x_out, float* drift_out, long long n, vaw_stream_t stream);

/* ---- K2: fused weighted-MSE forward + backward --------------------------------------------------------------
 * Replaces (target - out)**2 -> mean_flat (tools/nn.py:86-90) -> w * raw (tools/gaussian_diffusion.py:911-913)
 * and the autograd backward of that chain (seeded by trainer.py:107-108).  The target is rebuilt from x0/noise
 * in-kernel.  w_tab: [T] LUT gathered by t (or [N] per-sample when t == NULL; NULL = weight 1).
 *   mse[n]  = w_n * mean_i (target - out)^2            raw_mse[n] = the unweighted mean (nullable)
 *   grad_out[n,i] = gscale * gscale_n[n] * w_n * 2 (out - target) / chw   (nullable; same dtype as out)        */
int vaw_wmse_fwd_bwd(const void* out, int out_dtype, const float* x0, const float* noise, const long long* t,
                     const float* tab_alpha, const float* tab_sigma, const float* tab_c0, const float* tab_c1,
                     const float* w_tab, float* mse, float* raw_mse, void* grad_out, const float* gscale_n,
                     float gscale, int mean_type, long long N, long long chw, vaw_stream_t stream);

/* Same with the model output / gradient rows `out_stride` / `grad_stride` values apart (>= chw): the mean channels of a
 * [N, 2C, H, W] learned-variance output (tools/gaussian_diffusion.py:891 th.split) are read in place. */
int vaw_wmse_fwd_bwd_strided(const void* out, int out_dtype, long long out_stride, const float* x0, const float* noise,
                             const long long* t, const float* tab_alpha, const float* tab_sigma, const float* tab_c0,
                             const float* tab_c1, const float* w_tab, float* mse, float* raw_mse, void* grad_out,
                             long long grad_stride, const float* gscale_n, float gscale, int mean_type, long long N,
                             long long chw, vaw_stream_t stream);

/* ---- variational-bound term (learned variance / KL objectives) ----------------------------------------------
 * Replaces GaussianDiffusion._vb_terms_bpd (tools/gaussian_diffusion.py:775-808: q_posterior_mean_variance :254-276,
 * p_mean_variance :278-384 with clip_denoised=False, normal_kl and discretized_gaussian_log_likelihood of
 * tools/losses.py:12-77, mean_flat / ln 2, th.where(t == 0, nll, kl)) as used by training_losses (:862-875 LossType.KL /
 * RESCALED_KL; :886-906 the term added to the MSE objective with the mean prediction detached) and its autograd
 * backward.  out: model output, rows out_stride apart, mean channels [chw] followed (LEARNED / LEARNED_RANGE) by the
 * variance channels [chw].  tab: the [VAW_RT_ROWS][T] table of vaw_reverse_step.  vb[n] = out_scale * bound in bits.
 * grad (nullable, layout of out, rows grad_stride apart): gscale * gscale_n[n] * d vb[n] / d out; with detach_mean the
 * mean channels are left untouched (the caller's MSE gradient lives there).  mean_type PREVIOUS_X..VELOCITY. */
int vaw_vb_terms(const void* out, int out_dtype, long long out_stride, const float* x0, const float* x_t,
                 const long long* t, const float* tab, int T, float* vb, void* grad, long long grad_stride,
                 const float* gscale_n, float gscale, int mean_type, int var_type, int detach_mean, float out_scale,
                 long long N, long long chw, vaw_stream_t stream);

/* y[n,:] = x[n,:] * s[n] — applies a late per-sample upstream gradient to K2's grad_out */
int vaw_scale_rows(const void* x, const float* s, void* y, int dtype, long long N, long long chw,
                   vaw_stream_t stream);

/* HOST function: per-timestep loss-weight LUT, compute_mse_loss_weight (tools/gaussian_diffusion.py:1092-1148)
 * evaluated with the reference's fp32 operation order.  sqrt_ac/sqrt_1mac/lut are HOST arrays of length T.
 * Returns VAW_ERR_INVALID for (mean_type, weight_kind) pairs the reference rejects with ValueError (:1144-1145). */
int vaw_loss_weight_lut(const double* sqrt_ac, const double* sqrt_1mac, int T, int mean_type, int weight_kind,
                        double k, double p2_k, double p2_gamma, float* lut);

/* ---- K3: timestep importance sampling ------------------------------------------------------------------------
 * Replaces ScheduleSampler.sample (tools/resample.py:43-59) and LossSecondMomentResampler.weights (:142-149),
 * bit-exact with numpy (pairwise sums, sequential cumsum, searchsorted side='right').
 * mode 0: explicit weights w_in[T] (fp64).  mode 1: weights from history[T,H] (fp64) + counts[T] (int32).
 * u[B]: uniform doubles drawn on the host from numpy's global MT19937 (np.random.random_sample).
 * Outputs idx[B] int64, imp_w[B] fp32; w_out/p_out/cdf_out [T] fp64 are optional (NULL to skip).            */
int vaw_sampler_sample(int mode, const double* w_in, const double* history, const int* counts, int T, int H,
                       double uniform_prob, const double* u, long long B, long long* idx, float* imp_w,
                       double* w_out, double* p_out, double* cdf_out, vaw_stream_t stream);

/* Replaces LossSecondMomentResampler.update_with_all_losses (tools/resample.py:151-159): applies n_total gathered
 * (t, loss) entries in order; entries with t < 0 are padding.                                                  */
int vaw_sampler_update(double* history, int* counts, const int* ts, const float* losses, long long n_total, int T,
                       int H, vaw_stream_t stream);

/* Packs (int64 t, fp32 loss)[B] into int32/fp32 arrays padded to Bpad (t = -1) for the single all_gather that
 * replaces the three collectives of update_with_local_losses (tools/resample.py:85-106).                       */
int vaw_pack_tloss(const long long* t, const float* loss, int* t32, float* l32, long long B, long long Bpad,
                   vaw_stream_t stream);

/* ---- K4: tcgen05 GEMM with fused epilogues ---------------------------------------------------------------------
 * D[M,N] = A[M,K] B[N,K]^T, bf16 operands, fp32 accumulation in TMEM.  Replaces the nn.Linear GEMMs of the DiT /
 * U-ViT blocks (models/dit.py:118-137 via timm Attention/Mlp; models/uvit.py:55-121) in forward, dgrad and wgrad. */
#define VAW_EPI_BF16 0       /* out(bf16) = acc + bias */
#define VAW_EPI_F32 1        /* out(f32)  = acc + bias (+= when accumulate).  With out2 != NULL ("row-sum form", no bias): the
                                last 32 columns of B are not part of the output - out is [M, N - 32] (ldo, default N - 32) and
                                column N - 32 of the product goes to out2[M] (fp32, += when accumulate); with
                                B = [X | 1 0 ... 0] (vaw_ln_fwd_ex ones_block) one GEMM yields a Linear's weight AND bias
                                gradient (models/dit.py:126-128 qkv / fc1 under autograd) */
#define VAW_EPI_GELU_TANH 2  /* out(bf16) = pre ; out2(bf16) = gelu_tanh(pre) */
#define VAW_EPI_GELU_ERF 3   /* out(bf16) = pre ; out2(bf16) = gelu_erf(pre) */
#define VAW_EPI_GATE_RES 4   /* out(bf16) = y   ; out2(f32) = resid + gate[row/rows_per_sample] * y */
#define VAW_EPI_RES 5        /* out2(f32) = resid + bf16(acc + bias) */
#define VAW_EPI_DGELU_TANH 6 /* out(bf16) = acc * gelu_tanh'(aux) */
#define VAW_EPI_DGELU_ERF 7  /* out(bf16) = acc * gelu_erf'(aux) */
#define VAW_EPI_SILU 8       /* out(bf16) = pre ; out2(bf16) = silu(pre) */
#define VAW_EPI_DSILU 9      /* out(bf16) = acc * silu'(aux) */

typedef struct vaw_gemm_args {
  const void* A; /* bf16; a_mn=0: [M,K] row-major (lda); a_mn=1: [K,M] row-major (lda) */
  const void* B; /* bf16; b_mn=0: [N,K] row-major (ldb); b_mn=1: [K,N] row-major (ldb) */
  long long lda, ldb;
  int a_mn, b_mn;
  int M, N, K;
  int epilogue;
  void* out;
  void* out2;
  const float* bias;  /* [N] or NULL */
  const float* resid; /* [M,N] fp32 (ldo) */
  const float* gate;  /* [M/rows_per_sample, N] fp32 (ldg) */
  const void* aux;    /* bf16 [M,N] (ldo) */
  long long ldo, ldg; /* 0 = N */
  int rows_per_sample;
  int accumulate;
  int tile_n;    /* 0 = auto; 128 / 192 / 256 */
  int resid_mod; /* > 0: resid is a [resid_mod, N] table indexed by row % resid_mod (pos_embed) */
  int k_splits;  /* VAW_EPI_F32 only. > 1: split every tile's K loop; -1: split only the partial last wave
                    ("tail split"); partials are folded in fixed order (deterministic) */
  float* split_ws;          /* fp32 scratch for the partial slabs (128 x tile_n each) */
  long long split_ws_elems; /* capacity of split_ws in floats (0 = unchecked) */
  int cta_group; /* 0 = auto, 1 = one CTA per 128-row tile, 2 = SM pair per 256-row tile (tcgen05 cta_group::2) */
} vaw_gemm_args;

int vaw_gemm_bf16(const vaw_gemm_args* args, vaw_stream_t stream);
/* Weight gradient of a Linear fed one row per SAMPLE (adaLN-Zero modulation, models/dit.py:118-124; TimestepEmbedder
 * :41-79): out[M, N] fp32 (ldo) (+)= A^T B with A [K, M] (lda) and B [K, N] (ldb) bf16, K = batch size.  All epilogue
 * (1 GFLOP for 32 MB of output at DiT-XL/2): warp-level MMAs out of shared memory, fragments stored directly.  Returns
 * VAW_ERR_UNSUPPORTED for operands that are not 16-byte aligned / lda, ldb not multiples of 8 (use vaw_gemm_bf16). */
int vaw_wgrad_smallk(const void* A, long long lda, const void* B, long long ldb, float* out, long long ldo, int M, int N,
                     int K, int accumulate, vaw_stream_t stream);

/* ---- K4: flash-style attention ------------------------------------------------------------------------------------
 * Replaces F.scaled_dot_product_attention inside timm Attention (models/dit.py:126) / models/uvit.py:72-75 and its
 * backward.  qkv: bf16 [B*T, 3*H*hd], feature order (3, H, hd); o: bf16 [B*T, H*hd]; lse2: fp32 [B, H, T] (log2
 * domain); head_dim 64 or 72.  dqkv has the layout of qkv.                                                          */
int vaw_attn_fwd(const void* qkv, void* o, float* lse2, int B, int T, int H, int head_dim, vaw_stream_t stream);
int vaw_attn_bwd(const void* qkv, const void* o, const void* d_o, const float* lse2, void* dqkv, int B, int T, int H,
                 int head_dim, vaw_stream_t stream);
/* Same, with a caller-provided fp32 scratch of B*H*T elements: the row sums Delta = rowsum(dO * O) are then produced by
 * a separate coalesced pass instead of inside every CTA's prologue (faster; used by the engines). NULL = as above. */
int vaw_attn_bwd_ws(const void* qkv, const void* o, const void* d_o, const float* lse2, void* dqkv, float* delta_ws,
                    int B, int T, int H, int head_dim, vaw_stream_t stream);

/* ---- K4: LayerNorm (+ adaLN modulate | affine) forward / backward, residual-branch backward, reductions ------------
 * Replace nn.LayerNorm + modulate (models/dit.py:24-25,122-124,133-137,151-155) and nn.LayerNorm(affine)
 * (models/uvit.py:103-110) with their autograd backward.  x fp32 [M, D]; y / dy bf16 [M, D]; shift/scale/gate are
 * rows of the adaLN output (row = sample, leading dimension ld_mod); weight/bias fp32 [D].
 * Backward kernels leave [groups, chunks, 2, D] fp32 partial sums in `part`; vaw_finish_* fold them in fixed order. */
int vaw_ln_fwd(const float* x, const float* shift, const float* scale, long long ld_mod, int rows_per_sample,
               const float* weight, const float* bias, void* y, float* mean, float* rstd, int M, int D, float eps,
               vaw_stream_t stream);
/* Residual update of the previous branch fused into the LayerNorm + modulate of the next one (models/dit.py:133-137:
 * x = x + gate.unsqueeze(1) * branch, then modulate(norm(x), shift, scale)), one pass over the row:
 *   x_out[r, :] = x[r, :] + gate[r / rows_per_sample, :] * branch[r, :]      (branch bf16 [M, D], gate fp32 row stride ld_gate)
 *   y[r, :]     = LN(x_out[r, :]) * (1 + scale[n, :]) + shift[n, :]          (bf16), mean / rstd [M] as vaw_ln_fwd */
int vaw_ln_fwd_res(const float* x, const void* branch, const float* gate, long long ld_gate, float* x_out,
                   const float* shift, const float* scale, long long ld_mod, int rows_per_sample, void* y, float* mean,
                   float* rstd, int M, int D, float eps, vaw_stream_t stream);
/* General form of the two above.  branch == NULL: no folded-in residual update (gate / x_out unused).  ldy: row stride of
 * y (0 = D).  ones_block != 0: the 32 bf16 after the D outputs of each row are set to [1, 0, ..., 0] (ldy >= D + 32), so
 * that a weight-gradient GEMM reading y as [M, D + 32] also returns the bias gradient (VAW_EPI_F32 with out2). */
int vaw_ln_fwd_ex(const float* x, const void* branch, const float* gate, long long ld_gate, float* x_out,
                  const float* shift, const float* scale, long long ld_mod, int rows_per_sample, const float* weight,
                  const float* bias, void* y, long long ldy, int ones_block, float* mean, float* rstd, int M, int D,
                  float eps, vaw_stream_t stream);
int vaw_ln_bwd(const void* dy, const float* x, const float* mean, const float* rstd, const float* scale,
               long long ld_mod, const float* weight, float* dx_io, int add_into, float* part, int rows_per_group,
               int groups, int chunks, int M, int D, vaw_stream_t stream);
/* vaw_ln_bwd fused with the vaw_gate_bwd of the branch that follows in the backward pass (saves re-reading the residual
 * gradient): afterwards dx_io = dx', dy_next = bf16(dx' * gate_next[group]), part_gate = (sum dx', sum dx' * y_next). */
int vaw_ln_bwd_gate(const void* dy, const float* x, const float* mean, const float* rstd, const float* scale,
                    long long ld_mod, const float* weight, float* dx_io, int add_into, float* part, const void* y_next,
                    const float* gate_next, long long ld_gate, void* dy_next, float* part_gate, int rows_per_group,
                    int groups, int chunks, int M, int D, vaw_stream_t stream);
int vaw_gate_bwd(const float* dx, const void* y, const float* gate, long long ld_gate, void* dy, float* part,
                 int rows_per_group, int groups, int chunks, int M, int D, vaw_stream_t stream);
int vaw_finish_group(const float* part, int which, int groups, int chunks, int D, float* out, long long ld_out,
                     int accumulate, vaw_stream_t stream);
int vaw_finish_all(const float* part, int which, int groups, int chunks, int D, const float* w, long long ld_w,
                   float* out, int accumulate, vaw_stream_t stream);
/* All partial-sum buffers of one DiT block's backward in one launch: pA / pC = vaw_gate_bwd partials of the MLP /
 * attention branch, pB / pD = vaw_ln_bwd partials of the MLP / attention LayerNorm.  Writes d mod[B, 6*D] (fp32 and a
 * bf16 copy, row stride ldm) and (accumulate ? adds to : overwrites) the fc2.bias, proj.bias and adaLN-bias gradients. */
int vaw_dit_block_finish(const float* pA, const float* pB, const float* pC, const float* pD, int B, int chunks, int D,
                         const float* mod, long long ldm, float* dmod, void* dmod_b, float* g_fc2_b, float* g_proj_b,
                         float* g_ada_b, int accumulate, vaw_stream_t stream);
int vaw_colsum_bf16(const void* a, long long lda, int M, int N, float* part, int rows_per_chunk, float* out,
                    int accumulate, vaw_stream_t stream);
int vaw_colsum_f32_small(const float* a, long long lda, int rows, int N, float* out, int accumulate,
                         vaw_stream_t stream);

/* ---- embedders, (un)patchify, casts (models/dit.py:41-110,243-256; timm PatchEmbed; models/uvit.py:21-52) ---------- */
int vaw_patchify_in(const float* x, void* patches, int B, int C, int H, int W, int P, vaw_stream_t stream);
/* REPA teacher input: preprocess_raw_image (tools/align_utils.py:19-40, the mocov3 / mae / dinov1 branch:
 * x / 255 then torchvision Normalize(mean, std)) fused into the patchify; raw pixels fp32 [B, C, H, W] -> bf16 patches
 * [B*T, C*P*P] in Conv2d-weight order.  vaw_vit_assemble builds the ViT residual stream of
 * encoders/mocov3_vit.py:99-106 (timm VisionTransformer._pos_embed): row 0 of every sample = cls + pos[0], rows 1..L =
 * the patch tokens [B*L, D] (bias and pos[1..] already added by the patch-embed GEMM's residual-table epilogue). */
int vaw_patchify_norm(const float* x, const float* mean, const float* stdv, void* patches, int B, int C, int H, int W,
                      int P, vaw_stream_t stream);
int vaw_vit_assemble(const float* tok, const float* cls, const float* pos0, float* x, int B, int L, int D,
                     vaw_stream_t stream);
int vaw_unpatchify(void* tokens, void* image, int dtype, int B, int C, int H, int W, int P, int to_image,
                   vaw_stream_t stream);
int vaw_timestep_embedding(const float* t, void* out_bf16, float* out_f32, int B, int dim, vaw_stream_t stream);
int vaw_cond_combine(const float* t_emb, const float* table, const long long* labels, float* c, void* c_silu, int B,
                     int D, vaw_stream_t stream);
int vaw_cond_bwd(const float* dc_silu, const float* c, float* dc, void* dc_bf16, int n, vaw_stream_t stream);
int vaw_embedding_grad(const float* dc, const long long* labels, float* dtable, int rows, int B, int D, int accumulate,
                       vaw_stream_t stream);
int vaw_cast_f32_bf16(const float* src, void* dst, long long n, vaw_stream_t stream);
int vaw_cast_f32_bf16_2d(const float* src, long long lds, void* dst, long long ldd, int rows, int cols,
                         vaw_stream_t stream);
int vaw_add_bf16_into_f32(const void* src, float* dst, long long n, vaw_stream_t stream);
int vaw_add_f32(const float* src, float* dst, long long n, vaw_stream_t stream); /* dst += src (n % 4 == 0) */

/* ---- K5: REPA alignment loss, type 'mse' (tools/gaussian_diffusion.py:1011-1013) ------------------------------------
 * loss = mean((zs - feat)^2) (scalar, deterministic two-stage reduction); dzs = gscale * 2 (zs - feat) / n (nullable).
 * dtype codes VAW_F32 / VAW_BF16; part: fp32 scratch of >= 1024 elements.                                            */
int vaw_align_mse(const void* zs, int zs_dtype, const void* feat, int feat_dtype, void* dzs, float gscale, long long n,
                  float* part, float* loss, vaw_stream_t stream);

/* Fold the [ceil(M/32) * ceil(N/32)] partials of VAW_EPI_ALIGN_MSE in fixed order: loss = sum / n. */
int vaw_align_mse_finish(const float* part, long long nparts, long long n, float* loss, vaw_stream_t stream);
/* dzs = (*g) * 2 (zs - feat) / n with the upstream gradient g a DEVICE scalar (backward of the fused loss). */
int vaw_align_mse_bwd(const void* zs, int zs_dtype, const void* feat, int feat_dtype, const float* g, void* dzs,
                      long long n, vaw_stream_t stream);
/* Row-wise alignment losses over [rows, D] (tools/gaussian_diffusion.py:1008-1019; F.cosine_similarity eps 1e-8,
 * F.normalize eps 1e-12): kind 0 'cosine': loss = -mean_r cos(feat_r, zs_r); kind 1 'mse_l2': loss = mean over all
 * elements of (zs_r/|zs_r| - feat_r/|feat_r|)^2.  dzs (nullable, dtype of zs) = gscale * d loss / d zs.  part: scratch
 * of >= rows floats.  One warp per row, fixed-order two-stage reduction. */
int vaw_align_rowwise(const void* zs, int zs_dtype, const void* feat, int feat_dtype, int kind, void* dzs, float gscale,
                      long long rows, int D, float* part, float* loss, vaw_stream_t stream);

/* ---- fused AdamW over the flat parameter buffer (main.py:354 optim.AdamW; trainer.py:12-18 EMA; §8f-1) ----------------
 * One pass: p, m, v updated in place from g * grad_scale; p_bf16 (nullable) refreshed; ema (nullable) updated.        */
int vaw_adamw_step(float* p, const float* g, float* m, float* v, void* p_bf16, float* ema, long long n, double lr,
                   double beta1, double beta2, double eps, double weight_decay, long long step, double grad_scale,
                   double ema_decay, const float* clip_coef, vaw_stream_t stream);
/* Same, for torch.cuda.amp.GradScaler-driven steps (tools/trainer.py:124-129 scaler.unscale_ / scaler.step): inv_scale
 * (device, nullable) multiplies the gradients, found_inf (device, nullable) != 0 skips the whole step - parameters,
 * moments, shadow and EMA untouched - so neither the unscale nor the inf check costs a pass or a host sync. */
int vaw_adamw_step_amp(float* p, const float* g, float* m, float* v, void* p_bf16, float* ema, long long n, double lr,
                       double beta1, double beta2, double eps, double weight_decay, long long step, double grad_scale,
                       double ema_decay, const float* clip_coef, const float* inv_scale, const float* found_inf,
                       vaw_stream_t stream);
/* The same update over a list of element ranges of the flat buffers in ONE launch: ranges = DEVICE array of n_ranges
 * (offset, count) pairs (multiples of 4), max_count = the largest count.  Used by the sharded data-parallel optimizer
 * (every rank updates its 1/W slice of each block's weights plus the replicated small tensors). */
int vaw_adamw_step_ranges(float* p, const float* g, float* m, float* v, void* p_bf16, float* ema, const long long* ranges,
                          int n_ranges, long long max_count, double lr, double beta1, double beta2, double eps,
                          double weight_decay, long long step, double grad_scale, double ema_decay,
                          const float* clip_coef, const float* inv_scale, const float* found_inf, vaw_stream_t stream);
/* Global L2 norm of the flat gradient buffer and the clip_grad_norm_ coefficient (tools/trainer.py:60-62), kept on the
 * device: out[0] = ||grad_scale * g||, out[1] = min(1, max_norm / (out[0] + 1e-6)); pass out + 1 as clip_coef above.
 * part: 1024 floats of scratch. */
int vaw_grad_clip_coef(const float* g, long long n, double grad_scale, double max_norm, float* part, float* out,
                       vaw_stream_t stream);

/* ---- DiT engine: forward / backward of the whole denoiser as one call each (models/dit.py:157-280) ----------------- */
typedef struct vaw_dit_cfg {
  int B, T, D, H, depth, hidden;
  int C_in, C_out, P, img_h, img_w;
  int table_rows, freq_dim;
  int learn_align, encoder_depth, proj_dim, z_dim;
} vaw_dit_cfg;

/* Offsets / sizes (elements) of every parameter tensor in the flat buffers; order documented in csrc/dit_engine.cu. */
int vaw_dit_param_layout(const vaw_dit_cfg* cfg, long long* offsets, long long* numels, int cap, int* n_out,
                         long long* total);
int vaw_dit_workspace_bytes(const vaw_dit_cfg* cfg, long long* bytes);
/* P fp32 params, Pb their bf16 shadow, ws the workspace; x_t fp32 [B,C,H,W], t fp32 [B] (scaled like
 * _scale_timesteps, tools/gaussian_diffusion.py:417-420), y int64 [B]; out bf16 [B,C_out,H,W]; zs bf16 [B*T,z_dim]. */
int vaw_dit_forward(const vaw_dit_cfg* cfg, const float* P, const void* Pb, void* ws, const float* x_t, const float* t,
                    const long long* y, void* out, void* zs, vaw_stream_t stream);
/* vaw_dit_forward gated by events: wait_events = depth + 1 cudaEvent_t (NULL entries allowed); [0] is waited on before
 * the stacked adaLN weights are read, [1 + i] before block i's weights.  The sharded data-parallel optimizer records them
 * as the all-gather of each block's bf16 weights completes, so the gathers of later blocks overlap earlier blocks. */
int vaw_dit_forward_ev(const vaw_dit_cfg* cfg, const float* P, const void* Pb, void* ws, const float* x_t, const float* t,
                       const long long* y, void* out, void* zs, void** wait_events, vaw_stream_t stream);
/* Forward-only entry for sampling / evaluation (the reference runs its samplers under torch.no_grad(), tools/sampler.py,
 * tools/cfg_edm.py:109): identical results, but nothing is kept for a backward pass - one set of operand buffers shared
 * by all blocks, a three-buffer residual ring, no saved pre-activations / branch outputs (a third of the forward's HBM
 * writes).  ws: vaw_dit_infer_workspace_bytes (DiT-XL/2, B = 128: 1.3 GB instead of the 40 GB training workspace). */
int vaw_dit_infer_workspace_bytes(const vaw_dit_cfg* cfg, long long* bytes);
int vaw_dit_forward_infer(const vaw_dit_cfg* cfg, const float* P, const void* Pb, void* ws, const float* x_t,
                          const float* t, const long long* y, void* out, void* zs, vaw_stream_t stream);
/* Same, with the alignment loss fused into the last projector GEMM (VAW_EPI_ALIGN_MSE): feat bf16 [B*T, z_dim] are the
 * teacher features, align_loss (device scalar) receives mean((zs - feat)^2).  learn_align configs only. */
int vaw_dit_forward_align(const vaw_dit_cfg* cfg, const float* P, const void* Pb, void* ws, const float* x_t,
                          const float* t, const long long* y, void* out, void* zs, const void* feat, float* align_loss,
                          vaw_stream_t stream);
/* G fp32 gradients (same layout as P); dout bf16 [B,C_out,H,W]; dzs bf16 or NULL; accumulate 0 = overwrite G.
 * events: NULL or depth+1 cudaEvent_t recorded as each block's gradients (last block first) become final.          */
int vaw_dit_backward(const vaw_dit_cfg* cfg, const float* P, const void* Pb, float* G, void* ws, const void* dout,
                     const void* dzs, const long long* y, int accumulate, void** events, vaw_stream_t stream);

/* ---- U-ViT engine (models/uvit.py:139-250) and its glue kernels ---------------------------------------------------- */
typedef struct vaw_uvit_cfg {
  int B, T, D, H, depth, hidden; /* T counts the extra tokens; depth = number of blocks (odd) */
  int C, P, img_h, img_w;
  int extras, table_rows; /* 1 = time token, 2 = label + time tokens */
  int conv;               /* final 3x3 convolution (uvit.py:192) */
} vaw_uvit_cfg;
int vaw_uvit_param_layout(const vaw_uvit_cfg* cfg, long long* offsets, long long* numels, int cap, int* n_out,
                          long long* total);
int vaw_uvit_workspace_bytes(const vaw_uvit_cfg* cfg, long long* bytes);
/* x_t fp32 [B,C,H,W], t fp32 [B], y int64 [B] (extras == 2) -> out fp32 [B,C,H,W] */
int vaw_uvit_forward(const vaw_uvit_cfg* cfg, const float* P, const void* Pb, void* ws, const float* x_t,
                     const float* t, const long long* y, float* out, vaw_stream_t stream);
int vaw_uvit_backward(const vaw_uvit_cfg* cfg, const float* P, const void* Pb, float* G, void* ws, const float* dout,
                      const long long* y, int accumulate, vaw_stream_t stream);
/* events: NULL or depth+1 cudaEvent_t recorded as each block's gradients (last block first) become final, then the
 * embedder / head tensors - the hook the data-parallel bucketed all-reduce (main.py:347 DDP) overlaps on. */
int vaw_uvit_backward_ev(const vaw_uvit_cfg* cfg, const float* P, const void* Pb, float* G, void* ws,
                         const float* dout, const long long* y, int accumulate, void** events, vaw_stream_t stream);
/* token assembly [label, time, patches] + pos_embed (uvit.py:221-231) and its backward pieces */
int vaw_uvit_assemble(const float* patch_tok, const float* t, const float* table, const long long* labels,
                      const float* pos, float* x0, int B, int T, int extras, int D, vaw_stream_t stream);
int vaw_uvit_pos_grad(const float* dx0, float* dpos, int B, int T, int D, int accumulate, vaw_stream_t stream);
int vaw_uvit_gather_patch_grad(const float* dx0, void* dtok, int B, int T, int extras, int D, vaw_stream_t stream);
int vaw_embedding_grad_strided(const float* dc, long long ld, const long long* labels, float* dtable, int rows, int B,
                               int D, int accumulate, vaw_stream_t stream);
/* skip_linear operand cat([x, skip]) (uvit.py:117-118) and the column split of its gradient */
int vaw_cat_cast(const float* x, const float* skip, void* cat, long long M, int D, vaw_stream_t stream);
int vaw_unpack_cols(const void* src, long long ld, int col_off, float* dst, long long M, int D, int accumulate,
                    vaw_stream_t stream);
int vaw_unpatchify_strided(void* tokens, int tok_dtype, void* image, int img_dtype, int B, int C, int H, int W, int P,
                           int to_image, int row0, int rows_per_sample, int zero_extras, vaw_stream_t stream);
/* final 3x3 convolution (uvit.py:192,248): forward / input gradient (transpose = 1) / weight + bias gradient */
int vaw_conv3x3(const float* in, const float* w, const float* bias, float* out, int B, int C, int H, int W,
                int transpose, vaw_stream_t stream);
int vaw_conv3x3_wgrad(const float* in, const float* dout, float* dw, float* dbias, int B, int C, int H, int W,
                      int accumulate, vaw_stream_t stream);

unsigned long long vaw_launch_count(void); /* kernels launched by this library in this process */
/* debug knob: device buffer of 2 x 64 uint64 globaltimer stamps written by one CTA of the tcgen05 attention backward
 * (role 0 = an elementwise warp, role 1 = the MMA warp); NULL switches tracing off (the default). */
int vaw_attn_set_trace(void* device_buf);

#ifdef __cplusplus
}
#endif
#endif /* VAW_B200_H_ */

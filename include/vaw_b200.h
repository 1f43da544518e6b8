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
int vaw_qsample_target(const float* x0, const float* noise, const long long* t, const float* tab_alpha,
                       const float* tab_sigma, const float* tab_c0, const float* tab_c1, float* x_t, float* target,
                       int mean_type, long long N, long long chw, vaw_stream_t stream);

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
#define VAW_EPI_F32 1        /* out(f32)  = acc + bias (+= when accumulate) */
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
  int k_splits;  /* > 1 (VAW_EPI_F32 only): split-K over k_splits work items per tile (deterministic reduce) */
  float* split_ws; /* fp32 scratch, k_splits * M * ldo elements */
} vaw_gemm_args;

int vaw_gemm_bf16(const vaw_gemm_args* args, vaw_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VAW_B200_H_ */

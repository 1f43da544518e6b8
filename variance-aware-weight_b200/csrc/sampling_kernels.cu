// sampling_kernels.cu — the elementwise side of the reference's ODE / SDE samplers, fused around the denoiser call
// (the shipped recipe samples with the EDM Heun solver, run.sh `--solver heun`).
//
//   vaw_edm_pre / vaw_edm_post   tools/cfg_edm.py: Net.forward (:51-79) + one Euler / Heun stage of ablation_sampler
//                                (:187-208).  The state is float64 like the reference's (`latents.to(torch.float64)`),
//                                the preconditioning float32; every product / sum / quotient is rounded separately in
//                                the reference's order (no FMA contraction), so for a given denoiser output the
//                                trajectory is bit-identical to the eager code.
//   vaw_flow_sde_step            tools/gaussian_diffusion.py FlowMatching: convert_model_output_to_vector (:1206-1228),
//                                convert_model_output_to_score (:1230-1257), compute_drift (:1371-1375) and the
//                                Euler-Maruyama / stochastic-Heun updates of sde_sample (:1381-1408), float32.
// All coefficients are per-step scalars evaluated on the host (same torch operations as the reference applies to its
// 0-dim tensors); the kernels are HBM-bound elementwise passes (launch-latency-bound at sampling batch sizes): the
// reference spends ~25 launches per denoiser evaluation on the same work.
#include "vaw_common.cuh"

namespace {

enum : int { PT_START_X = 2, PT_EPSILON = 3, PT_VELOCITY = 4, PT_VECTOR = 5, PT_SCORE = 6 };
enum : int { EDM_DENOISE = -1, EDM_EULER = 0, EDM_PREDICT = 1, EDM_CORRECT = 2 };

// x_hat = a x_cur + c noise (float64); x_in = c_in * float32(x_hat / s_hat)
__global__ void __launch_bounds__(256)
edm_pre_kernel(const double* __restrict__ x_cur, const double* __restrict__ noise, double a, double c, double s_hat,
               float c_in, double* __restrict__ x_hat, float* __restrict__ x_in, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double xh = __dmul_rn(a, x_cur[i]);
    if (noise) xh = __dadd_rn(xh, __dmul_rn(c, noise[i]));
    if (x_hat) x_hat[i] = xh;
    x_in[i] = __fmul_rn(c_in, (float)__ddiv_rn(xh, s_hat));
  }
}

template <typename OutT>
__global__ void __launch_bounds__(256)
edm_post_kernel(const OutT* __restrict__ out, long long out_stride, const double* __restrict__ x_src, double s_src,
                float c_skip, float c_out, int pred, double A, double Bc, int mode, double h,
                const double* __restrict__ x_hat, const double* __restrict__ d_prev, double c1, double c2,
                double* __restrict__ x_out, double* __restrict__ d_out, double s_next, float c_in_next,
                float* __restrict__ x_in_next, long long N, long long chw) {
  const long long total = N * chw;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long nn = i / chw;
    const float mo = (float)out[nn * out_stride + (i - nn * chw)];   // model_output[:, :img_channels].to(float32)
    const double xs = x_src[i];
    const float x32 = (float)__ddiv_rn(xs, s_src);                   // net(x / s(t), ...): x.to(torch.float32)
    // Net.forward :62-77
    const float den32 = pred == PT_START_X ? mo : __fadd_rn(__fmul_rn(c_skip, x32), __fmul_rn(c_out, mo));
    if (mode == EDM_DENOISE) {
      x_in_next[i] = den32;
      continue;
    }
    const double den = (double)den32;
    const double d = __dsub_rn(__dmul_rn(A, xs), __dmul_rn(Bc, den));             // :199 / :206
    double xo;
    if (mode == EDM_CORRECT) {
      const double mix = __dadd_rn(__dmul_rn(c1, d_prev[i]), __dmul_rn(c2, d));   // :207
      xo = __dadd_rn(x_hat[i], __dmul_rn(h, mix));
    } else {
      xo = __dadd_rn(x_hat[i], __dmul_rn(h, d));                                  // :200 (h = alpha h) / :204
      if (mode == EDM_PREDICT) {
        d_out[i] = d;
        x_in_next[i] = __fmul_rn(c_in_next, (float)__ddiv_rn(xo, s_next));
      }
    }
    x_out[i] = xo;
  }
}

struct FlowCoef {
  float a, s, da, ds, diff;   // alpha_t, sigma_t, d alpha_t, d sigma_t, 2 sigma_t d sigma_t
};

// vector - 0.5 * diffusion * score for one element (reference op order, float32)
__device__ __forceinline__ float flow_drift(int mt, const FlowCoef& k, float mo, float x) {
  float vec, score;
  if (mt == PT_VECTOR) {
    vec = mo;
    const float den = __fsub_rn(__fmul_rn(k.s, k.da), __fmul_rn(k.a, k.ds));
    const float noise = __fdiv_rn(__fsub_rn(__fmul_rn(k.da, x), __fmul_rn(k.a, mo)), den);
    score = __fdiv_rn(-noise, k.s);
  } else if (mt == PT_START_X) {
    const float r = __fsub_rn(x, __fmul_rn(k.a, mo));
    const float noise = __fdiv_rn(r, k.s);
    vec = __fadd_rn(__fmul_rn(k.da, mo), __fmul_rn(k.ds, noise));
    score = __fdiv_rn(-r, __fmul_rn(k.s, k.s));
  } else if (mt == PT_EPSILON) {
    const float xs = __fdiv_rn(__fsub_rn(x, __fmul_rn(k.s, mo)), k.a);
    vec = __fadd_rn(__fmul_rn(k.da, xs), __fmul_rn(k.ds, mo));
    score = __fdiv_rn(-mo, k.s);
  } else if (mt == PT_VELOCITY) {
    const float den = __fadd_rn(__fmul_rn(k.a, k.a), __fmul_rn(k.s, k.s));
    const float xs = __fdiv_rn(__fsub_rn(__fmul_rn(k.a, x), __fmul_rn(k.s, mo)), den);
    const float noise = __fdiv_rn(__fadd_rn(__fmul_rn(k.s, x), __fmul_rn(k.a, mo)), den);
    vec = __fadd_rn(__fmul_rn(k.da, xs), __fmul_rn(k.ds, noise));
    score = __fdiv_rn(-noise, k.s);
  } else {   // SCORE has no vector form in the reference (convert_model_output_to_vector raises)
    vec = 0.f;
    score = mo;
  }
  return __fsub_rn(vec, __fmul_rn(__fmul_rn(0.5f, k.diff), score));
}

template <typename OutT>
__global__ void __launch_bounds__(256)
flow_sde_step_kernel(const OutT* __restrict__ out, const float* __restrict__ x_eval, FlowCoef k, int mt, int mode,
                     const float* __restrict__ x_base, const float* __restrict__ drift_prev,
                     const float* __restrict__ noise, float step, float noise_scale, float sqrt_abs_step,
                     float* __restrict__ x_out, float* __restrict__ drift_out, long long n) {
  // noise_scale = th.sqrt(diffusion) of the step's CURRENT time (the Heun corrector evaluates its drift at the next
  // time but re-uses the predictor's noise term); NaN for a negative coefficient, exactly like the reference
  const float sd = noise_scale;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float d = flow_drift(mt, k, (float)out[i], x_eval[i]);
    if (drift_out) drift_out[i] = d;
    float upd = d;
    if (mode == 2) upd = __fmul_rn(0.5f, __fadd_rn(drift_prev[i], d));            // Heun corrector :1396
    float xo = __fadd_rn(x_base[i], __fmul_rn(upd, step));
    if (noise) xo = __fadd_rn(xo, __fmul_rn(__fmul_rn(sd, noise[i]), sqrt_abs_step));   // :1387
    x_out[i] = xo;
  }
}

inline unsigned grid_of(long long work) {
  long long b = (work + 255) / 256;
  const long long cap = (long long)vaw_num_sms() * 16;
  if (b > cap) b = cap;
  return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace

extern "C" int vaw_edm_pre(const double* x_cur, const double* noise, double a, double c, double s_hat, float c_in,
                           double* x_hat, float* x_in, long long n, cudaStream_t stream) {
  VAW_CHECK_ARG(x_cur && x_in && n >= 0, "vaw_edm_pre: bad arguments");
  if (n == 0) return VAW_OK;
  edm_pre_kernel<<<grid_of(n), 256, 0, stream>>>(x_cur, noise, a, c, s_hat, c_in, x_hat, x_in, n);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

extern "C" int vaw_edm_post(const void* out, int out_dtype, long long out_stride, const double* x_src, double s_src,
                            float c_skip, float c_out, int pred, double A, double Bc, int mode, double h,
                            const double* x_hat, const double* d_prev, double c1, double c2, double* x_out,
                            double* d_out, double s_next, float c_in_next, float* x_in_next, long long N, long long chw,
                            cudaStream_t stream) {
  VAW_CHECK_ARG(out && x_src && N >= 0 && chw > 0 && out_stride >= chw, "vaw_edm_post: bad arguments");
  VAW_CHECK_ARG(out_dtype == 0 || out_dtype == 1, "vaw_edm_post: out_dtype must be 0 (f32) or 1 (bf16)");
  VAW_CHECK_ARG(pred == PT_START_X || pred == PT_EPSILON || pred == PT_VELOCITY, "vaw_edm_post: bad pred type %d", pred);
  VAW_CHECK_ARG(mode >= EDM_DENOISE && mode <= EDM_CORRECT, "vaw_edm_post: bad mode %d", mode);
  VAW_CHECK_ARG(mode == EDM_DENOISE ? x_in_next != nullptr : (x_hat && x_out), "vaw_edm_post: missing output");
  VAW_CHECK_ARG(mode != EDM_PREDICT || (d_out && x_in_next), "vaw_edm_post: the predictor writes d_out and x_in_next");
  VAW_CHECK_ARG(mode != EDM_CORRECT || d_prev, "vaw_edm_post: the corrector needs d_prev");
  if (N == 0) return VAW_OK;
  if (out_dtype == 0)
    edm_post_kernel<float><<<grid_of(N * chw), 256, 0, stream>>>((const float*)out, out_stride, x_src, s_src, c_skip,
                                                                 c_out, pred, A, Bc, mode, h, x_hat, d_prev, c1, c2, x_out,
                                                                 d_out, s_next, c_in_next, x_in_next, N, chw);
  else
    edm_post_kernel<bf16><<<grid_of(N * chw), 256, 0, stream>>>((const bf16*)out, out_stride, x_src, s_src, c_skip, c_out,
                                                                pred, A, Bc, mode, h, x_hat, d_prev, c1, c2, x_out, d_out,
                                                                s_next, c_in_next, x_in_next, N, chw);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

extern "C" int vaw_flow_sde_step(const void* out, int out_dtype, const float* x_eval, const float* coef, int mean_type,
                                 int mode, const float* x_base, const float* drift_prev, const float* noise, float step,
                                 float noise_scale, float sqrt_abs_step, float* x_out, float* drift_out, long long n,
                                 cudaStream_t stream) {
  VAW_CHECK_ARG(out && x_eval && coef && x_base && x_out && n >= 0, "vaw_flow_sde_step: bad arguments");
  VAW_CHECK_ARG(out_dtype == 0 || out_dtype == 1, "vaw_flow_sde_step: out_dtype must be 0 (f32) or 1 (bf16)");
  VAW_CHECK_ARG(mean_type >= PT_START_X && mean_type <= PT_VECTOR, "vaw_flow_sde_step: mean type %d has no vector form",
                mean_type);
  VAW_CHECK_ARG(mode >= 0 && mode <= 2 && (mode != 2 || drift_prev), "vaw_flow_sde_step: bad mode");
  if (n == 0) return VAW_OK;
  const FlowCoef k{coef[0], coef[1], coef[2], coef[3], coef[4]};   // HOST array: alpha, sigma, d_alpha, d_sigma, diffusion
  if (out_dtype == 0)
    flow_sde_step_kernel<float><<<grid_of(n), 256, 0, stream>>>((const float*)out, x_eval, k, mean_type, mode, x_base,
                                                                drift_prev, noise, step, noise_scale, sqrt_abs_step, x_out, drift_out, n);
  else
    flow_sde_step_kernel<bf16><<<grid_of(n), 256, 0, stream>>>((const bf16*)out, x_eval, k, mean_type, mode, x_base,
                                                               drift_prev, noise, step, noise_scale, sqrt_abs_step, x_out,
                                                               drift_out, n);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

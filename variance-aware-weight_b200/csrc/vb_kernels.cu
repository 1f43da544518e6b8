// vb_kernels.cu — the variational-bound term of the training objective, forward value and gradient in one pass.
//
// Replaces, inside GaussianDiffusion.training_losses (reference tools/gaussian_diffusion.py:862-875 for LossType.KL /
// RESCALED_KL and :886-906 for the learned-variance term added to the MSE objective):
//   q_posterior_mean_variance (:254-276) + p_mean_variance with clip_denoised=False (:278-384) + normal_kl
//   (tools/losses.py:12-40) + discretized_gaussian_log_likelihood (tools/losses.py:43-77) + mean_flat / ln 2 +
//   th.where(t == 0, decoder_nll, kl) (:794-807), and the autograd backward of that chain (~60 elementwise launches
//   and 7 table uploads per step in the reference).
// One CTA per sample: reads the model output (mean channels and, for the LEARNED* variance types, the variance channels
// behind them), x_0 and x_t once, accumulates the per-sample bound with a fixed reduction tree and writes
// d vb_n / d out in the same pass (hand-derived; tests check it against autograd over the oracle).  The arithmetic is
// fp32 in the reference's operation order: the decoder NLL at t = 0 lives where fp32 tanh saturates (cdf values of
// exactly 1, the 1e-12 clamps), so the order is part of the result.  For a bf16 model output the (v + 1) / 2 and
// 1 - frac intermediates are rounded to bf16 like the reference's ops on a bf16 tensor (:321-323).
#include <type_traits>

#include "vaw_common.cuh"

namespace {

enum : int { MT_PREVIOUS_X = 1, MT_START_X = 2, MT_EPSILON = 3, MT_VELOCITY = 4 };
enum : int { VT_LEARNED = 1, VT_FIXED_SMALL = 2, VT_FIXED_LARGE = 3, VT_LEARNED_RANGE = 4 };
// rows of the [VAW_RT_ROWS][T] table (include/vaw_b200.h)
enum : int { RT_SQRT_RECIP_AC = 0, RT_SQRT_RECIPM1_AC, RT_SQRT_AC, RT_SQRT_1MAC, RT_INV_COEF1, RT_COEF2_OVER_COEF1,
             RT_COEF1, RT_COEF2, RT_LOGVAR, RT_MAX_LOG, RT_VARIANCE, RT_AC, RT_AC_PREV, RT_AC_NEXT, RT_TRUE_LOGVAR,
             RT_ROWS };

__device__ __forceinline__ float bf16r(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// approx_standard_normal_cdf (losses.py:43-48) and its derivative
__device__ __forceinline__ float cdf_approx(float x, float* dcdf) {
  const float k = 0.7978845608028654f;  // sqrt(2 / pi)
  const float x3 = __fmul_rn(__fmul_rn(x, x), x);   // th.pow(x, 3); every op rounded separately like the eager chain
  const float u = __fmul_rn(k, __fadd_rn(x, __fmul_rn(0.044715f, x3)));
  const float th = tanhf(u);
  *dcdf = 0.5f * (1.f - th * th) * k * (1.f + 3.f * 0.044715f * x * x);
  return __fmul_rn(0.5f, __fadd_rn(1.f, th));
}

template <typename OutT, bool BF>
__global__ void __launch_bounds__(256)
vb_terms_kernel(const OutT* __restrict__ out, long long out_stride, const float* __restrict__ x0,
                const float* __restrict__ x_t, const long long* __restrict__ t, const float* __restrict__ tab, int T,
                float* __restrict__ vb, OutT* __restrict__ grad, long long grad_stride,
                const float* __restrict__ gscale_n, float gscale, int mean_type, int var_type, int detach_mean,
                float out_scale, long long chw) {
  const long long n = blockIdx.x;
  long long tt = t[n];
  const bool first = tt == 0;
  tt = tt < 0 ? 0 : (tt >= T ? T - 1 : tt);   // memory safety only (the reference's gather would device-assert)
  auto row = [&](int r) { return __ldg(tab + (long long)r * T + tt); };
  const float c1 = row(RT_COEF1), c2 = row(RT_COEF2);
  const float tlv = row(RT_TRUE_LOGVAR);
  const float maxlog = row(RT_MAX_LOG), fixed_lv = row(RT_LOGVAR);
  float p = 1.f, q = 0.f;   // pred_xstart = p * x_t - q * o  (EPSILON, VELOCITY); START_X: o
  if (mean_type == MT_EPSILON) { p = row(RT_SQRT_RECIP_AC); q = row(RT_SQRT_RECIPM1_AC); }
  else if (mean_type == MT_VELOCITY) { p = row(RT_SQRT_AC); q = row(RT_SQRT_1MAC); }
  // d mean / d o
  const float dmean_do = mean_type == MT_PREVIOUS_X ? 1.f : (mean_type == MT_START_X ? c1 : -c1 * q);
  const bool learned = var_type == VT_LEARNED || var_type == VT_LEARNED_RANGE;
  const float dlv_dv = var_type == VT_LEARNED ? 1.f : 0.5f * (maxlog - tlv);
  const float inv_n = 1.0f / (float)chw;
  const float kLn2 = 0.6931471805599453f;
  const float G = (gscale_n ? __ldg(gscale_n + n) : 1.f) * gscale * out_scale * inv_n / kLn2;
  const OutT* on = out + n * out_stride;
  OutT* gn = grad ? grad + n * grad_stride : nullptr;
  const long long base = n * chw;
  float acc = 0.f;
  for (long long i = threadIdx.x; i < chw; i += blockDim.x) {
    const float o = (float)on[i];
    const float xs0 = x0[base + i], xt = x_t[base + i];
    // ---- model log-variance (:315-324) ----
    float lv;
    if (var_type == VT_LEARNED) {
      lv = (float)on[chw + i];
    } else if (var_type == VT_LEARNED_RANGE) {
      const float v = (float)on[chw + i];
      float frac = __fadd_rn(v, 1.f);
      if (BF) frac = bf16r(frac);
      frac = __fdiv_rn(frac, 2.f);
      if (BF) frac = bf16r(frac);
      float om = __fsub_rn(1.f, frac);
      if (BF) om = bf16r(om);
      lv = __fadd_rn(__fmul_rn(frac, maxlog), __fmul_rn(om, tlv));
    } else {
      lv = fixed_lv;
    }
    // ---- model mean (:354-384 with clip_denoised=False) and the true posterior mean (:254-276) ----
    float mean;
    if (mean_type == MT_PREVIOUS_X) {
      mean = o;
    } else {
      const float xs = mean_type == MT_START_X ? o : __fsub_rn(__fmul_rn(p, xt), __fmul_rn(q, o));
      mean = __fadd_rn(__fmul_rn(c1, xs), __fmul_rn(c2, xt));
    }
    const float tm = __fadd_rn(__fmul_rn(c1, xs0), __fmul_rn(c2, xt));
    float elem, d_lv, d_mean;
    if (!first) {
      // normal_kl (losses.py:33-39): 0.5 * (-1 + lv2 - lv1 + exp(lv1 - lv2) + (m1 - m2)^2 * exp(-lv2))
      const float e1 = expf(__fsub_rn(tlv, lv));
      const float e2 = expf(-lv);
      const float dm = __fsub_rn(tm, mean);
      const float sq = __fmul_rn(dm, dm);
      float s = __fadd_rn(-1.f, lv);
      s = __fsub_rn(s, tlv);
      s = __fadd_rn(s, e1);
      s = __fadd_rn(s, __fmul_rn(sq, e2));
      elem = __fmul_rn(0.5f, s);
      d_lv = 0.5f * (1.f - e1 - sq * e2);
      d_mean = -dm * e2;
    } else {
      // -discretized_gaussian_log_likelihood(x_start, means, log_scales = 0.5 * lv) (losses.py:51-77)
      const float cx = __fsub_rn(xs0, mean);
      const float inv = expf(-__fmul_rn(0.5f, lv));
      const float pin = __fmul_rn(inv, __fadd_rn(cx, 1.0f / 255.0f));
      const float min_ = __fmul_rn(inv, __fsub_rn(cx, 1.0f / 255.0f));
      float dcp, dcm;
      const float cp = cdf_approx(pin, &dcp);
      const float cm = cdf_approx(min_, &dcm);
      float lp, dl_dpin = 0.f, dl_dmin = 0.f;   // log prob and its derivatives w.r.t. plus_in / min_in
      if (xs0 < -0.999f) {
        lp = logf(fmaxf(cp, 1e-12f));
        if (cp >= 1e-12f) dl_dpin = dcp / cp;
      } else if (xs0 > 0.999f) {
        const float om = __fsub_rn(1.f, cm);
        lp = logf(fmaxf(om, 1e-12f));
        if (om >= 1e-12f) dl_dmin = -dcm / om;
      } else {
        const float delta = __fsub_rn(cp, cm);
        lp = logf(fmaxf(delta, 1e-12f));
        if (delta >= 1e-12f) { dl_dpin = dcp / delta; dl_dmin = -dcm / delta; }
      }
      elem = -lp;
      // d plus_in / d lv = -0.5 plus_in, d plus_in / d mean = -inv (same for min_in)
      d_lv = 0.5f * (dl_dpin * pin + dl_dmin * min_);
      d_mean = inv * (dl_dpin + dl_dmin);
    }
    acc += elem;
    if (gn) {
      if (learned) gn[chw + i] = (OutT)(G * d_lv * dlv_dv);
      if (!detach_mean) gn[i] = (OutT)(G * d_mean * dmean_do);
    }
  }
  __shared__ float red[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i];
    vb[n] = out_scale * ((s * inv_n) / kLn2);
  }
}

}  // namespace

extern "C" int vaw_vb_terms(const void* out, int out_dtype, long long out_stride, const float* x0, const float* x_t,
                            const long long* t, const float* tab, int T, float* vb, void* grad, long long grad_stride,
                            const float* gscale_n, float gscale, int mean_type, int var_type, int detach_mean,
                            float out_scale, long long N, long long chw, cudaStream_t stream) {
  VAW_CHECK_ARG(out && x0 && x_t && t && tab && vb && T > 0, "vaw_vb_terms: null pointer");
  VAW_CHECK_ARG(out_dtype == 0 || out_dtype == 1, "vaw_vb_terms: out_dtype must be 0 (f32) or 1 (bf16)");
  VAW_CHECK_ARG(N >= 0 && chw > 0, "vaw_vb_terms: bad shape N=%lld chw=%lld", N, chw);
  VAW_CHECK_ARG(mean_type >= MT_PREVIOUS_X && mean_type <= MT_VELOCITY, "vaw_vb_terms: mean_type %d has no posterior", mean_type);
  VAW_CHECK_ARG(var_type >= VT_LEARNED && var_type <= VT_LEARNED_RANGE, "vaw_vb_terms: bad var_type %d", var_type);
  const bool learned = var_type == VT_LEARNED || var_type == VT_LEARNED_RANGE;
  VAW_CHECK_ARG(out_stride >= (learned ? 2 : 1) * chw && (!grad || grad_stride >= (learned ? 2 : 1) * chw),
                "vaw_vb_terms: strides too small for the channel layout");
  if (N == 0) return VAW_OK;
  if (out_dtype == 0)
    vb_terms_kernel<float, false><<<(unsigned)N, 256, 0, stream>>>(
        (const float*)out, out_stride, x0, x_t, t, tab, T, vb, (float*)grad, grad_stride, gscale_n, gscale, mean_type,
        var_type, detach_mean, out_scale, chw);
  else
    vb_terms_kernel<bf16, true><<<(unsigned)N, 256, 0, stream>>>(
        (const bf16*)out, out_stride, x0, x_t, t, tab, T, vb, (bf16*)grad, grad_stride, gscale_n, gscale, mean_type,
        var_type, detach_mean, out_scale, chw);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

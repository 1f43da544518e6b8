// diffusion_kernels.cu — K1 (fused q_sample + target) and K2 (fused weighted-MSE forward + backward).
//
// Replaces, on the hot path of GaussianDiffusion.training_losses (reference
// tools/gaussian_diffusion.py:834-930):
//   K1: q_sample (:234-252) + compute_target (:818-832) + the four _extract_into_tensor gathers (:1059-1072)
//   K2: (target - out)**2 -> mean_flat (tools/nn.py:86-90) -> w * raw_mse (:911-913) and its autograd backward
//
// Rounding contract (bit-exact with the reference's fp32 eager path):
//   x_t    = fl( fl(a*x0) + fl(s*eps) )      two multiplies and one add, never contracted into an FMA
//   v-tgt  = fl( fl(a*eps) - fl(s*x0) )
// a = float(sqrt_alphas_cumprod[t]), s = float(sqrt_one_minus_alphas_cumprod[t]) (f64 table rounded once to f32).
#include <type_traits>

#include "vaw_common.cuh"

namespace {

// ModelMeanType codes follow the reference enum (enum.auto() starts at 1), gaussian_diffusion.py:21-32
enum : int { MT_PREVIOUS_X = 1, MT_START_X = 2, MT_EPSILON = 3, MT_VELOCITY = 4, MT_VECTOR = 5, MT_SCORE = 6 };

struct Coef {
  float a, s;    // alpha_t, sigma_t
  float c0, c1;  // posterior_mean_coef1/2 (PREVIOUS_X) or d_alpha/d_sigma (VECTOR)
};

__device__ __forceinline__ Coef load_coef(const long long* __restrict__ t, const float* __restrict__ tab_a,
                                          const float* __restrict__ tab_s, const float* __restrict__ tab_c0,
                                          const float* __restrict__ tab_c1, long long n) {
  // t != null: tables are [T] and indexed by the integer timestep (diffusion mode)
  // t == null: the "tables" are per-sample arrays [N] (flow-matching mode, continuous time)
  long long i = t ? t[n] : n;
  Coef c;
  c.a = __ldg(tab_a + i);
  c.s = __ldg(tab_s + i);
  c.c0 = tab_c0 ? __ldg(tab_c0 + i) : 0.f;
  c.c1 = tab_c1 ? __ldg(tab_c1 + i) : 0.f;
  return c;
}

__device__ __forceinline__ float mix2(float p, float u, float q, float v) {
  return __fadd_rn(__fmul_rn(p, u), __fmul_rn(q, v));
}

__device__ __forceinline__ float target_of(int mean_type, const Coef& c, float x0, float eps, float xt) {
  switch (mean_type) {
    case MT_START_X: return x0;
    case MT_EPSILON: return eps;
    case MT_VELOCITY: return __fsub_rn(__fmul_rn(c.a, eps), __fmul_rn(c.s, x0));
    case MT_PREVIOUS_X: return mix2(c.c0, x0, c.c1, xt);
    case MT_VECTOR: return mix2(c.c0, x0, c.c1, eps);
    case MT_SCORE: return __fdiv_rn(-eps, c.s);
    default: return eps;
  }
}

// ------------------------------------------------------------------------------------------------
// K1: x_t (and optionally the regression target) from x0, eps, t.  12 B/element (eps/x0 target need
// no extra write), 16 B/element when the target is materialised.
// ------------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(256)
qsample_target_kernel(const float* __restrict__ x0, const float* __restrict__ eps, const long long* __restrict__ t,
                      const float* __restrict__ tab_a, const float* __restrict__ tab_s,
                      const float* __restrict__ tab_c0, const float* __restrict__ tab_c1, float* __restrict__ x_t,
                      float* __restrict__ target, int mean_type, long long N, long long chw) {
  if (VEC) {
    const long long chw4 = chw >> 2;
    const long long total4 = N * chw4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += stride) {
      const long long n = i / chw4;
      const Coef c = load_coef(t, tab_a, tab_s, tab_c0, tab_c1, n);
      const float4 x = ldg_stream_f4(reinterpret_cast<const float4*>(x0) + i);
      const float4 e = ldg_stream_f4(reinterpret_cast<const float4*>(eps) + i);
      float4 o;
      o.x = mix2(c.a, x.x, c.s, e.x);
      o.y = mix2(c.a, x.y, c.s, e.y);
      o.z = mix2(c.a, x.z, c.s, e.z);
      o.w = mix2(c.a, x.w, c.s, e.w);
      stg_stream_f4(reinterpret_cast<float4*>(x_t) + i, o);
      if (target) {
        float4 g;
        g.x = target_of(mean_type, c, x.x, e.x, o.x);
        g.y = target_of(mean_type, c, x.y, e.y, o.y);
        g.z = target_of(mean_type, c, x.z, e.z, o.z);
        g.w = target_of(mean_type, c, x.w, e.w, o.w);
        stg_stream_f4(reinterpret_cast<float4*>(target) + i, g);
      }
    }
  } else {
    const long long total = N * chw;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
      const long long n = i / chw;
      const Coef c = load_coef(t, tab_a, tab_s, tab_c0, tab_c1, n);
      const float x = x0[i], e = eps[i];
      const float o = mix2(c.a, x, c.s, e);
      x_t[i] = o;
      if (target) target[i] = target_of(mean_type, c, x, e, o);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K2: one CTA per sample.  Reads the model output (fp32 or bf16), rebuilds the target on the fly from
// x0/eps (no target tensor in HBM), accumulates sum((target-out)^2) with a fixed reduction tree
// (thread-serial -> warp shuffle -> 8-warp shared reduce), and writes
//   mse[n]  = w_n * sum / chw          (terms["mse"], reference :913)
//   grad[n,i] = g_n * w_n * 2 (out - target) / chw     (d loss_n / d out, times the upstream per-sample factor)
// in the same pass.  12 B/element fp32, 8 B/element bf16 (+ x0/eps re-read when the target needs both).
// ------------------------------------------------------------------------------------------------
template <typename OutT>
struct Vec4;
template <>
struct Vec4<float> {
  static __device__ __forceinline__ float4 load(const float* p, long long i4) {
    return ldg_stream_f4(reinterpret_cast<const float4*>(p) + i4);
  }
  static __device__ __forceinline__ void store(float* p, long long i4, float4 v) {
    stg_stream_f4(reinterpret_cast<float4*>(p) + i4, v);
  }
};
template <>
struct Vec4<bf16> {
  static __device__ __forceinline__ float4 load(const bf16* p, long long i4) {
    uint2 u = ldg_stream_u2(reinterpret_cast<const uint2*>(p) + i4);
    float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
    return make_float4(a.x, a.y, b.x, b.y);
  }
  static __device__ __forceinline__ void store(bf16* p, long long i4, float4 v) {
    uint2 u;
    u.x = pack_bf16(v.x, v.y);
    u.y = pack_bf16(v.z, v.w);
    stg_stream_u2(reinterpret_cast<uint2*>(p) + i4, u);
  }
};

template <typename OutT>
__global__ void __launch_bounds__(256)
wmse_fwd_bwd_kernel(const OutT* __restrict__ out, const float* __restrict__ x0, const float* __restrict__ eps,
                    const long long* __restrict__ t, const float* __restrict__ tab_a, const float* __restrict__ tab_s,
                    const float* __restrict__ tab_c0, const float* __restrict__ tab_c1,
                    const float* __restrict__ w_tab, float* __restrict__ mse, float* __restrict__ raw_mse,
                    OutT* __restrict__ grad, const float* __restrict__ gscale_n, float gscale, int mean_type,
                    long long chw, int vec, long long out_stride, long long grad_stride) {
  const long long n = blockIdx.x;
  out += n * (out_stride - chw);     // rows of the model output / gradient may be further apart than chw (the mean
  if (grad) grad += n * (grad_stride - chw);   // channels of a [N, 2C, H, W] learned-variance output)
  const Coef c = load_coef(t, tab_a, tab_s, tab_c0, tab_c1, n);
  const float w = w_tab ? __ldg(w_tab + (t ? t[n] : n)) : 1.f;
  const float inv = 1.0f / (float)chw;
  const float g = (gscale_n ? __ldg(gscale_n + n) : 1.f) * gscale * w * 2.f * inv;
  const bool need_x0 = (mean_type != MT_EPSILON && mean_type != MT_SCORE);
  const bool need_eps = (mean_type != MT_START_X);
  const long long base = n * chw;
  float acc = 0.f;
  if (vec) {
    const long long chw4 = chw >> 2, base4 = base >> 2;
    for (long long i = threadIdx.x; i < chw4; i += blockDim.x) {
      const float4 o = Vec4<OutT>::load(out, base4 + i);
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f), e = x;
      if (need_x0) x = ldg_stream_f4(reinterpret_cast<const float4*>(x0) + base4 + i);
      if (need_eps) e = ldg_stream_f4(reinterpret_cast<const float4*>(eps) + base4 + i);
      float4 tg;
      tg.x = target_of(mean_type, c, x.x, e.x, mix2(c.a, x.x, c.s, e.x));
      tg.y = target_of(mean_type, c, x.y, e.y, mix2(c.a, x.y, c.s, e.y));
      tg.z = target_of(mean_type, c, x.z, e.z, mix2(c.a, x.z, c.s, e.z));
      tg.w = target_of(mean_type, c, x.w, e.w, mix2(c.a, x.w, c.s, e.w));
      const float dx = tg.x - o.x, dy = tg.y - o.y, dz = tg.z - o.z, dw = tg.w - o.w;
      acc += dx * dx;
      acc += dy * dy;
      acc += dz * dz;
      acc += dw * dw;
      if (grad) Vec4<OutT>::store(grad, base4 + i, make_float4(-g * dx, -g * dy, -g * dz, -g * dw));
    }
  } else {
    for (long long i = threadIdx.x; i < chw; i += blockDim.x) {
      const float o = (float)out[base + i];
      const float x = need_x0 ? x0[base + i] : 0.f, e = need_eps ? eps[base + i] : 0.f;
      const float tg = target_of(mean_type, c, x, e, mix2(c.a, x, c.s, e));
      const float d = tg - o;
      acc += d * d;
      if (grad) grad[base + i] = (OutT)(-g * d);
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
    const float m = s * inv;
    if (raw_mse) raw_mse[n] = m;
    mse[n] = w * m;
  }
}

// per-sample row scale: y[n, :] = x[n, :] * s[n] (backward of K2 when the upstream grad arrives late)
template <typename T>
__global__ void __launch_bounds__(256)
scale_rows_kernel(const T* __restrict__ x, const float* __restrict__ s, T* __restrict__ y, long long N,
                  long long chw) {
  const long long total = N * chw;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    y[i] = (T)((float)x[i] * s[i / chw]);
  }
}

// sample_from_latent (tools/trainer.py:21-25): latent [N, 2C, H, W] = (mean | std) along channels;
// out = (mean + std * eps) * scale with every product / sum rounded separately (torch's elementwise sequence).
__global__ void __launch_bounds__(256)
sample_from_latent_kernel(const float* __restrict__ latent, const float* __restrict__ eps, float* __restrict__ out,
                          long long N, long long chw, float scale) {
  const long long total = N * chw;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long n = i / chw, r = i - n * chw;
    const float mean = latent[n * 2 * chw + r], sd = latent[n * 2 * chw + chw + r];
    out[i] = __fmul_rn(__fadd_rn(mean, __fmul_rn(sd, eps[i])), scale);
  }
}

// ------------------------------------------------------------------------------------------------
// K8: one fused reverse-process step.  Replaces the elementwise tail of p_mean_variance
// (gaussian_diffusion.py:278-384) + p_sample (:455-506) / ddim_sample (:603-651) / ddim_reverse_sample (:653-689):
// the ~25 elementwise launches and 8 table uploads of the reference become one pass that reads the model
// output, x_t (and the noise) once and writes the next sample (and, on request, pred_xstart / mean / variance).
// Every product, sum, quotient and square root is rounded separately in the reference's order, so the fp32
// results are bit-identical to the eager path; only exp() (device libm vs host libm) may differ by an ulp.
// For a bf16 model output the variance branch keeps the reference's bf16 intermediates.
// ------------------------------------------------------------------------------------------------
enum : int { VT_LEARNED = 1, VT_FIXED_SMALL = 2, VT_FIXED_LARGE = 3, VT_LEARNED_RANGE = 4 };
enum : int { RS_DDPM = 0, RS_DDIM = 1, RS_DDIM_REVERSE = 2, RS_MOMENTS = 3 };
enum : int { RT_SQRT_RECIP_AC = 0, RT_SQRT_RECIPM1_AC, RT_SQRT_AC, RT_SQRT_1MAC, RT_INV_COEF1, RT_COEF2_OVER_COEF1,
             RT_COEF1, RT_COEF2, RT_LOGVAR, RT_MAX_LOG, RT_VARIANCE, RT_AC, RT_AC_PREV, RT_AC_NEXT, RT_ROWS };

struct StepCoef {
  float p, q;          // pred_xstart = p * (x_t or xprev) - q * (model output or x_t)
  float c1, c2;        // posterior mean
  float lv, maxlog, var;
  float recip, recipm1;
  float k_xs, k_eps, sig;  // DDIM: sample = k_xs * xs + k_eps * eps + sig * noise
  float mask, fixed_scale;
};

__device__ __forceinline__ float bf16r(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

__device__ __forceinline__ StepCoef load_step_coef(const float* __restrict__ tab, int T, long long tt, int mean_type,
                                                   int mode, float eta) {
  StepCoef c;
  tt = tt < 0 ? 0 : (tt >= T ? T - 1 : tt);  // memory safety only: the reference device-asserts on such an index
  auto row = [&](int r) { return __ldg(tab + (long long)r * T + tt); };
  c.recip = row(RT_SQRT_RECIP_AC);
  c.recipm1 = row(RT_SQRT_RECIPM1_AC);
  if (mean_type == MT_EPSILON) { c.p = c.recip; c.q = c.recipm1; }
  else if (mean_type == MT_VELOCITY) { c.p = row(RT_SQRT_AC); c.q = row(RT_SQRT_1MAC); }
  else if (mean_type == MT_PREVIOUS_X) { c.p = row(RT_INV_COEF1); c.q = row(RT_COEF2_OVER_COEF1); }
  else { c.p = 1.f; c.q = 0.f; }
  c.c1 = row(RT_COEF1);
  c.c2 = row(RT_COEF2);
  c.lv = row(RT_LOGVAR);
  c.maxlog = row(RT_MAX_LOG);
  c.var = row(RT_VARIANCE);
  c.mask = tt != 0 ? 1.f : 0.f;
  c.fixed_scale = mode == RS_DDPM ? expf(__fmul_rn(0.5f, c.lv)) : 0.f;  // fixed variance types: exp(0.5 logvar[t])
  c.k_xs = c.k_eps = c.sig = 0.f;
  if (mode == RS_DDIM) {
    const float ab = row(RT_AC), abp = row(RT_AC_PREV);
    const float om_abp = __fsub_rn(1.f, abp);
    const float s1 = __fsqrt_rn(__fdiv_rn(om_abp, __fsub_rn(1.f, ab)));
    const float s2 = __fsqrt_rn(__fsub_rn(1.f, __fdiv_rn(ab, abp)));
    const float sigma = __fmul_rn(__fmul_rn(eta, s1), s2);
    c.k_xs = __fsqrt_rn(abp);
    c.k_eps = __fsqrt_rn(__fsub_rn(om_abp, __fmul_rn(sigma, sigma)));
    c.sig = __fmul_rn(c.mask, sigma);
  } else if (mode == RS_DDIM_REVERSE) {
    const float abn = row(RT_AC_NEXT);
    c.k_xs = __fsqrt_rn(abn);
    c.k_eps = __fsqrt_rn(__fsub_rn(1.f, abn));
  }
  return c;
}

struct StepOut { float sample, xs, mean, lv, var; };

// BF: the model output arrived as bf16, so the reference's variance arithmetic on it rounds to bf16 per op
template <bool BF>
__device__ __forceinline__ StepOut step_elem(const StepCoef& c, float o, float vv, float x, float z, int mean_type,
                                             int var_type, int mode, bool clip) {
  StepOut r;
  float xs;
  if (mean_type == MT_START_X) xs = o;
  else if (mean_type == MT_PREVIOUS_X) xs = __fsub_rn(__fmul_rn(c.p, o), __fmul_rn(c.q, x));
  else xs = __fsub_rn(__fmul_rn(c.p, x), __fmul_rn(c.q, o));
  if (clip) xs = xs < -1.f ? -1.f : (xs > 1.f ? 1.f : xs);  // NaN stays NaN, as torch.clamp
  r.xs = xs;
  r.mean = mean_type == MT_PREVIOUS_X ? o : __fadd_rn(__fmul_rn(c.c1, xs), __fmul_rn(c.c2, x));
  float half_lv_exp;  // exp(0.5 * log_variance) in the dtype the reference computes it in
  if (var_type == VT_LEARNED) {
    r.lv = vv;
    r.var = BF ? bf16r(expf(vv)) : expf(vv);
    const float h = __fmul_rn(0.5f, vv);
    half_lv_exp = BF ? bf16r(expf(bf16r(h))) : expf(h);
  } else if (var_type == VT_LEARNED_RANGE) {
    float frac = __fadd_rn(vv, 1.f);
    if (BF) frac = bf16r(frac);
    frac = __fdiv_rn(frac, 2.f);
    if (BF) frac = bf16r(frac);
    float om = __fsub_rn(1.f, frac);
    if (BF) om = bf16r(om);
    r.lv = __fadd_rn(__fmul_rn(frac, c.maxlog), __fmul_rn(om, c.lv));
    r.var = expf(r.lv);
    half_lv_exp = expf(__fmul_rn(0.5f, r.lv));
  } else {
    r.lv = c.lv;
    r.var = c.var;
    half_lv_exp = c.fixed_scale;
  }
  if (mode == RS_DDPM) {
    r.sample = __fadd_rn(r.mean, __fmul_rn(__fmul_rn(c.mask, half_lv_exp), z));
  } else if (mode == RS_DDIM || mode == RS_DDIM_REVERSE) {
    const float eps = __fdiv_rn(__fsub_rn(__fmul_rn(c.recip, x), xs), c.recipm1);
    const float mp = __fadd_rn(__fmul_rn(xs, c.k_xs), __fmul_rn(c.k_eps, eps));
    r.sample = mode == RS_DDIM ? __fadd_rn(mp, __fmul_rn(c.sig, z)) : mp;
  } else {
    r.sample = r.mean;
  }
  return r;
}

struct StepPtrs { float *sample, *xs, *mean, *lv, *var; };

// Vector path.  The kernel is instruction-issue-bound if every thread re-derives the per-sample coefficients (14
// table gathers behind a dependent t[n] load and, for DDIM, five IEEE sqrt/div: ~390 warp instructions per float4,
// ncu in profiles/r01m_k8_reverse_step.txt), so each CTA owns a CONTIGUOUS run of 256-float4 tiles, derives the
// coefficients of the few samples that run touches once (one thread per sample, in parallel) into shared memory,
// and then streams its tiles with nothing but the element math in the loop.
constexpr int RS_MAX_SAMPLES = 128;  // samples one CTA's run may touch (host sizes the grid accordingly)

template <typename OutT, int MODE>
__global__ void __launch_bounds__(256, 4)
reverse_step_tile_kernel(const OutT* __restrict__ out, long long out_stride, const float* __restrict__ x,
                         const float* __restrict__ noise, const long long* __restrict__ t,
                         const float* __restrict__ tab, int T, StepPtrs dst, int mean_type, int var_type, float eta,
                         int clip, long long N, long long chw, long long tiles_per_sample, long long tiles_per_cta) {
  constexpr bool BF = !std::is_same<OutT, float>::value;
  __shared__ StepCoef sc[RS_MAX_SAMPLES];
  const bool learned = var_type == VT_LEARNED || var_type == VT_LEARNED_RANGE;
  const long long chw4 = chw >> 2, os4 = out_stride >> 2, tiles = N * tiles_per_sample;
  const long long t0 = (long long)blockIdx.x * tiles_per_cta;
  const long long t1 = t0 + tiles_per_cta < tiles ? t0 + tiles_per_cta : tiles;
  if (t0 >= t1) return;
  const long long n0 = t0 / tiles_per_sample, n1 = (t1 - 1) / tiles_per_sample;
  for (long long s = threadIdx.x; s <= n1 - n0; s += blockDim.x)
    sc[s] = load_step_coef(tab, T, t[n0 + s], mean_type, MODE, eta);
  __syncthreads();
  long long n = n0, k = t0 - n0 * tiles_per_sample;  // tile = n * tiles_per_sample + k
#pragma unroll 2
  for (long long tile = t0; tile < t1; ++tile) {
    const long long r = k * 256 + threadIdx.x;
    if (r < chw4) {
      const long long i = n * chw4 + r;
      const float4 o = Vec4<OutT>::load(out, n * os4 + r);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f), z = v;
      if (learned) v = Vec4<OutT>::load(out, n * os4 + chw4 + r);
      const float4 xv = ldg_stream_f4(reinterpret_cast<const float4*>(x) + i);
      if (noise) z = ldg_stream_f4(reinterpret_cast<const float4*>(noise) + i);
      const StepCoef& c = sc[n - n0];
      const StepOut a = step_elem<BF>(c, o.x, v.x, xv.x, z.x, mean_type, var_type, MODE, clip != 0);
      const StepOut b = step_elem<BF>(c, o.y, v.y, xv.y, z.y, mean_type, var_type, MODE, clip != 0);
      const StepOut d = step_elem<BF>(c, o.z, v.z, xv.z, z.z, mean_type, var_type, MODE, clip != 0);
      const StepOut e = step_elem<BF>(c, o.w, v.w, xv.w, z.w, mean_type, var_type, MODE, clip != 0);
      if (dst.sample) stg_stream_f4(reinterpret_cast<float4*>(dst.sample) + i, make_float4(a.sample, b.sample, d.sample, e.sample));
      if (dst.xs) stg_stream_f4(reinterpret_cast<float4*>(dst.xs) + i, make_float4(a.xs, b.xs, d.xs, e.xs));
      if (dst.mean) stg_stream_f4(reinterpret_cast<float4*>(dst.mean) + i, make_float4(a.mean, b.mean, d.mean, e.mean));
      if (dst.lv) stg_stream_f4(reinterpret_cast<float4*>(dst.lv) + i, make_float4(a.lv, b.lv, d.lv, e.lv));
      if (dst.var) stg_stream_f4(reinterpret_cast<float4*>(dst.var) + i, make_float4(a.var, b.var, d.var, e.var));
    }
    if (++k == tiles_per_sample) { k = 0; ++n; }
  }
}

// Scalar path (chw not a multiple of 4, or unaligned views).
template <typename OutT>
__global__ void __launch_bounds__(256)
reverse_step_scalar_kernel(const OutT* __restrict__ out, long long out_stride, const float* __restrict__ x,
                           const float* __restrict__ noise, const long long* __restrict__ t,
                           const float* __restrict__ tab, int T, StepPtrs dst, int mean_type, int var_type, int mode,
                           float eta, int clip, long long N, long long chw) {
  constexpr bool BF = !std::is_same<OutT, float>::value;
  const bool learned = var_type == VT_LEARNED || var_type == VT_LEARNED_RANGE;
  const long long stride = (long long)gridDim.x * blockDim.x, total = N * chw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long n = i / chw, r = i - n * chw;
    const StepCoef c = load_step_coef(tab, T, t[n], mean_type, mode, eta);
    const float o = (float)out[n * out_stride + r];
    const float v = learned ? (float)out[n * out_stride + chw + r] : 0.f;
    const StepOut a = step_elem<BF>(c, o, v, x[i], noise ? noise[i] : 0.f, mean_type, var_type, mode, clip != 0);
    if (dst.sample) dst.sample[i] = a.sample;
    if (dst.xs) dst.xs[i] = a.xs;
    if (dst.mean) dst.mean[i] = a.mean;
    if (dst.lv) dst.lv[i] = a.lv;
    if (dst.var) dst.var[i] = a.var;
  }
}

// IntervalCFG combine (tools/sampler.py:46-48): uncond + scale * (cond - uncond) on the two halves of a doubled
// batch, each op rounded in the tensor's own dtype as the eager expression does.
template <bool BF>
__device__ __forceinline__ float cfg_elem(float c, float u, float scale) {
  float d = __fsub_rn(c, u);
  if (BF) d = bf16r(d);
  float m = __fmul_rn(scale, d);
  if (BF) m = bf16r(m);
  return __fadd_rn(u, m);  // the store rounds to the tensor dtype
}

template <typename T, bool VEC>
__global__ void __launch_bounds__(256)
cfg_combine_kernel(const T* __restrict__ both, T* __restrict__ y, float scale, long long half) {
  constexpr bool BF = !std::is_same<T, float>::value;
  const long long stride = (long long)gridDim.x * blockDim.x;
  if (VEC) {
    const long long half4 = half >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < half4; i += stride) {
      const float4 c = Vec4<T>::load(both, i), u = Vec4<T>::load(both, half4 + i);
      Vec4<T>::store(y, i, make_float4(cfg_elem<BF>(c.x, u.x, scale), cfg_elem<BF>(c.y, u.y, scale),
                                       cfg_elem<BF>(c.z, u.z, scale), cfg_elem<BF>(c.w, u.w, scale)));
    }
  } else {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += stride)
      y[i] = (T)cfg_elem<BF>((float)both[i], (float)both[half + i], scale);
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" int vaw_sample_from_latent(const float* latent, const float* eps, float* out, long long N, long long chw,
                                      float scale, cudaStream_t stream) {
  VAW_CHECK_ARG(latent && eps && out && N >= 0 && chw >= 0, "vaw_sample_from_latent: bad arguments");
  if (N * chw == 0) return VAW_OK;
  long long blocks = (N * chw + 255) / 256;
  const long long cap = (long long)vaw_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  sample_from_latent_kernel<<<(unsigned)blocks, 256, 0, stream>>>(latent, eps, out, N, chw, scale);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

extern "C" int vaw_reverse_step(const void* model_out, int out_dtype, long long out_stride, const float* x,
                                const float* noise, const long long* t, const float* tab, int T, float* sample,
                                float* pred_xstart, float* mean, float* log_variance, float* variance, int mean_type,
                                int var_type, int mode, float eta, int clip, long long N, long long chw,
                                cudaStream_t stream) {
  VAW_CHECK_ARG(model_out && x && t && tab && T > 0, "vaw_reverse_step: null pointer");
  VAW_CHECK_ARG(out_dtype == 0 || out_dtype == 1, "vaw_reverse_step: out_dtype must be 0 (f32) or 1 (bf16)");
  VAW_CHECK_ARG(N >= 0 && chw > 0, "vaw_reverse_step: bad shape N=%lld chw=%lld", N, chw);
  VAW_CHECK_ARG(mean_type >= MT_PREVIOUS_X && mean_type <= MT_VELOCITY,
                "vaw_reverse_step: mean_type %d has no reverse step (NotImplementedError in the reference)", mean_type);
  VAW_CHECK_ARG(var_type >= VT_LEARNED && var_type <= VT_LEARNED_RANGE, "vaw_reverse_step: bad var_type %d", var_type);
  VAW_CHECK_ARG(mode >= RS_DDPM && mode <= RS_MOMENTS, "vaw_reverse_step: bad mode %d", mode);
  const bool learned = var_type == VT_LEARNED || var_type == VT_LEARNED_RANGE;
  VAW_CHECK_ARG(out_stride == (learned ? 2 : 1) * chw,
                "vaw_reverse_step: model output has %lld values per sample, expected %lld", out_stride,
                (learned ? 2 : 1) * chw);
  VAW_CHECK_ARG(noise || (mode != RS_DDPM && !(mode == RS_DDIM && eta != 0.f)), "vaw_reverse_step: this mode needs noise");
  VAW_CHECK_ARG(mode != RS_DDIM_REVERSE || eta == 0.f, "Reverse ODE only for deterministic path");
  VAW_CHECK_ARG(sample || pred_xstart || mean || log_variance || variance, "vaw_reverse_step: no output requested");
  if (N == 0) return VAW_OK;
  if (mode == RS_DDIM && eta == 0.f) noise = nullptr;  // sigma == 0: the reference adds 0 * noise
  const uintptr_t al = (uintptr_t)model_out | (uintptr_t)x | (uintptr_t)noise | (uintptr_t)sample |
                       (uintptr_t)pred_xstart | (uintptr_t)mean | (uintptr_t)log_variance | (uintptr_t)variance;
  const bool vec = (chw % 4 == 0) && ((al & 15) == 0);
  const StepPtrs dst{sample, pred_xstart, mean, log_variance, variance};
  const long long cap = (long long)vaw_num_sms() * 8;
  if (vec) {
    const long long tps = (chw / 4 + 255) / 256, tiles = N * tps;
    long long blocks = tiles < cap ? tiles : cap;
    const long long min_blocks = (N + RS_MAX_SAMPLES - 3) / (RS_MAX_SAMPLES - 2);  // a run spans <= per/tps + 2 samples
    if (blocks < min_blocks) blocks = min_blocks;
    const long long per = (tiles + blocks - 1) / blocks;
    blocks = (tiles + per - 1) / per;
#define VAW_RS_LAUNCH(TY, MD)                                                                                  \
  reverse_step_tile_kernel<TY, MD><<<(unsigned)blocks, 256, 0, stream>>>(                                      \
      (const TY*)model_out, out_stride, x, noise, t, tab, T, dst, mean_type, var_type, eta, clip, N, chw, tps, per)
#define VAW_RS_PICK(TY)                                                \
  do {                                                                 \
    if (mode == RS_DDPM) VAW_RS_LAUNCH(TY, RS_DDPM);                   \
    else if (mode == RS_DDIM) VAW_RS_LAUNCH(TY, RS_DDIM);              \
    else if (mode == RS_DDIM_REVERSE) VAW_RS_LAUNCH(TY, RS_DDIM_REVERSE); \
    else VAW_RS_LAUNCH(TY, RS_MOMENTS);                                \
  } while (0)
    if (out_dtype == 0) VAW_RS_PICK(float); else VAW_RS_PICK(bf16);
#undef VAW_RS_PICK
#undef VAW_RS_LAUNCH
  } else {
    long long blocks = (N * chw + 255) / 256;
    if (blocks > 2 * cap) blocks = 2 * cap;
#define VAW_RS_LAUNCH(TY)                                                                                           \
  reverse_step_scalar_kernel<TY><<<(unsigned)blocks, 256, 0, stream>>>((const TY*)model_out, out_stride, x, noise, t, \
                                                                       tab, T, dst, mean_type, var_type, mode, eta,  \
                                                                       clip, N, chw)
    if (out_dtype == 0) VAW_RS_LAUNCH(float); else VAW_RS_LAUNCH(bf16);
#undef VAW_RS_LAUNCH
  }
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

extern "C" int vaw_cfg_combine(const void* both, void* y, int dtype, float scale, long long half,
                               cudaStream_t stream) {
  VAW_CHECK_ARG(both && y && half >= 0, "vaw_cfg_combine: bad arguments");
  VAW_CHECK_ARG(dtype == 0 || dtype == 1, "vaw_cfg_combine: dtype must be 0 (f32) or 1 (bf16)");
  if (half == 0) return VAW_OK;
  const bool vec = (half % 4 == 0) && ((((uintptr_t)both | (uintptr_t)y) & 15) == 0) &&
                   ((half * (dtype == 0 ? 4 : 2)) % 16 == 0);
  long long blocks = ((vec ? half / 4 : half) + 255) / 256;
  const long long cap = (long long)vaw_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (dtype == 0) {
    if (vec) cfg_combine_kernel<float, true><<<(unsigned)blocks, 256, 0, stream>>>((const float*)both, (float*)y, scale, half);
    else cfg_combine_kernel<float, false><<<(unsigned)blocks, 256, 0, stream>>>((const float*)both, (float*)y, scale, half);
  } else {
    if (vec) cfg_combine_kernel<bf16, true><<<(unsigned)blocks, 256, 0, stream>>>((const bf16*)both, (bf16*)y, scale, half);
    else cfg_combine_kernel<bf16, false><<<(unsigned)blocks, 256, 0, stream>>>((const bf16*)both, (bf16*)y, scale, half);
  }
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

extern "C" int vaw_qsample_target(const float* x0, const float* noise, const long long* t, const float* tab_alpha,
                                  const float* tab_sigma, const float* tab_c0, const float* tab_c1, float* x_t,
                                  float* target, int mean_type, long long N, long long chw, cudaStream_t stream) {
  VAW_CHECK_ARG(x0 && noise && tab_alpha && tab_sigma && x_t, "vaw_qsample_target: null pointer");
  VAW_CHECK_ARG(N >= 0 && chw > 0, "vaw_qsample_target: bad shape N=%lld chw=%lld", N, chw);
  VAW_CHECK_ARG(mean_type >= MT_PREVIOUS_X && mean_type <= MT_SCORE, "vaw_qsample_target: bad mean_type %d",
                mean_type);
  if (N == 0) return VAW_OK;
  const bool vec = (chw % 4 == 0) && ((((uintptr_t)x0 | (uintptr_t)noise | (uintptr_t)x_t | (uintptr_t)target) & 15) == 0);
  const long long work = vec ? N * (chw / 4) : N * chw;
  long long blocks = (work + 255) / 256;
  const long long cap = (long long)vaw_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (vec)
    qsample_target_kernel<true><<<(unsigned)blocks, 256, 0, stream>>>(x0, noise, t, tab_alpha, tab_sigma, tab_c0,
                                                                      tab_c1, x_t, target, mean_type, N, chw);
  else
    qsample_target_kernel<false><<<(unsigned)blocks, 256, 0, stream>>>(x0, noise, t, tab_alpha, tab_sigma, tab_c0,
                                                                       tab_c1, x_t, target, mean_type, N, chw);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

extern "C" int vaw_wmse_fwd_bwd_strided(const void* out, int out_dtype, long long out_stride, const float* x0,
                                        const float* noise, const long long* t, const float* tab_alpha,
                                        const float* tab_sigma, const float* tab_c0, const float* tab_c1,
                                        const float* w_tab, float* mse, float* raw_mse, void* grad_out,
                                        long long grad_stride, const float* gscale_n, float gscale, int mean_type,
                                        long long N, long long chw, cudaStream_t stream) {
  VAW_CHECK_ARG(out && x0 && noise && tab_alpha && tab_sigma && mse, "vaw_wmse_fwd_bwd: null pointer");
  VAW_CHECK_ARG(out_stride >= chw && (!grad_out || grad_stride >= chw), "vaw_wmse_fwd_bwd: row strides below chw");
  VAW_CHECK_ARG(out_dtype == 0 || out_dtype == 1, "vaw_wmse_fwd_bwd: out_dtype must be 0 (f32) or 1 (bf16)");
  VAW_CHECK_ARG(N >= 0 && chw > 0, "vaw_wmse_fwd_bwd: bad shape N=%lld chw=%lld", N, chw);
  VAW_CHECK_ARG(mean_type >= MT_PREVIOUS_X && mean_type <= MT_SCORE, "vaw_wmse_fwd_bwd: bad mean_type %d", mean_type);
  if (N == 0) return VAW_OK;
  // 128-bit (fp32) / 64-bit (bf16) accesses need chw % 4 == 0 AND suitably aligned bases (an offset view of a larger
  // buffer may not be); otherwise the scalar path runs
  const uintptr_t f32_ptrs = (uintptr_t)x0 | (uintptr_t)noise;
  const uintptr_t out_ptrs = (uintptr_t)out | (uintptr_t)grad_out;
  const int vec = (chw % 4 == 0) && ((f32_ptrs & 15) == 0) && ((out_ptrs & (out_dtype == 0 ? 15 : 7)) == 0) &&
                  (out_stride % 4 == 0) && (grad_stride % 4 == 0);
  if (out_dtype == 0)
    wmse_fwd_bwd_kernel<float><<<(unsigned)N, 256, 0, stream>>>(
        (const float*)out, x0, noise, t, tab_alpha, tab_sigma, tab_c0, tab_c1, w_tab, mse, raw_mse, (float*)grad_out,
        gscale_n, gscale, mean_type, chw, vec, out_stride, grad_out ? grad_stride : chw);
  else
    wmse_fwd_bwd_kernel<bf16><<<(unsigned)N, 256, 0, stream>>>(
        (const bf16*)out, x0, noise, t, tab_alpha, tab_sigma, tab_c0, tab_c1, w_tab, mse, raw_mse, (bf16*)grad_out,
        gscale_n, gscale, mean_type, chw, vec, out_stride, grad_out ? grad_stride : chw);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

extern "C" int vaw_wmse_fwd_bwd(const void* out, int out_dtype, const float* x0, const float* noise,
                                const long long* t, const float* tab_alpha, const float* tab_sigma,
                                const float* tab_c0, const float* tab_c1, const float* w_tab, float* mse,
                                float* raw_mse, void* grad_out, const float* gscale_n, float gscale, int mean_type,
                                long long N, long long chw, cudaStream_t stream) {
  return vaw_wmse_fwd_bwd_strided(out, out_dtype, chw, x0, noise, t, tab_alpha, tab_sigma, tab_c0, tab_c1, w_tab, mse,
                                  raw_mse, grad_out, chw, gscale_n, gscale, mean_type, N, chw, stream);
}

extern "C" int vaw_scale_rows(const void* x, const float* s, void* y, int dtype, long long N, long long chw,
                              cudaStream_t stream) {
  VAW_CHECK_ARG(x && s && y, "vaw_scale_rows: null pointer");
  VAW_CHECK_ARG(dtype == 0 || dtype == 1, "vaw_scale_rows: dtype must be 0 (f32) or 1 (bf16)");
  if (N * chw == 0) return VAW_OK;
  long long blocks = (N * chw + 255) / 256;
  const long long cap = (long long)vaw_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (dtype == 0)
    scale_rows_kernel<float><<<(unsigned)blocks, 256, 0, stream>>>((const float*)x, s, (float*)y, N, chw);
  else
    scale_rows_kernel<bf16><<<(unsigned)blocks, 256, 0, stream>>>((const bf16*)x, s, (bf16*)y, N, chw);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

// ------------------------------------------------------------------------------------------------
// Host: per-timestep loss-weight LUT.  compute_mse_loss_weight (reference :1092-1148) depends on t only,
// so w[t] is evaluated once per diffusion object with the reference's fp32 operation order:
//   alpha = float(sqrt_ac[t]); sigma = float(sqrt_1mac[t]); snr = (alpha/sigma)^2 (fp32) ...
// kind codes: see include/vaw_b200.h (VAW_W_*).  Returns VAW_ERR_INVALID for combinations the reference
// rejects with ValueError (:1144-1145).
// ------------------------------------------------------------------------------------------------
extern "C" int vaw_loss_weight_lut(const double* sqrt_ac, const double* sqrt_1mac, int T, int mean_type,
                                   int weight_kind, double k, double p2_k, double p2_gamma, float* lut) {
  VAW_CHECK_ARG(sqrt_ac && sqrt_1mac && lut && T > 0, "vaw_loss_weight_lut: bad arguments");
  enum { W_CONSTANT = 0, W_LAMBDA, W_MIN_SNR, W_MAX_SNR, W_DEBIAS, W_MIN_DEBIAS, W_MAX_DEBIAS, W_P2, W_TRUNC_SNR,
         W_SNR, W_INV_SNR };
  for (int i = 0; i < T; ++i) {
    volatile float alpha = (float)sqrt_ac[i];
    volatile float sigma = (float)sqrt_1mac[i];
    volatile float q = alpha / sigma;
    volatile float snr = q * q;
    const float kf = (float)k;
    float w = 0.f;
    bool ok = false;
    if (weight_kind == W_CONSTANT) {
      lut[i] = 1.f;  // early return in the reference: no snr==0 fix-up
      continue;
    }
    if (mean_type == MT_EPSILON) {
      ok = true;
      switch (weight_kind) {
        case W_MIN_SNR: { volatile float m = fminf(snr, kf); w = m / snr; } break;
        case W_MAX_SNR: { volatile float m = fmaxf(snr, kf); w = m / snr; } break;
        case W_LAMBDA: w = sigma; break;
        case W_DEBIAS: w = sigma / alpha; break;
        case W_P2: { volatile float b = (float)p2_k + snr; volatile float p = powf(b, (float)p2_gamma); w = 1.f / p; } break;
        case W_MIN_DEBIAS: { volatile float r = sigma / alpha; w = fminf(r, 1.f); } break;
        case W_MAX_DEBIAS: { volatile float r = sigma / alpha; w = fmaxf(r, 1.f); } break;
        default: ok = false;
      }
    } else if (mean_type == MT_START_X) {
      ok = true;
      switch (weight_kind) {
        case W_TRUNC_SNR: w = fmaxf(snr, 1.f); break;
        case W_SNR: w = snr; break;
        case W_INV_SNR: w = 1.f / snr; break;
        case W_MIN_SNR: w = fminf(snr, kf); break;
        case W_MAX_SNR: w = fmaxf(snr, kf); break;
        case W_LAMBDA: w = alpha; break;
        default: ok = false;
      }
    } else if (mean_type == MT_VECTOR) {
      if (weight_kind == W_LAMBDA) { ok = true; w = 1.f; }
    } else if (mean_type == MT_VELOCITY) {
      ok = true;
      switch (weight_kind) {
        case W_MIN_SNR: { volatile float m = fminf(snr, kf); volatile float d = snr + 1.f; w = m / d; } break;
        case W_LAMBDA: w = alpha * sigma; break;
        default: ok = false;
      }
    }
    if (!ok) {
      vaw_set_error("Invalid mse_loss_weight_type: kind=%d for mean_type=%d", weight_kind, mean_type);
      return VAW_ERR_INVALID;
    }
    if (snr == 0.f) w = 1.f;  // reference :1147
    lut[i] = w;
  }
  return VAW_OK;
}

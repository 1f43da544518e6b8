// philox_kernels.cu — K1 / K2 with the Gaussian noise drawn INSIDE the kernels, bit-compatible with the tensors
// `torch.randn_like` / `torch.randint` would have produced from the same CUDA generator state.
//
// The reference's step draws, in this order, from the device generator (tools/trainer.py:21-25,
// tools/gaussian_diffusion.py:849-852):
//     eps1  = randn_like(mean)      sample_from_latent (latent models only)
//     noise = randn_like(x_start)   training_losses
//     t     = randint(0, T, (N,))   sample_t
// ATen fills such tensors with distribution_elementwise_grid_stride_kernel (ATen/native/cuda/DistributionTemplates.h):
// block 256, grid = min(SMs * (maxThreadsPerSM / 256), ceil(numel / 256)); thread `idx` owns Philox4x32-10 subsequence
// `idx` of (seed, offset), and its k-th curand_normal4 / curand4 call supplies elements idx + G (4k + ii), ii = 0..3,
// G = grid * 256.  Reproducing that mapping with cuRAND's device API gives the identical values, so the noise tensor
// never has to exist in HBM: K1 consumes it as it is drawn (12 -> 8 B/element) and K2 re-draws the value it needs for
// the eps target instead of reading it back (8 -> 4 B/element fp32-out).  The host side advances the generator offset by
// exactly what the ATen kernels would have consumed, so every later draw of the program is unchanged.
#include <curand_kernel.h>

#include "vaw_common.cuh"

namespace {

enum : int { MT_PREVIOUS_X = 1, MT_START_X = 2, MT_EPSILON = 3, MT_VELOCITY = 4, MT_VECTOR = 5, MT_SCORE = 6 };

__device__ __forceinline__ float mix2(float p, float u, float q, float v) {
  return __fadd_rn(__fmul_rn(p, u), __fmul_rn(q, v));
}

__device__ __forceinline__ float target_of(int mt, float a, float s, float c0, float c1, float x0, float eps, float xt) {
  switch (mt) {
    case MT_START_X: return x0;
    case MT_EPSILON: return eps;
    case MT_VELOCITY: return __fsub_rn(__fmul_rn(a, eps), __fmul_rn(s, x0));
    case MT_PREVIOUS_X: return mix2(c0, x0, c1, xt);
    case MT_VECTOR: return mix2(c0, x0, c1, eps);
    case MT_SCORE: return __fdiv_rn(-eps, s);
    default: return eps;
  }
}

// element `li` of a tensor of `numel` normals filled by ATen with (seed, offset) on a grid of G threads
__device__ __forceinline__ float aten_normal_at(unsigned long long seed, unsigned long long offset, long long G,
                                                long long li) {
  const long long idx = li % G;
  const long long q = li / G;          // = 4 * call + ii
  curandStatePhilox4_32_10_t st;
  curand_init(seed, (unsigned long long)idx, offset + 4ULL * (unsigned long long)(q >> 2), &st);
  const float4 r = curand_normal4(&st);
  const int ii = (int)(q & 3);
  return ii == 0 ? r.x : ii == 1 ? r.y : ii == 2 ? r.z : r.w;
}

// K1 on ATen's thread / element mapping.  x0 given, or rebuilt from an 8-channel latent (mean | std) with its own draw
// (sample_from_latent, trainer.py:21-25): x0 = fl(fl(mean + fl(std * eps1)) * scale).
__global__ void __launch_bounds__(256)
qsample_philox_kernel(const float* __restrict__ x0, const float* __restrict__ latent, float latent_scale,
                      unsigned long long seed, unsigned long long off_latent, unsigned long long off_noise,
                      const long long* __restrict__ t, const float* __restrict__ tab_a, const float* __restrict__ tab_s,
                      const float* __restrict__ tab_c0, const float* __restrict__ tab_c1, float* __restrict__ x_start_out,
                      float* __restrict__ noise_out, float* __restrict__ x_t, float* __restrict__ target, int mean_type,
                      long long numel, long long chw) {
  const long long G = (long long)gridDim.x * blockDim.x;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  curandStatePhilox4_32_10_t sn, sl;
  curand_init(seed, (unsigned long long)idx, off_noise, &sn);
  if (latent) curand_init(seed, (unsigned long long)idx, off_latent, &sl);
  const long long rounded = ((numel - 1) / (G * 4) + 1) * G * 4;
  for (long long base = idx; base < rounded; base += G * 4) {
    const float4 rn = curand_normal4(&sn);
    float4 rl = make_float4(0.f, 0.f, 0.f, 0.f);
    if (latent) rl = curand_normal4(&sl);
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const long long li = base + G * ii;
      if (li >= numel) continue;
      const float eps = ii == 0 ? rn.x : ii == 1 ? rn.y : ii == 2 ? rn.z : rn.w;
      const long long n = li / chw;
      float x;
      if (latent) {
        const float e1 = ii == 0 ? rl.x : ii == 1 ? rl.y : ii == 2 ? rl.z : rl.w;
        const long long r = li - n * chw;
        const float mean = latent[n * 2 * chw + r], sd = latent[n * 2 * chw + chw + r];
        x = __fmul_rn(__fadd_rn(mean, __fmul_rn(sd, e1)), latent_scale);
      } else {
        x = x0[li];
      }
      const long long ti = t ? t[n] : n;
      const float a = __ldg(tab_a + ti), s = __ldg(tab_s + ti);
      const float xt = mix2(a, x, s, eps);
      x_t[li] = xt;
      if (x_start_out) x_start_out[li] = x;
      if (noise_out) noise_out[li] = eps;
      if (target) {
        const float c0 = tab_c0 ? __ldg(tab_c0 + ti) : 0.f, c1 = tab_c1 ? __ldg(tab_c1 + ti) : 0.f;
        target[li] = target_of(mean_type, a, s, c0, c1, x, eps, xt);
      }
    }
  }
}

// torch.randint(low, low + range, (n,)) with range < 2^32 (random_from_to_kernel: curand4, `rand % range + base`)
__global__ void __launch_bounds__(256)
randint_philox_kernel(unsigned long long seed, unsigned long long offset, unsigned int range, long long base_value,
                      long long* __restrict__ out, long long numel) {
  const long long G = (long long)gridDim.x * blockDim.x;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  curandStatePhilox4_32_10_t st;
  curand_init(seed, (unsigned long long)idx, offset, &st);
  const long long rounded = ((numel - 1) / (G * 4) + 1) * G * 4;
  for (long long b = idx; b < rounded; b += G * 4) {
    const uint4 r = curand4(&st);
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const long long li = b + G * ii;
      if (li >= numel) continue;
      const unsigned int v = ii == 0 ? r.x : ii == 1 ? r.y : ii == 2 ? r.z : r.w;
      out[li] = (long long)(v % range) + base_value;
    }
  }
}

// K2 that re-draws the noise element it needs instead of reading a noise tensor (one CTA per sample, as K2).
template <typename OutT>
__global__ void __launch_bounds__(256)
wmse_philox_kernel(const OutT* __restrict__ out, long long out_stride, const float* __restrict__ x0,
                   unsigned long long seed, unsigned long long off_noise, long long G, const long long* __restrict__ t,
                   const float* __restrict__ tab_a, const float* __restrict__ tab_s, const float* __restrict__ tab_c0,
                   const float* __restrict__ tab_c1, const float* __restrict__ w_tab, float* __restrict__ mse,
                   OutT* __restrict__ grad, long long grad_stride, float gscale, int mean_type, long long chw, int vec) {
  const long long n = blockIdx.x;
  const long long ti = t ? t[n] : n;
  const float a = __ldg(tab_a + ti), s = __ldg(tab_s + ti);
  const float c0 = tab_c0 ? __ldg(tab_c0 + ti) : 0.f, c1 = tab_c1 ? __ldg(tab_c1 + ti) : 0.f;
  const float w = w_tab ? __ldg(w_tab + ti) : 1.f;
  const float inv = 1.0f / (float)chw;
  const float g = gscale * w * 2.f * inv;
  const bool need_x0 = (mean_type != MT_EPSILON && mean_type != MT_SCORE);
  const bool need_eps = (mean_type != MT_START_X);
  const OutT* on = out + n * out_stride;
  OutT* gn = grad ? grad + n * grad_stride : nullptr;
  float acc = 0.f;
  auto one = [&](long long i) {
    const long long li = n * chw + i;
    const float o = (float)on[i];
    const float x = need_x0 ? x0[li] : 0.f;
    const float e = need_eps ? aten_normal_at(seed, off_noise, G, li) : 0.f;
    const float tg = target_of(mean_type, a, s, c0, c1, x, e, mix2(a, x, s, e));
    const float d = tg - o;
    acc += d * d;
    if (gn) gn[i] = (OutT)(-g * d);
  };
  if (vec) {
    // same thread -> element assignment and the same accumulation order as K2's 128-bit path (wmse_fwd_bwd_kernel), so
    // a step with in-kernel noise returns bit-identical losses to the step fed the equivalent noise tensor
    for (long long i4 = threadIdx.x; i4 < (chw >> 2); i4 += blockDim.x) {
#pragma unroll
      for (int k = 0; k < 4; ++k) one(4 * i4 + k);
    }
  } else {
    for (long long i = threadIdx.x; i < chw; i += blockDim.x) one(i);
  }
  __shared__ float red[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) sum += red[i];
    mse[n] = w * (sum * inv);
  }
}

// ATen's launch geometry for a distribution kernel over `numel` elements on the current device
int aten_grid(long long numel, unsigned* grid_out) {
  int dev = 0;
  VAW_CUDA_TRY(cudaGetDevice(&dev));
  static int cached_dev = -1, sms = 0, max_threads = 0;
  if (cached_dev != dev) {
    VAW_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    VAW_CUDA_TRY(cudaDeviceGetAttribute(&max_threads, cudaDevAttrMaxThreadsPerMultiProcessor, dev));
    cached_dev = dev;
  }
  unsigned long long grid = (unsigned long long)((numel + 255) / 256);
  const unsigned long long cap = (unsigned long long)sms * (unsigned long long)(max_threads / 256);
  if (grid > cap) grid = cap;
  *grid_out = (unsigned)grid;
  return VAW_OK;
}

}  // namespace

// The generator offset ATen would consume for a float normal / 32-bit integer draw of `numel` elements (always a multiple
// of 4): the host advances torch's generator by this amount after each in-kernel draw.
extern "C" int vaw_philox_offset_increment(long long numel, unsigned long long* increment) {
  VAW_CHECK_ARG(increment && numel >= 0, "vaw_philox_offset_increment: bad arguments");
  if (numel == 0) { *increment = 0; return VAW_OK; }
  unsigned grid = 0;
  int rc = aten_grid(numel, &grid);
  if (rc) return rc;
  *increment = (unsigned long long)(((numel - 1) / (256LL * grid * 4) + 1) * 4);
  return VAW_OK;
}

extern "C" int vaw_qsample_philox(const float* x0, const float* latent, float latent_scale, unsigned long long seed,
                                  unsigned long long offset_latent, unsigned long long offset_noise, const long long* t,
                                  const float* tab_alpha, const float* tab_sigma, const float* tab_c0,
                                  const float* tab_c1, float* x_start_out, float* noise_out, float* x_t, float* target,
                                  int mean_type, long long N, long long chw, cudaStream_t stream) {
  VAW_CHECK_ARG((x0 != nullptr) != (latent != nullptr), "vaw_qsample_philox: pass exactly one of x0 / latent");
  VAW_CHECK_ARG(tab_alpha && tab_sigma && x_t && N >= 0 && chw > 0, "vaw_qsample_philox: bad arguments");
  VAW_CHECK_ARG(mean_type >= MT_PREVIOUS_X && mean_type <= MT_SCORE, "vaw_qsample_philox: bad mean_type %d", mean_type);
  if (N == 0) return VAW_OK;
  unsigned grid = 0;
  int rc = aten_grid(N * chw, &grid);
  if (rc) return rc;
  qsample_philox_kernel<<<grid, 256, 0, stream>>>(x0, latent, latent_scale, seed, offset_latent, offset_noise, t,
                                                  tab_alpha, tab_sigma, tab_c0, tab_c1, x_start_out, noise_out, x_t,
                                                  target, mean_type, N * chw, chw);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

extern "C" int vaw_randint_philox(unsigned long long seed, unsigned long long offset, long long low, long long high,
                                  long long* out, long long n, cudaStream_t stream) {
  VAW_CHECK_ARG(out && n >= 0 && high > low && (high - low) < (1LL << 32), "vaw_randint_philox: bad arguments");
  if (n == 0) return VAW_OK;
  unsigned grid = 0;
  int rc = aten_grid(n, &grid);
  if (rc) return rc;
  randint_philox_kernel<<<grid, 256, 0, stream>>>(seed, offset, (unsigned int)(high - low), low, out, n);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

extern "C" int vaw_wmse_fwd_bwd_philox(const void* out, int out_dtype, long long out_stride, const float* x0,
                                       unsigned long long seed, unsigned long long offset_noise, const long long* t,
                                       const float* tab_alpha, const float* tab_sigma, const float* tab_c0,
                                       const float* tab_c1, const float* w_tab, float* mse, void* grad_out,
                                       long long grad_stride, float gscale, int mean_type, long long N, long long chw,
                                       cudaStream_t stream) {
  VAW_CHECK_ARG(out && tab_alpha && tab_sigma && mse && N >= 0 && chw > 0, "vaw_wmse_fwd_bwd_philox: bad arguments");
  VAW_CHECK_ARG(out_dtype == 0 || out_dtype == 1, "vaw_wmse_fwd_bwd_philox: out_dtype must be 0 (f32) or 1 (bf16)");
  VAW_CHECK_ARG(mean_type >= MT_PREVIOUS_X && mean_type <= MT_SCORE, "vaw_wmse_fwd_bwd_philox: bad mean_type");
  VAW_CHECK_ARG(x0 || mean_type == MT_EPSILON || mean_type == MT_SCORE, "vaw_wmse_fwd_bwd_philox: this target needs x0");
  VAW_CHECK_ARG(out_stride >= chw && (!grad_out || grad_stride >= chw), "vaw_wmse_fwd_bwd_philox: row strides below chw");
  if (N == 0) return VAW_OK;
  unsigned grid = 0;
  int rc = aten_grid(N * chw, &grid);
  if (rc) return rc;
  const long long G = 256LL * grid;
  // mirrors the vector-path condition of vaw_wmse_fwd_bwd_strided (which decides the summation order)
  const uintptr_t f32_ptrs = (uintptr_t)x0;
  const uintptr_t out_ptrs = (uintptr_t)out | (uintptr_t)grad_out;
  const int vec = (chw % 4 == 0) && ((f32_ptrs & 15) == 0) && ((out_ptrs & (out_dtype == 0 ? 15 : 7)) == 0) &&
                  (out_stride % 4 == 0) && (grad_stride % 4 == 0);
  if (out_dtype == 0)
    wmse_philox_kernel<float><<<(unsigned)N, 256, 0, stream>>>((const float*)out, out_stride, x0, seed, offset_noise, G, t,
                                                              tab_alpha, tab_sigma, tab_c0, tab_c1, w_tab, mse,
                                                              (float*)grad_out, grad_stride, gscale, mean_type, chw, vec);
  else
    wmse_philox_kernel<bf16><<<(unsigned)N, 256, 0, stream>>>((const bf16*)out, out_stride, x0, seed, offset_noise, G, t,
                                                             tab_alpha, tab_sigma, tab_c0, tab_c1, w_tab, mse,
                                                             (bf16*)grad_out, grad_stride, gscale, mean_type, chw, vec);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

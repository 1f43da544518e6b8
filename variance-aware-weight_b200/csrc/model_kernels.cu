// model_kernels.cu — small kernels around the transformer blocks: (un)patchify, timestep / label conditioning,
// dtype casts of the flat parameter buffer, the fused AdamW(+bf16 shadow) update, and the REPA alignment loss.
// Reference: models/dit.py:41-110 (embedders), :243-256 (unpatchify), timm PatchEmbed,
// models/uvit.py:21-52, tools/gaussian_diffusion.py:1007-1013 (compute_align_loss, 'mse'), main.py:354 (AdamW).
#include <type_traits>

#include "vaw_common.cuh"

namespace {

// x [B, C, H, W] fp32 -> patches [B*T, C*P*P] bf16, feature order (c, p, q) = Conv2d weight flatten order
__global__ void __launch_bounds__(256)
patchify_in_kernel(const float* __restrict__ x, bf16* __restrict__ patches, int B, int C, int H, int W, int P) {
  const int Wg = W / P, Hg = H / P, Kp = C * P * P;
  const long long total = (long long)B * Hg * Wg * Kp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i % Kp);
    const long long tok = i / Kp;
    const int w = (int)(tok % Wg), h = (int)((tok / Wg) % Hg), n = (int)(tok / ((long long)Wg * Hg));
    const int q = f % P, p = (f / P) % P, c = f / (P * P);
    patches[i] = __float2bfloat16_rn(x[(((long long)n * C + c) * H + h * P + p) * W + w * P + q]);
  }
}

// REPA teacher input: preprocess_raw_image (tools/align_utils.py:19-40, mocov3 / mae / dinov1 branch)
// fused into the patchify: raw pixels 0..255 fp32 [B, C, H, W] -> ((x / 255) - mean[c]) / std[c] (each op rounded like
// the eager sequence x / 255., Normalize's sub_ and div_) -> bf16 patches [B*T, C*P*P] in Conv2d weight order.
__global__ void __launch_bounds__(256)
patchify_norm_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ stdv,
                     bf16* __restrict__ patches, int B, int C, int H, int W, int P) {
  const int Wg = W / P, Hg = H / P, Kp = C * P * P;
  const long long total = (long long)B * Hg * Wg * Kp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i % Kp);
    const long long tok = i / Kp;
    const int w = (int)(tok % Wg), h = (int)((tok / Wg) % Hg), n = (int)(tok / ((long long)Wg * Hg));
    const int q = f % P, p = (f / P) % P, c = f / (P * P);
    const float v = x[(((long long)n * C + c) * H + h * P + p) * W + w * P + q];
    patches[i] = __float2bfloat16_rn(__fdiv_rn(__fsub_rn(__fdiv_rn(v, 255.f), __ldg(mean + c)), __ldg(stdv + c)));
  }
}

// ViT token assembly (timm VisionTransformer._pos_embed): x[n, 0, :] = cls + pos[0], x[n, 1 + l, :] = tok[n, l, :]
// (the patch-embed GEMM has already added bias and pos[1 + l] through its residual-table epilogue).
__global__ void __launch_bounds__(256)
vit_assemble_kernel(const float* __restrict__ tok, const float* __restrict__ cls, const float* __restrict__ pos0,
                    float* __restrict__ x, int B, int L, int D) {
  const long long total = (long long)B * (L + 1) * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i % D);
    const long long row = i / D;
    const int l = (int)(row % (L + 1));
    const long long n = row / (L + 1);
    x[i] = l == 0 ? __fadd_rn(cls[d], pos0[d]) : tok[(n * L + (l - 1)) * D + d];
  }
}

// tokens [B*T, P*P*C] (feature order (p, q, c)) <-> image [B, C, H, W]   (dit.py:243-256, uvit.py:47-52)
template <typename T>
__global__ void __launch_bounds__(256)
unpatchify_kernel(T* tok, T* img, int B, int C, int H, int W, int P, int to_image) {
  const int Wg = W / P, Hg = H / P, F = C * P * P;
  const long long total = (long long)B * C * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xw = (int)(i % W), yh = (int)((i / W) % H), c = (int)((i / ((long long)W * H)) % C);
    const int n = (int)(i / ((long long)W * H * C));
    const int h = yh / P, p = yh % P, w = xw / P, q = xw % P;
    const long long ti = ((long long)n * Hg * Wg + (long long)h * Wg + w) * F + (p * P + q) * C + c;
    if (to_image) img[i] = tok[ti];
    else tok[ti] = img[i];
  }
}

// sinusoidal timestep features [B, dim] bf16: cat(cos(t f_k), sin(t f_k)), f_k = exp(-ln(10000) k / half)  (dit.py:53-73)
__global__ void timestep_embedding_kernel(const float* __restrict__ t, bf16* __restrict__ out, float* __restrict__ out_f32,
                                          int B, int dim) {
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * dim) return;
  const int n = i / dim, j = i % dim;
  float v = 0.f;
  if (j < 2 * half) {
    const int k = j < half ? j : j - half;
    const float freq = expf(-9.210340371976184f * (float)k / (float)half);
    const float a = t[n] * freq;
    v = j < half ? cosf(a) : sinf(a);
  }
  if (out) out[i] = __float2bfloat16_rn(v);
  if (out_f32) out_f32[i] = v;
}

// c = bf16round(t_emb) + table[label] ; c_silu = bf16(silu(c))                      (dit.py:261-263,118-131)
__global__ void cond_combine_kernel(const float* __restrict__ t_emb, const float* __restrict__ table,
                                    const long long* __restrict__ labels, float* __restrict__ c,
                                    bf16* __restrict__ c_silu, int B, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int n = i / D, j = i % D;
  float v = __bfloat162float(__float2bfloat16_rn(t_emb[i]));
  if (table) v += table[labels[n] * (long long)D + j];
  c[i] = v;
  c_silu[i] = __float2bfloat16_rn(silu_f(v));
}
// dc = dc_silu * silu'(c) ; written as fp32 and bf16
__global__ void cond_bwd_kernel(const float* __restrict__ dcs, const float* __restrict__ c, float* __restrict__ dc,
                                bf16* __restrict__ dc_bf16, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = dcs[i] * silu_grad_f(c[i]);
  dc[i] = v;
  dc_bf16[i] = __float2bfloat16_rn(v);
}
// dense embedding gradient, deterministic: one CTA per table row scans the batch in order
__global__ void embedding_grad_kernel(const float* __restrict__ dc, long long ld, const long long* __restrict__ labels,
                                      float* __restrict__ dtable, int B, int D, int accumulate) {
  const int r = blockIdx.x;
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    float s = 0.f;
    for (int n = 0; n < B; ++n)
      if (labels[n] == r) s += dc[(long long)n * ld + j];
    float* o = dtable + (long long)r * D + j;
    *o = accumulate ? *o + s : s;
  }
}

__global__ void __launch_bounds__(256)
cast_f32_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n) {
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = ldg_stream_f4(reinterpret_cast<const float4*>(src) + i);
    uint2 o;
    o.x = pack_bf16(v.x, v.y);
    o.y = pack_bf16(v.z, v.w);
    reinterpret_cast<uint2*>(dst)[i] = o;
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = __float2bfloat16_rn(src[i]);
}

// strided 2-D cast: dst[r, c] = bf16(src[r, c]) for a [rows, cols] window of row-major matrices
__global__ void __launch_bounds__(256)
cast_f32_bf16_2d_kernel(const float* __restrict__ src, long long lds, bf16* __restrict__ dst, long long ldd, int rows,
                        int cols) {
  const int c4 = cols >> 2;
  const long long total = (long long)rows * c4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / c4), c = (int)(i % c4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(src + (long long)r * lds + c);
    uint2 o;
    o.x = pack_bf16(v.x, v.y);
    o.y = pack_bf16(v.z, v.w);
    *reinterpret_cast<uint2*>(dst + (long long)r * ldd + c) = o;
  }
}

// dst(fp32) += src(bf16)
__global__ void __launch_bounds__(256)
add_bf16_into_f32_kernel(const bf16* __restrict__ src, float* __restrict__ dst, long long n) {
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const uint2 u = reinterpret_cast<const uint2*>(src)[i];
    const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
    float4 d = reinterpret_cast<float4*>(dst)[i];
    d.x += a.x; d.y += a.y; d.z += b.x; d.w += b.y;
    reinterpret_cast<float4*>(dst)[i] = d;
  }
}

// column sum of a small fp32 [rows, N] matrix (rows = batch): out[c] (+)= sum_r a[r, c]
__global__ void colsum_f32_small_kernel(const float* __restrict__ a, long long lda, int rows, int N,
                                        float* __restrict__ out, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  float s = 0.f;
  for (int r = 0; r < rows; ++r) s += a[(long long)r * lda + c];
  out[c] = accumulate ? out[c] + s : s;
}

// Fused AdamW over the flat parameter buffer (torch.optim.AdamW semantics, decoupled weight decay) that also
// refreshes the bf16 shadow used by the GEMMs and optionally the EMA copy (trainer.py:12-18).
__device__ __forceinline__ void adamw_span(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                           float* __restrict__ v, bf16* __restrict__ p_bf16, float* __restrict__ ema,
                                           long long n, long long first, long long stride, float lr, float beta1,
                                           float beta2, float eps, float wd, float bc1, float bc2_sqrt, float grad_scale,
                                           float ema_decay, const float* __restrict__ clip_coef,
                                           const float* __restrict__ inv_scale, const float* __restrict__ found_inf) {
  // GradScaler semantics without a host round trip (trainer.py:124-129): a step whose gradients held an inf / nan is
  // skipped as a whole - parameters, moments, the bf16 shadow and the EMA copy stay as they were
  if (found_inf && __ldg(found_inf) != 0.f) return;
  if (clip_coef) grad_scale *= __ldg(clip_coef);   // device-side clip_grad_norm_ coefficient (vaw_grad_clip_coef)
  if (inv_scale) grad_scale *= __ldg(inv_scale);   // 1 / loss scale, kept on the device by torch's GradScaler
  const long long n4 = n >> 2;
  for (long long i = first; i < n4; i += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = ldg_stream_f4(reinterpret_cast<const float4*>(g) + i);
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float pa[4] = {pv.x, pv.y, pv.z, pv.w};
    const float ga[4] = {gv.x * grad_scale, gv.y * grad_scale, gv.z * grad_scale, gv.w * grad_scale};
    float ma[4] = {mv.x, mv.y, mv.z, mv.w};
    float va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      pa[j] *= (1.f - lr * wd);
      ma[j] = beta1 * ma[j] + (1.f - beta1) * ga[j];
      va[j] = beta2 * va[j] + (1.f - beta2) * ga[j] * ga[j];
      const float denom = sqrtf(va[j]) / bc2_sqrt + eps;
      pa[j] -= (lr / bc1) * (ma[j] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = make_float4(pa[0], pa[1], pa[2], pa[3]);
    reinterpret_cast<float4*>(m)[i] = make_float4(ma[0], ma[1], ma[2], ma[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(va[0], va[1], va[2], va[3]);
    if (p_bf16) {
      uint2 o;
      o.x = pack_bf16(pa[0], pa[1]);
      o.y = pack_bf16(pa[2], pa[3]);
      reinterpret_cast<uint2*>(p_bf16)[i] = o;
    }
    if (ema) {
      float4 e = reinterpret_cast<float4*>(ema)[i];
      e.x = e.x * ema_decay + pa[0] * (1.f - ema_decay);
      e.y = e.y * ema_decay + pa[1] * (1.f - ema_decay);
      e.z = e.z * ema_decay + pa[2] * (1.f - ema_decay);
      e.w = e.w * ema_decay + pa[3] * (1.f - ema_decay);
      reinterpret_cast<float4*>(ema)[i] = e;
    }
  }
}

__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             bf16* __restrict__ p_bf16, float* __restrict__ ema, long long n, float lr, float beta1, float beta2,
             float eps, float wd, float bc1, float bc2_sqrt, float grad_scale, float ema_decay,
             const float* __restrict__ clip_coef, const float* __restrict__ inv_scale,
             const float* __restrict__ found_inf) {
  adamw_span(p, g, m, v, p_bf16, ema, n, (long long)blockIdx.x * blockDim.x + threadIdx.x,
             (long long)gridDim.x * blockDim.x, lr, beta1, beta2, eps, wd, bc1, bc2_sqrt, grad_scale, ema_decay, clip_coef,
             inv_scale, found_inf);
}

// The same update over a LIST of element ranges of the flat buffers in one launch (blockIdx.y = range): the sharded
// data-parallel optimizer owns ~60 slices per rank plus the replicated small tensors; one launch instead of ~120.
__global__ void __launch_bounds__(256)
adamw_ranges_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                    bf16* __restrict__ p_bf16, float* __restrict__ ema, const long long* __restrict__ ranges, float lr,
                    float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt, float grad_scale,
                    float ema_decay, const float* __restrict__ clip_coef, const float* __restrict__ inv_scale,
                    const float* __restrict__ found_inf) {
  const long long off = ranges[2 * blockIdx.y], n = ranges[2 * blockIdx.y + 1];
  adamw_span(p + off, g + off, m + off, v + off, p_bf16 ? p_bf16 + off : nullptr, ema ? ema + off : nullptr, n,
             (long long)blockIdx.x * blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x, lr, beta1, beta2, eps,
             wd, bc1, bc2_sqrt, grad_scale, ema_decay, clip_coef, inv_scale, found_inf);
}

// Global gradient norm of the flat gradient buffer and the clip_grad_norm_ coefficient (trainer.py:60-62 ->
// torch.nn.utils.clip_grad_norm_: coef = min(1, max_norm / (||g||_2 + 1e-6))), kept on the device so that the optimizer
// step needs no host synchronisation.  Two deterministic stages (fixed grid, fixed reduction tree).
constexpr int kNormBlocks = 1024;
__global__ void __launch_bounds__(256)
grad_sqnorm_stage1(const float* __restrict__ g, long long n, float scale, float* __restrict__ part) {
  __shared__ float red[8];
  const long long n4 = n >> 2;
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)kNormBlocks * 256) {
    const float4 v = ldg_stream_f4(reinterpret_cast<const float4*>(g) + i);
    const float a = v.x * scale, b = v.y * scale, c = v.z * scale, d = v.w * scale;
    acc += (a * a + b * b) + (c * c + d * d);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float a = g[(n4 << 2) + threadIdx.x] * scale;
    acc += a * a;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i];
    part[blockIdx.x] = s;
  }
}
__global__ void __launch_bounds__(256)
grad_sqnorm_stage2(const float* __restrict__ part, float max_norm, float* __restrict__ out) {
  __shared__ double red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < kNormBlocks; i += 256) s += (double)part[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float norm = (float)sqrt(red[0]);
    out[0] = norm;
    out[1] = max_norm > 0.f ? fminf(1.f, max_norm / (norm + 1e-6f)) : 1.f;
  }
}

// REPA alignment loss, type 'mse': loss = mean((zs - feat)^2); dzs = gscale * 2 (zs - feat) / numel.
// two deterministic stages: per-CTA partial sums, then a single-CTA finish.
template <typename ZT, typename FT>
__global__ void __launch_bounds__(256)
align_mse_stage1(const ZT* __restrict__ zs, const FT* __restrict__ feat, ZT* __restrict__ dzs, float gcoef,
                 long long n, float* __restrict__ part) {
  float acc = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float d = (float)zs[i] - (float)feat[i];
    acc += d * d;
    if (dzs) dzs[i] = (ZT)(gcoef * d);
  }
  __shared__ float red[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    part[blockIdx.x] = s;
  }
}
__global__ void align_mse_stage2(const float* __restrict__ part, int nparts, float inv_n, float* __restrict__ loss) {
  __shared__ float red[256];
  float s = 0.f;
  for (int i = threadIdx.x; i < nparts; i += 256) s += part[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = red[0] * inv_n;
}

// backward of the fused alignment loss: dzs = g * 2 (zs - feat) / n, g a device scalar
template <typename ZT, typename FT>
__global__ void __launch_bounds__(256)
align_mse_bwd_kernel(const ZT* __restrict__ zs, const FT* __restrict__ feat, const float* __restrict__ g,
                     ZT* __restrict__ dzs, float two_over_n, long long n) {
  const float coef = __ldg(g) * two_over_n;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dzs[i] = (ZT)(coef * ((float)zs[i] - (float)feat[i]));
}

// Row-wise alignment losses (gaussian_diffusion.py:1008-1019), one warp per row of D values:
//   kind 0 'cosine': c = <t, o> / (max(|t|, 1e-8) max(|o|, 1e-8)); row value -c;          d/do = -(t / (|t||o|) - c o / |o|^2)
//   kind 1 'mse_l2': row value |o/|o| - t/|t||^2 (the mean runs over rows * D elements);  d/do = (2 / |o|)(oh <oh, th> - th)
template <typename ZT, typename FT>
__global__ void __launch_bounds__(256)
align_rowwise_kernel(const ZT* __restrict__ zs, const FT* __restrict__ feat, int kind, ZT* __restrict__ dzs,
                     float gcoef, long long rows, int D, float* __restrict__ part) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const ZT* o = zs + r * D;
  const FT* t = feat + r * D;
  float dot = 0.f, oo = 0.f, tt = 0.f;
  for (int i = lane; i < D; i += 32) {
    const float a = (float)o[i], b = (float)t[i];
    dot += a * b;
    oo += a * a;
    tt += b * b;
  }
  dot = warp_sum(dot);
  oo = warp_sum(oo);
  tt = warp_sum(tt);
  const float eps = kind == 0 ? 1e-8f : 1e-12f;
  const float no = fmaxf(sqrtf(oo), eps), nt = fmaxf(sqrtf(tt), eps);
  const float c = dot / (no * nt);
  float val, ko, kt;   // gradient row = ko * o + kt * t
  if (kind == 0) {
    val = -c;
    ko = c / (no * no);
    kt = -1.f / (no * nt);
  } else {
    // sum_i (o_i / |o| - t_i / |t|)^2, accumulated element by element like the reference (the closed form 2 - 2c
    // cancels catastrophically once the projector has learned to align)
    const float io = 1.f / no, it = 1.f / nt;
    float acc = 0.f;
    for (int i = lane; i < D; i += 32) {
      const float d = (float)o[i] * io - (float)t[i] * it;
      acc += d * d;
    }
    val = warp_sum(acc);
    ko = 2.f * c / (no * no);
    kt = -2.f / (no * nt);
  }
  if (lane == 0) part[r] = val;
  if (dzs) {
    ZT* d = dzs + r * D;
    for (int i = lane; i < D; i += 32) d[i] = (ZT)(gcoef * (ko * (float)o[i] + kt * (float)t[i]));
  }
}

inline unsigned grid_for(long long work, int per_block = 256) {
  long long b = (work + per_block - 1) / per_block;
  const long long cap = (long long)vaw_num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace

extern "C" int vaw_patchify_in(const float* x, void* patches, int B, int C, int H, int W, int P, cudaStream_t stream) {
  VAW_CHECK_ARG(x && patches && B > 0 && C > 0 && P > 0 && H % P == 0 && W % P == 0, "vaw_patchify_in: bad arguments");
  patchify_in_kernel<<<grid_for((long long)B * C * H * W), 256, 0, stream>>>(x, (bf16*)patches, B, C, H, W, P);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

extern "C" int vaw_patchify_norm(const float* x, const float* mean, const float* stdv, void* patches, int B, int C,
                                 int H, int W, int P, cudaStream_t stream) {
  VAW_CHECK_ARG(x && mean && stdv && patches && B > 0 && C > 0 && P > 0 && H % P == 0 && W % P == 0,
                "vaw_patchify_norm: bad arguments");
  patchify_norm_kernel<<<grid_for((long long)B * C * H * W), 256, 0, stream>>>(x, mean, stdv, (bf16*)patches, B, C, H,
                                                                               W, P);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

extern "C" int vaw_vit_assemble(const float* tok, const float* cls, const float* pos0, float* x, int B, int L, int D,
                                cudaStream_t stream) {
  VAW_CHECK_ARG(tok && cls && pos0 && x && B > 0 && L > 0 && D > 0, "vaw_vit_assemble: bad arguments");
  vit_assemble_kernel<<<grid_for((long long)B * (L + 1) * D), 256, 0, stream>>>(tok, cls, pos0, x, B, L, D);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

// to_image = 1: tokens -> image; 0: image -> tokens.  dtype: 0 fp32, 1 bf16
extern "C" int vaw_unpatchify(void* tokens, void* image, int dtype, int B, int C, int H, int W, int P, int to_image,
                              cudaStream_t stream) {
  VAW_CHECK_ARG(tokens && image && B > 0 && C > 0 && P > 0 && H % P == 0 && W % P == 0, "vaw_unpatchify: bad arguments");
  const unsigned g = grid_for((long long)B * C * H * W);
  if (dtype == 0) unpatchify_kernel<float><<<g, 256, 0, stream>>>((float*)tokens, (float*)image, B, C, H, W, P, to_image);
  else if (dtype == 1) unpatchify_kernel<bf16><<<g, 256, 0, stream>>>((bf16*)tokens, (bf16*)image, B, C, H, W, P, to_image);
  else { vaw_set_error("vaw_unpatchify: bad dtype %d", dtype); return VAW_ERR_INVALID; }
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

extern "C" int vaw_timestep_embedding(const float* t, void* out_bf16, float* out_f32, int B, int dim,
                                      cudaStream_t stream) {
  VAW_CHECK_ARG(t && (out_bf16 || out_f32) && B > 0 && dim > 0, "vaw_timestep_embedding: bad arguments");
  timestep_embedding_kernel<<<(B * dim + 255) / 256, 256, 0, stream>>>(t, (bf16*)out_bf16, out_f32, B, dim);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

extern "C" int vaw_cond_combine(const float* t_emb, const float* table, const long long* labels, float* c,
                                void* c_silu, int B, int D, cudaStream_t stream) {
  VAW_CHECK_ARG(t_emb && c && c_silu && B > 0 && D > 0 && (!table || labels), "vaw_cond_combine: bad arguments");
  cond_combine_kernel<<<(B * D + 255) / 256, 256, 0, stream>>>(t_emb, table, labels, c, (bf16*)c_silu, B, D);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

extern "C" int vaw_cond_bwd(const float* dc_silu, const float* c, float* dc, void* dc_bf16, int n, cudaStream_t stream) {
  VAW_CHECK_ARG(dc_silu && c && dc && dc_bf16 && n > 0, "vaw_cond_bwd: bad arguments");
  cond_bwd_kernel<<<(n + 255) / 256, 256, 0, stream>>>(dc_silu, c, dc, (bf16*)dc_bf16, n);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

extern "C" int vaw_embedding_grad(const float* dc, const long long* labels, float* dtable, int rows, int B, int D,
                                  int accumulate, cudaStream_t stream) {
  VAW_CHECK_ARG(dc && labels && dtable && rows > 0 && B > 0 && D > 0, "vaw_embedding_grad: bad arguments");
  embedding_grad_kernel<<<rows, 128, 0, stream>>>(dc, (long long)D, labels, dtable, B, D, accumulate);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

// same with a row stride ld for dc (U-ViT: the label token's gradient rows sit T*D apart)
extern "C" int vaw_embedding_grad_strided(const float* dc, long long ld, const long long* labels, float* dtable, int rows,
                                          int B, int D, int accumulate, cudaStream_t stream) {
  VAW_CHECK_ARG(dc && labels && dtable && rows > 0 && B > 0 && D > 0 && ld >= D, "vaw_embedding_grad_strided: bad arguments");
  embedding_grad_kernel<<<rows, 128, 0, stream>>>(dc, ld, labels, dtable, B, D, accumulate);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

extern "C" int vaw_cast_f32_bf16(const float* src, void* dst, long long n, cudaStream_t stream) {
  VAW_CHECK_ARG(src && dst && n >= 0, "vaw_cast_f32_bf16: bad arguments");
  VAW_CHECK_ARG((((uintptr_t)src & 15) | ((uintptr_t)dst & 7)) == 0, "vaw_cast_f32_bf16: misaligned buffers");
  if (n == 0) return VAW_OK;
  cast_f32_bf16_kernel<<<grid_for(n / 4 + 1), 256, 0, stream>>>(src, (bf16*)dst, n);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

extern "C" int vaw_cast_f32_bf16_2d(const float* src, long long lds, void* dst, long long ldd, int rows, int cols,
                                    cudaStream_t stream) {
  VAW_CHECK_ARG(src && dst && rows > 0 && cols > 0 && cols % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0,
                "vaw_cast_f32_bf16_2d: bad arguments (cols, lds, ldd multiples of 4)");
  cast_f32_bf16_2d_kernel<<<grid_for((long long)rows * cols / 4), 256, 0, stream>>>(src, lds, (bf16*)dst, ldd, rows, cols);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

// dst += src over a flat fp32 buffer (gradient accumulation around a CUDA-graph replay, vaw_b200/graph.py)
__global__ void __launch_bounds__(256) add_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, long long n) {
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 a = ldg_stream_f4(reinterpret_cast<const float4*>(src) + i);
    float4 d = reinterpret_cast<float4*>(dst)[i];
    d.x += a.x; d.y += a.y; d.z += a.z; d.w += a.w;
    reinterpret_cast<float4*>(dst)[i] = d;
  }
}

extern "C" int vaw_add_f32(const float* src, float* dst, long long n, cudaStream_t stream) {
  VAW_CHECK_ARG(src && dst && n >= 0 && n % 4 == 0, "vaw_add_f32: n must be a multiple of 4");
  if (n == 0) return VAW_OK;
  add_f32_kernel<<<grid_for(n / 4), 256, 0, stream>>>(src, dst, n);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

extern "C" int vaw_add_bf16_into_f32(const void* src, float* dst, long long n, cudaStream_t stream) {
  VAW_CHECK_ARG(src && dst && n >= 0 && n % 4 == 0, "vaw_add_bf16_into_f32: n must be a multiple of 4");
  if (n == 0) return VAW_OK;
  add_bf16_into_f32_kernel<<<grid_for(n / 4), 256, 0, stream>>>((const bf16*)src, dst, n);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

extern "C" int vaw_colsum_f32_small(const float* a, long long lda, int rows, int N, float* out, int accumulate,
                                    cudaStream_t stream) {
  VAW_CHECK_ARG(a && out && rows > 0 && N > 0, "vaw_colsum_f32_small: bad arguments");
  colsum_f32_small_kernel<<<(N + 255) / 256, 256, 0, stream>>>(a, lda, rows, N, out, accumulate);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

// step: 1-based optimizer step (bias corrections are computed on the host in double like torch does)
extern "C" int vaw_adamw_step_amp(float* p, const float* g, float* m, float* v, void* p_bf16, float* ema, long long n,
                                  double lr, double beta1, double beta2, double eps, double weight_decay,
                                  long long step, double grad_scale, double ema_decay, const float* clip_coef,
                                  const float* inv_scale, const float* found_inf, cudaStream_t stream) {
  VAW_CHECK_ARG(p && g && m && v && n >= 0 && n % 4 == 0 && step >= 1, "vaw_adamw_step: bad arguments (n %% 4 == 0)");
  if (n == 0) return VAW_OK;
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  adamw_kernel<<<grid_for(n / 4), 256, 0, stream>>>(p, g, m, v, (bf16*)p_bf16, ema, n, (float)lr, (float)beta1,
                                                    (float)beta2, (float)eps, (float)weight_decay, (float)bc1,
                                                    (float)sqrt(bc2), (float)grad_scale, (float)ema_decay, clip_coef,
                                                    inv_scale, found_inf);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

// ranges: DEVICE array of n_ranges (offset, count) pairs in elements; offsets and counts multiples of 4
extern "C" int vaw_adamw_step_ranges(float* p, const float* g, float* m, float* v, void* p_bf16, float* ema,
                                     const long long* ranges, int n_ranges, long long max_count, double lr, double beta1,
                                     double beta2, double eps, double weight_decay, long long step, double grad_scale,
                                     double ema_decay, const float* clip_coef, const float* inv_scale,
                                     const float* found_inf, cudaStream_t stream) {
  VAW_CHECK_ARG(p && g && m && v && ranges && n_ranges >= 0 && n_ranges <= 65535 && step >= 1 && max_count >= 0,
                "vaw_adamw_step_ranges: bad arguments");
  if (n_ranges == 0 || max_count == 0) return VAW_OK;
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  unsigned gx = grid_for(max_count / 4);
  if (gx > 256) gx = 256;
  adamw_ranges_kernel<<<dim3(gx, (unsigned)n_ranges), 256, 0, stream>>>(
      p, g, m, v, (bf16*)p_bf16, ema, ranges, (float)lr, (float)beta1, (float)beta2, (float)eps, (float)weight_decay,
      (float)bc1, (float)sqrt(bc2), (float)grad_scale, (float)ema_decay, clip_coef, inv_scale, found_inf);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

extern "C" int vaw_adamw_step(float* p, const float* g, float* m, float* v, void* p_bf16, float* ema, long long n,
                              double lr, double beta1, double beta2, double eps, double weight_decay, long long step,
                              double grad_scale, double ema_decay, const float* clip_coef, cudaStream_t stream) {
  return vaw_adamw_step_amp(p, g, m, v, p_bf16, ema, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale,
                            ema_decay, clip_coef, nullptr, nullptr, stream);
}

// out[0] = ||grad_scale * g||_2, out[1] = clip coefficient min(1, max_norm / (out[0] + 1e-6)) (1 if max_norm <= 0);
// part: scratch of 1024 floats.
extern "C" int vaw_grad_clip_coef(const float* g, long long n, double grad_scale, double max_norm, float* part,
                                  float* out, cudaStream_t stream) {
  VAW_CHECK_ARG(g && part && out && n >= 0, "vaw_grad_clip_coef: bad arguments");
  VAW_CHECK_ARG((reinterpret_cast<uintptr_t>(g) & 15) == 0, "vaw_grad_clip_coef: g must be 16-byte aligned");
  grad_sqnorm_stage1<<<kNormBlocks, 256, 0, stream>>>(g, n, (float)grad_scale, part);
  VAW_LAUNCH_CHECK();
  grad_sqnorm_stage2<<<1, 256, 0, stream>>>(part, (float)max_norm, out);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

extern "C" int vaw_align_mse_finish(const float* part, long long nparts, long long n, float* loss, cudaStream_t stream) {
  VAW_CHECK_ARG(part && loss && nparts > 0 && nparts < (1LL << 31) && n > 0, "vaw_align_mse_finish: bad arguments");
  align_mse_stage2<<<1, 256, 0, stream>>>(part, (int)nparts, 1.0f / (float)n, loss);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

#define TRY_RC(expr) do { int _rc = (expr); if (_rc != VAW_OK) return _rc; } while (0)

namespace {
// dtype dispatch of the (zs, feat) pairs: F is called with two typed null pointers that only carry the element types
template <typename F>
int dispatch_zf(int zs_dtype, int feat_dtype, F&& f) {
  if (zs_dtype == 1 && feat_dtype == 1) f((bf16*)nullptr, (bf16*)nullptr);
  else if (zs_dtype == 1 && feat_dtype == 0) f((bf16*)nullptr, (float*)nullptr);
  else if (zs_dtype == 0 && feat_dtype == 0) f((float*)nullptr, (float*)nullptr);
  else if (zs_dtype == 0 && feat_dtype == 1) f((float*)nullptr, (bf16*)nullptr);
  else { vaw_set_error("bad dtypes %d %d", zs_dtype, feat_dtype); return VAW_ERR_INVALID; }
  return VAW_OK;
}
}  // namespace

extern "C" int vaw_align_mse_bwd(const void* zs, int zs_dtype, const void* feat, int feat_dtype, const float* g,
                                 void* dzs, long long n, cudaStream_t stream) {
  VAW_CHECK_ARG(zs && feat && g && dzs && n > 0, "vaw_align_mse_bwd: bad arguments");
  const unsigned blocks = grid_for(n);
  const float k = 2.f / (float)n;
  TRY_RC(dispatch_zf(zs_dtype, feat_dtype, [&](auto* z, auto* f) {
    using ZT = std::remove_pointer_t<decltype(z)>;
    using FT = std::remove_pointer_t<decltype(f)>;
    align_mse_bwd_kernel<ZT, FT><<<blocks, 256, 0, stream>>>((const ZT*)zs, (const FT*)feat, g, (ZT*)dzs, k, n);
  }));
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

extern "C" int vaw_align_rowwise(const void* zs, int zs_dtype, const void* feat, int feat_dtype, int kind, void* dzs,
                                 float gscale, long long rows, int D, float* part, float* loss, cudaStream_t stream) {
  VAW_CHECK_ARG(zs && feat && part && loss && rows > 0 && rows < (1LL << 31) && D > 0 && (kind == 0 || kind == 1),
                "vaw_align_rowwise: bad arguments");
  const unsigned blocks = (unsigned)((rows + 7) / 8);
  // 'cosine' averages over rows, 'mse_l2' over rows * D elements
  const float inv = kind == 0 ? 1.0f / (float)rows : 1.0f / ((float)rows * (float)D);
  const float gcoef = gscale * inv;
  TRY_RC(dispatch_zf(zs_dtype, feat_dtype, [&](auto* z, auto* f) {
    using ZT = std::remove_pointer_t<decltype(z)>;
    using FT = std::remove_pointer_t<decltype(f)>;
    align_rowwise_kernel<ZT, FT><<<blocks, 256, 0, stream>>>((const ZT*)zs, (const FT*)feat, kind, (ZT*)dzs, gcoef, rows,
                                                             D, part);
  }));
  VAW_LAUNCH_CHECK();
  align_mse_stage2<<<1, 256, 0, stream>>>(part, (int)rows, inv, loss);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

// zs/feat dtype codes: 0 fp32, 1 bf16.  part: scratch of >= 1024 floats.  gscale multiplies the gradient.
extern "C" int vaw_align_mse(const void* zs, int zs_dtype, const void* feat, int feat_dtype, void* dzs, float gscale,
                             long long n, float* part, float* loss, cudaStream_t stream) {
  VAW_CHECK_ARG(zs && feat && part && loss && n > 0, "vaw_align_mse: bad arguments");
  unsigned blocks = grid_for(n);
  if (blocks > 1024) blocks = 1024;
  const float gcoef = gscale * 2.f / (float)n;
  if (zs_dtype == 1 && feat_dtype == 1)
    align_mse_stage1<bf16, bf16><<<blocks, 256, 0, stream>>>((const bf16*)zs, (const bf16*)feat, (bf16*)dzs, gcoef, n, part);
  else if (zs_dtype == 1 && feat_dtype == 0)
    align_mse_stage1<bf16, float><<<blocks, 256, 0, stream>>>((const bf16*)zs, (const float*)feat, (bf16*)dzs, gcoef, n, part);
  else if (zs_dtype == 0 && feat_dtype == 0)
    align_mse_stage1<float, float><<<blocks, 256, 0, stream>>>((const float*)zs, (const float*)feat, (float*)dzs, gcoef, n, part);
  else if (zs_dtype == 0 && feat_dtype == 1)
    align_mse_stage1<float, bf16><<<blocks, 256, 0, stream>>>((const float*)zs, (const bf16*)feat, (float*)dzs, gcoef, n, part);
  else { vaw_set_error("vaw_align_mse: bad dtypes %d %d", zs_dtype, feat_dtype); return VAW_ERR_INVALID; }
  VAW_LAUNCH_CHECK();
  align_mse_stage2<<<1, 256, 0, stream>>>(part, (int)blocks, 1.0f / (float)n, loss);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

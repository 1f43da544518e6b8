// uvit_kernels.cu — the U-ViT-specific glue kernels (models/uvit.py): token assembly (label / time tokens + patches +
// pos_embed) and its backward, the [x, skip] concat for skip_linear and its split backward, strided (un)patchify past
// the extra tokens, and the final 3x3 convolution (forward, input gradient, weight gradient).
// All are small and memory-bound; the transformer blocks themselves run on the shared GEMM / attention / LN kernels.
#include "vaw_common.cuh"

namespace {

// x0[b, tok, :] = token(b, tok) + pos[tok, :]   (uvit.py:221-231)
//   tok <  extras: label embedding (tok 0 when class-conditional) / raw sinusoidal time embedding (last extra token)
//   tok >= extras: bf16-rounded patch embedding (the Conv2d output is a low-precision tensor under autocast)
__global__ void __launch_bounds__(256)
uvit_assemble_kernel(const float* __restrict__ patch_tok, const float* __restrict__ t, const float* __restrict__ table,
                     const long long* __restrict__ labels, const float* __restrict__ pos, float* __restrict__ x0, int B,
                     int T, int extras, int D) {
  const long long total = (long long)B * T * D;
  const int half = D / 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % D), tok = (int)((i / D) % T), b = (int)(i / ((long long)D * T));
    float v;
    if (tok >= extras) {
      v = __bfloat162float(__float2bfloat16_rn(patch_tok[((long long)b * (T - extras) + (tok - extras)) * D + j]));
    } else if (tok == extras - 1) {  // time token: cat(cos, sin) of t * exp(-ln(1e4) k / half)   (uvit.py:21-39)
      v = 0.f;
      if (j < 2 * half) {
        const int k = j < half ? j : j - half;
        const float a = t[b] * expf(-9.210340371976184f * (float)k / (float)half);
        v = j < half ? cosf(a) : sinf(a);
      }
    } else {
      v = table[labels[b] * (long long)D + j];
    }
    x0[i] = v + pos[(long long)tok * D + j];
  }
}

// dpos[tok, j] (+)= sum_b dx0[b, tok, j]  (fixed order over b)
__global__ void __launch_bounds__(256)
uvit_pos_grad_kernel(const float* __restrict__ dx0, float* __restrict__ dpos, int B, int T, int D, int accumulate) {
  const long long n = (long long)T * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += dx0[(long long)b * n + i];
    dpos[i] = accumulate ? dpos[i] + s : s;
  }
}

// rows [b, extras + l] of dx0 -> dtok bf16 [B*L, D] (operand of the patch-embed wgrad)
__global__ void __launch_bounds__(256)
uvit_gather_patch_grad_kernel(const float* __restrict__ dx0, bf16* __restrict__ dtok, int B, int T, int extras, int D) {
  const int Lp = T - extras;
  const long long total = (long long)B * Lp * (D / 4);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % (D / 4));
    const long long row = i / (D / 4);
    const int l = (int)(row % Lp), b = (int)(row / Lp);
    const float4 v = *reinterpret_cast<const float4*>(dx0 + ((long long)b * T + extras + l) * D + c4 * 4);
    uint2 o;
    o.x = pack_bf16(v.x, v.y);
    o.y = pack_bf16(v.z, v.w);
    *reinterpret_cast<uint2*>(dtok + row * D + c4 * 4) = o;
  }
}

// cat[row, :] = bf16([x[row, :], skip[row, :]])   (uvit.py:117-118, operand of skip_linear)
__global__ void __launch_bounds__(256)
cat_cast_kernel(const float* __restrict__ x, const float* __restrict__ skip, bf16* __restrict__ cat, long long M, int D) {
  const long long total = M * (D / 4) * 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % (D / 2));  // float4 index inside the 2D-wide row
    const long long row = i / (D / 2);
    const bool second = c4 >= D / 4;
    const float4 v = *reinterpret_cast<const float4*>((second ? skip : x) + row * D + (second ? c4 - D / 4 : c4) * 4);
    uint2 o;
    o.x = pack_bf16(v.x, v.y);
    o.y = pack_bf16(v.z, v.w);
    *reinterpret_cast<uint2*>(cat + row * 2 * D + c4 * 4) = o;
  }
}

// dst[row, :] (+)= f32(src[row, col_off : col_off + D])  with src bf16 of leading dimension ld
__global__ void __launch_bounds__(256)
unpack_cols_kernel(const bf16* __restrict__ src, long long ld, int col_off, float* __restrict__ dst, long long M, int D,
                   int accumulate) {
  const long long total = M * (D / 4);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % (D / 4));
    const long long row = i / (D / 4);
    const uint2 u = *reinterpret_cast<const uint2*>(src + row * ld + col_off + c4 * 4);
    const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
    float4* d = reinterpret_cast<float4*>(dst + row * D + c4 * 4);
    float4 o = make_float4(a.x, a.y, b.x, b.y);
    if (accumulate) {
      const float4 p = *d;
      o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
    }
    *d = o;
  }
}

// tokens [B, rows_per_sample, F] (feature order (p, q, c), patch rows start at row0) <-> image [B, C, H, W]
template <typename TT, typename TI>
__global__ void __launch_bounds__(256)
unpatchify_strided_kernel(TT* tok, TI* img, int B, int C, int H, int W, int P, int to_image, int row0,
                          int rows_per_sample, int zero_extras) {
  const int Wg = W / P, Hg = H / P, F = C * P * P;
  const long long total = (long long)B * C * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xw = (int)(i % W), yh = (int)((i / W) % H), c = (int)((i / ((long long)W * H)) % C);
    const int n = (int)(i / ((long long)W * H * C));
    const int h = yh / P, p = yh % P, w = xw / P, q = xw % P;
    const long long ti = ((long long)n * rows_per_sample + row0 + (long long)h * Wg + w) * F + (p * P + q) * C + c;
    if (to_image) img[i] = (TI)(float)tok[ti];
    else tok[ti] = (TT)(float)img[i];
  }
  if (!to_image && zero_extras) {  // gradient rows of the extra tokens are zero
    const long long ez = (long long)B * row0 * F;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < ez; i += (long long)gridDim.x * blockDim.x) {
      const int f = (int)(i % F);
      const int r = (int)((i / F) % row0), n = (int)(i / ((long long)F * row0));
      tok[((long long)n * rows_per_sample + r) * F + f] = (TT)0.f;
    }
  }
}

// 3x3 convolution, padding 1, C channels in and out (C <= 8), fp32  (uvit.py:192, final_layer)
//   transpose = 0: out[b,co,y,x] = bias[co] + sum_{ci,dy,dx} w[co,ci,dy,dx] * in[b,ci,y+dy-1,x+dx-1]
//   transpose = 1: input gradient: out[b,ci,y,x] = sum_{co,dy,dx} w[co,ci,dy,dx] * in[b,co,y-dy+1,x-dx+1]
__global__ void __launch_bounds__(256)
conv3x3_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
               float* __restrict__ out, int B, int C, int H, int W, int transpose) {
  const long long total = (long long)B * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H), b = (int)(i / ((long long)W * H));
    float acc[8];
    for (int c = 0; c < C; ++c) acc[c] = (bias && !transpose) ? bias[c] : 0.f;
    for (int ci = 0; ci < C; ++ci)
      for (int dy = 0; dy < 3; ++dy)
        for (int dx = 0; dx < 3; ++dx) {
          const int yy = transpose ? y - dy + 1 : y + dy - 1, xx = transpose ? x - dx + 1 : x + dx - 1;
          if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
          const float v = in[(((long long)b * C + ci) * H + yy) * W + xx];
          for (int co = 0; co < C; ++co)
            acc[co] += v * (transpose ? w[((ci * C + co) * 3 + dy) * 3 + dx] : w[((co * C + ci) * 3 + dy) * 3 + dx]);
        }
    for (int c = 0; c < C; ++c) out[(((long long)b * C + c) * H + y) * W + x] = acc[c];
  }
}

// weight / bias gradient of the 3x3 convolution: one CTA per (co, ci, dy, dx) tap (+ C CTAs for the bias),
// fixed-order block reduction over all (b, y, x)
// 1024 threads per tap and four positions in flight per thread: the kernel is a latency-bound reduction over B*H*W
// positions served from L2 (84 CTAs, 6 MB of inputs); with 256 threads and one dependent load pair per iteration it took
// 880 us of a 25 ms U-ViT step.
constexpr int kConvThreads = 1024;
__global__ void __launch_bounds__(kConvThreads)
conv3x3_wgrad_kernel(const float* __restrict__ in, const float* __restrict__ dout, float* __restrict__ dw,
                     float* __restrict__ dbias, int B, int C, int H, int W, int accumulate) {
  const int tap = blockIdx.x;
  const int ntaps = C * C * 9;
  const bool is_bias = tap >= ntaps;
  const int co = is_bias ? tap - ntaps : tap / (C * 9);
  const int ci = is_bias ? 0 : (tap / 9) % C, dy = is_bias ? 0 : (tap / 3) % 3, dx = is_bias ? 0 : tap % 3;
  float s = 0.f;
  const long long total = (long long)B * H * W;
  const int hw = H * W;
#pragma unroll 4
  for (long long i = threadIdx.x; i < total; i += kConvThreads) {
    const int b = (int)(i / hw), r = (int)(i - (long long)b * hw);
    const int y = r / W, x = r - y * W;
    const float g = __ldg(dout + ((long long)b * C + co) * hw + r);
    if (is_bias) {
      s += g;
    } else {
      const int yy = y + dy - 1, xx = x + dx - 1;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) s += g * __ldg(in + ((long long)b * C + ci) * hw + yy * W + xx);
    }
  }
  __shared__ float red[kConvThreads];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = kConvThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float* dst = is_bias ? dbias + co : dw + tap;
    *dst = accumulate ? *dst + red[0] : red[0];
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

extern "C" int vaw_uvit_assemble(const float* patch_tok, const float* t, const float* table, const long long* labels,
                                 const float* pos, float* x0, int B, int T, int extras, int D, cudaStream_t stream) {
  VAW_CHECK_ARG(patch_tok && t && pos && x0 && B > 0 && T > extras && extras >= 1 && extras <= 2 && D > 0,
                "vaw_uvit_assemble: bad arguments");
  VAW_CHECK_ARG(extras == 1 || (table && labels), "vaw_uvit_assemble: labels required when class-conditional");
  uvit_assemble_kernel<<<grid_for((long long)B * T * D), 256, 0, stream>>>(patch_tok, t, table, labels, pos, x0, B, T,
                                                                          extras, D);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

extern "C" int vaw_uvit_pos_grad(const float* dx0, float* dpos, int B, int T, int D, int accumulate,
                                 cudaStream_t stream) {
  VAW_CHECK_ARG(dx0 && dpos && B > 0 && T > 0 && D > 0, "vaw_uvit_pos_grad: bad arguments");
  uvit_pos_grad_kernel<<<grid_for((long long)T * D), 256, 0, stream>>>(dx0, dpos, B, T, D, accumulate);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

extern "C" int vaw_uvit_gather_patch_grad(const float* dx0, void* dtok, int B, int T, int extras, int D,
                                          cudaStream_t stream) {
  VAW_CHECK_ARG(dx0 && dtok && B > 0 && T > extras && D % 4 == 0, "vaw_uvit_gather_patch_grad: bad arguments");
  uvit_gather_patch_grad_kernel<<<grid_for((long long)B * (T - extras) * D / 4), 256, 0, stream>>>(dx0, (bf16*)dtok, B,
                                                                                                 T, extras, D);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

extern "C" int vaw_cat_cast(const float* x, const float* skip, void* cat, long long M, int D, cudaStream_t stream) {
  VAW_CHECK_ARG(x && skip && cat && M > 0 && D % 4 == 0, "vaw_cat_cast: bad arguments");
  cat_cast_kernel<<<grid_for(M * D / 2), 256, 0, stream>>>(x, skip, (bf16*)cat, M, D);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

extern "C" int vaw_unpack_cols(const void* src, long long ld, int col_off, float* dst, long long M, int D,
                               int accumulate, cudaStream_t stream) {
  VAW_CHECK_ARG(src && dst && M > 0 && D % 4 == 0 && ld % 4 == 0 && col_off % 4 == 0, "vaw_unpack_cols: bad arguments");
  unpack_cols_kernel<<<grid_for(M * D / 4), 256, 0, stream>>>((const bf16*)src, ld, col_off, dst, M, D, accumulate);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

// tok_dtype / img_dtype: 0 fp32, 1 bf16.  to_image = 0 also zeroes the rows of the `row0` extra tokens when zero_extras.
extern "C" int vaw_unpatchify_strided(void* tokens, int tok_dtype, void* image, int img_dtype, int B, int C, int H,
                                      int W, int P, int to_image, int row0, int rows_per_sample, int zero_extras,
                                      cudaStream_t stream) {
  VAW_CHECK_ARG(tokens && image && B > 0 && C > 0 && P > 0 && H % P == 0 && W % P == 0 && row0 >= 0 &&
                    rows_per_sample >= row0 + (H / P) * (W / P),
                "vaw_unpatchify_strided: bad arguments");
  const unsigned g = grid_for((long long)B * C * H * W);
  if (tok_dtype == 0 && img_dtype == 0)
    unpatchify_strided_kernel<float, float><<<g, 256, 0, stream>>>((float*)tokens, (float*)image, B, C, H, W, P, to_image, row0, rows_per_sample, zero_extras);
  else if (tok_dtype == 1 && img_dtype == 0)
    unpatchify_strided_kernel<bf16, float><<<g, 256, 0, stream>>>((bf16*)tokens, (float*)image, B, C, H, W, P, to_image, row0, rows_per_sample, zero_extras);
  else if (tok_dtype == 1 && img_dtype == 1)
    unpatchify_strided_kernel<bf16, bf16><<<g, 256, 0, stream>>>((bf16*)tokens, (bf16*)image, B, C, H, W, P, to_image, row0, rows_per_sample, zero_extras);
  else {
    vaw_set_error("vaw_unpatchify_strided: unsupported dtype combination %d/%d", tok_dtype, img_dtype);
    return VAW_ERR_INVALID;
  }
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

extern "C" int vaw_conv3x3(const float* in, const float* w, const float* bias, float* out, int B, int C, int H, int W,
                           int transpose, cudaStream_t stream) {
  VAW_CHECK_ARG(in && w && out && B > 0 && C > 0 && C <= 8 && H > 0 && W > 0, "vaw_conv3x3: bad arguments (C <= 8)");
  conv3x3_kernel<<<grid_for((long long)B * H * W), 256, 0, stream>>>(in, w, bias, out, B, C, H, W, transpose);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

extern "C" int vaw_conv3x3_wgrad(const float* in, const float* dout, float* dw, float* dbias, int B, int C, int H,
                                 int W, int accumulate, cudaStream_t stream) {
  VAW_CHECK_ARG(in && dout && dw && dbias && B > 0 && C > 0 && C <= 8, "vaw_conv3x3_wgrad: bad arguments");
  conv3x3_wgrad_kernel<<<C * C * 9 + C, kConvThreads, 0, stream>>>(in, dout, dw, dbias, B, C, H, W, accumulate);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

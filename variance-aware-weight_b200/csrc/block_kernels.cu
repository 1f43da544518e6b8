// block_kernels.cu — the memory-bound pieces of the DiT / U-ViT block that cannot live in a GEMM epilogue:
//   LayerNorm (+ adaLN modulate, or affine) forward / backward          (models/dit.py:24-25,122-124,133-137;
//   gate * branch backward (adaLN-Zero) and residual bookkeeping          models/uvit.py:96-121)
//   deterministic column reductions for bias / shift / scale / gate gradients
// All kernels are HBM-bound: one pass over the [rows, D] activation, 128-bit accesses, fp32 math, and two-level
// (per-CTA partial -> fixed-order finish) reductions so gradients are bit-reproducible run to run.
#include "vaw_common.cuh"

namespace {

constexpr int kMaxD = 2048;  // rows live in registers: KV = ceil(D / 128) float4 per lane, KV in {3, 6, 9, 16}

// ---------------------------------------------------------------------------------------------------
// LayerNorm forward: y = xhat * A + Bv, xhat = (x - mean) * rstd
//   modulate (DiT):  A = 1 + scale[n, :], Bv = shift[n, :]     (n = row / rows_per_sample; ld_mod = row stride)
//   affine (U-ViT):  A = weight[:],       Bv = bias[:]
//   plain:           A = 1, Bv = 0
// one warp per row, the row lives in registers (two-pass variance like torch's layer_norm).
// ---------------------------------------------------------------------------------------------------
template <int KV>
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ shift, const float* __restrict__ scale,
              long long ld_mod, int rows_per_sample, const float* __restrict__ weight,
              const float* __restrict__ bias, bf16* __restrict__ y, float* __restrict__ mean_out,
              float* __restrict__ rstd_out, int M, int D, float eps) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (long long)row * D);
  const int nv = D >> 2;  // float4 per row
  float4 v[KV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < KV; ++i) {
    const int idx = i * 32 + lane;
    if (idx < nv) {
      v[i] = xr[idx];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float mean = warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < KV; ++i) {
    const int idx = i * 32 + lane;
    if (idx < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
  if (lane == 0) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
  const long long mo = scale ? (long long)(row / rows_per_sample) * ld_mod : 0;
  uint2* yr = reinterpret_cast<uint2*>(y + (long long)row * D);
#pragma unroll
  for (int i = 0; i < KV; ++i) {
    const int idx = i * 32 + lane;
    if (idx < nv) {
      float4 A = make_float4(1.f, 1.f, 1.f, 1.f), Bv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (scale) {
        const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + mo) + idx);
        const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + mo) + idx);
        A = make_float4(1.f + sc.x, 1.f + sc.y, 1.f + sc.z, 1.f + sc.w);
        Bv = sh;
      } else if (weight) {
        A = __ldg(reinterpret_cast<const float4*>(weight) + idx);
        Bv = __ldg(reinterpret_cast<const float4*>(bias) + idx);
      }
      const float a = (v[i].x - mean) * rstd, b = (v[i].y - mean) * rstd, c = (v[i].z - mean) * rstd,
                  d = (v[i].w - mean) * rstd;
      uint2 o;
      o.x = pack_bf16(fmaf(a, A.x, Bv.x), fmaf(b, A.y, Bv.y));
      o.y = pack_bf16(fmaf(c, A.z, Bv.z), fmaf(d, A.w, Bv.w));
      yr[idx] = o;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// LayerNorm backward.  grid = (chunks, groups): group = sample (modulate) or arbitrary row range (affine);
// every warp walks rows of its chunk, keeps per-column partial sums of dB = sum dy and dA = sum dy * xhat in
// its own shared-memory slice, the CTA then folds its 8 slices in fixed order into part[group, chunk, {dB,dA}, D].
//   g = dy * A ;  dx = rstd * (g - mean(g) - xhat * mean(g * xhat)) ;  dx_io = (add_into ? dx_io : 0) + dx
// ---------------------------------------------------------------------------------------------------
template <int KV>
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const bf16* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean_in,
              const float* __restrict__ rstd_in, const float* __restrict__ scale, long long ld_mod,
              const float* __restrict__ weight, float* __restrict__ dx_io, int add_into, float* __restrict__ part,
              int rows_per_group, int chunks, int M, int D) {
  extern __shared__ float sm_acc[];  // [8 warps][2][D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.x, group = blockIdx.y;
  const int nv = D >> 2;
  float* acc = sm_acc + (long long)warp * 2 * D;
  for (int i = lane; i < 2 * D; i += 32) acc[i] = 0.f;
  __syncwarp();
  const int rows_per_chunk = (rows_per_group + chunks - 1) / chunks;
  const int r_begin = group * rows_per_group + chunk * rows_per_chunk;
  const int r_end = min(min(r_begin + rows_per_chunk, (group + 1) * rows_per_group), M);
  const long long mo = scale ? (long long)group * ld_mod : 0;
  for (int row = r_begin + warp; row < r_end; row += 8) {
    const float4* xr = reinterpret_cast<const float4*>(x + (long long)row * D);
    const uint2* dyr = reinterpret_cast<const uint2*>(dy + (long long)row * D);
    const float mean = mean_in[row], rstd = rstd_in[row];
    float4 xh[KV], gv[KV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < KV; ++i) {
      const int idx = i * 32 + lane;
      if (idx < nv) {
        const float4 xv = xr[idx];
        const uint2 du = dyr[idx];
        const float2 d0 = unpack_bf16(du.x), d1 = unpack_bf16(du.y);
        xh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
        float4 A = make_float4(1.f, 1.f, 1.f, 1.f);
        if (scale) {
          const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + mo) + idx);
          A = make_float4(1.f + sc.x, 1.f + sc.y, 1.f + sc.z, 1.f + sc.w);
        } else if (weight) {
          A = __ldg(reinterpret_cast<const float4*>(weight) + idx);
        }
        gv[i] = make_float4(d0.x * A.x, d0.y * A.y, d1.x * A.z, d1.y * A.w);
        s1 += (gv[i].x + gv[i].y) + (gv[i].z + gv[i].w);
        s2 += (gv[i].x * xh[i].x + gv[i].y * xh[i].y) + (gv[i].z * xh[i].z + gv[i].w * xh[i].w);
        // column partials: dB += dy, dA += dy * xhat (each lane owns its columns -> no conflicts)
        float4* aB = reinterpret_cast<float4*>(acc) + idx;
        float4* aA = reinterpret_cast<float4*>(acc + D) + idx;
        float4 b = *aB, a = *aA;
        b.x += d0.x; b.y += d0.y; b.z += d1.x; b.w += d1.y;
        a.x += d0.x * xh[i].x; a.y += d0.y * xh[i].y; a.z += d1.x * xh[i].z; a.w += d1.y * xh[i].w;
        *aB = b;
        *aA = a;
      }
    }
    const float m1 = warp_sum(s1) / (float)D, m2 = warp_sum(s2) / (float)D;
    float4* dxr = reinterpret_cast<float4*>(dx_io + (long long)row * D);
#pragma unroll
    for (int i = 0; i < KV; ++i) {
      const int idx = i * 32 + lane;
      if (idx < nv) {
        float4 o;
        o.x = rstd * (gv[i].x - m1 - xh[i].x * m2);
        o.y = rstd * (gv[i].y - m1 - xh[i].y * m2);
        o.z = rstd * (gv[i].z - m1 - xh[i].z * m2);
        o.w = rstd * (gv[i].w - m1 - xh[i].w * m2);
        if (add_into) {
          const float4 p = dxr[idx];
          o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
        }
        dxr[idx] = o;
      }
    }
  }
  __syncthreads();
  if (part) {
    float* dst = part + ((long long)group * chunks + chunk) * 2 * D;
    for (int i = threadIdx.x; i < 2 * D; i += 256) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += sm_acc[(long long)w * 2 * D + i];
      dst[i] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Residual-branch backward: dx (fp32 grad of the block output x_out = x_in + gate * y) ->
//   dy[row, :]   = bf16(dx * gate[n, :])           (operand of the branch's dgrad / wgrad GEMMs)
//   part[n, chunk, 0, :] = sum_rows dx             (-> bias gradient: db = sum_n gate[n] * s[n])
//   part[n, chunk, 1, :] = sum_rows dx * y         (-> dgate[n])
// gate == null: plain residual (U-ViT): dy = bf16(dx), only the column sum is produced.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gate_bwd_kernel(const float* __restrict__ dx, const bf16* __restrict__ y, const float* __restrict__ gate,
                long long ld_gate, bf16* __restrict__ dy, float* __restrict__ part, int rows_per_group, int chunks,
                int M, int D) {
  extern __shared__ float sm_acc[];  // [8][2][D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.x, group = blockIdx.y;
  const int nv = D >> 2;
  float* acc = sm_acc + (long long)warp * 2 * D;
  for (int i = lane; i < 2 * D; i += 32) acc[i] = 0.f;
  __syncwarp();
  const int rows_per_chunk = (rows_per_group + chunks - 1) / chunks;
  const int r_begin = group * rows_per_group + chunk * rows_per_chunk;
  const int r_end = min(min(r_begin + rows_per_chunk, (group + 1) * rows_per_group), M);
  const float4* gp = gate ? reinterpret_cast<const float4*>(gate + (long long)group * ld_gate) : nullptr;
  for (int row = r_begin + warp; row < r_end; row += 8) {
    const float4* dxr = reinterpret_cast<const float4*>(dx + (long long)row * D);
    const uint2* yr = y ? reinterpret_cast<const uint2*>(y + (long long)row * D) : nullptr;
    uint2* dyr = reinterpret_cast<uint2*>(dy + (long long)row * D);
    for (int idx = lane; idx < nv; idx += 32) {
      const float4 d = dxr[idx];
      float4 gt = make_float4(1.f, 1.f, 1.f, 1.f);
      if (gp) gt = __ldg(gp + idx);
      uint2 o;
      o.x = pack_bf16(d.x * gt.x, d.y * gt.y);
      o.y = pack_bf16(d.z * gt.z, d.w * gt.w);
      dyr[idx] = o;
      float4* aS = reinterpret_cast<float4*>(acc) + idx;
      float4 s = *aS;
      s.x += d.x; s.y += d.y; s.z += d.z; s.w += d.w;
      *aS = s;
      if (yr) {
        const uint2 yu = yr[idx];
        const float2 y0 = unpack_bf16(yu.x), y1 = unpack_bf16(yu.y);
        float4* aG = reinterpret_cast<float4*>(acc + D) + idx;
        float4 a = *aG;
        a.x += d.x * y0.x; a.y += d.y * y0.y; a.z += d.z * y1.x; a.w += d.w * y1.y;
        *aG = a;
      }
    }
  }
  __syncthreads();
  float* dst = part + ((long long)group * chunks + chunk) * 2 * D;
  for (int i = threadIdx.x; i < 2 * D; i += 256) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += sm_acc[(long long)w * 2 * D + i];
    dst[i] = s;
  }
}

// ---------------------------------------------------------------------------------------------------
// Finish kernels for the [groups, chunks, 2, D] partial buffers (fixed summation order).
//   which = 0/1 selects the first / second D-vector of each partial.
//   per-group:  out[g * ld_out + col] (+)= sum_c part[g, c, which, col]
//   all-groups: out[col] (+)= sum_g (w ? w[g * ld_w + col] : 1) * sum_c part[g, c, which, col]
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
finish_group_kernel(const float* __restrict__ part, int which, int groups, int chunks, int D, float* __restrict__ out,
                    long long ld_out, int accumulate) {
  const int col = blockIdx.x * 256 + threadIdx.x;
  const int g = blockIdx.y;
  if (col >= D) return;
  float s = 0.f;
  for (int c = 0; c < chunks; ++c) s += part[(((long long)g * chunks + c) * 2 + which) * D + col];
  float* o = out + (long long)g * ld_out + col;
  *o = accumulate ? *o + s : s;
}
__global__ void __launch_bounds__(1024)
finish_all_kernel(const float* __restrict__ part, int which, int groups, int chunks, int D,
                  const float* __restrict__ w, long long ld_w, float* __restrict__ out, int accumulate) {
  // block = 32 columns x 32 group-lanes; each group-lane folds groups gy, gy+32, ... in order, then lane 0 folds
  // the 32 partials in order: fixed summation tree -> deterministic
  __shared__ float red[32][33];
  const int cx = threadIdx.x & 31, gy = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + cx;
  float tot = 0.f;
  if (col < D) {
    for (int g = gy; g < groups; g += 32) {
      float s = 0.f;
      for (int c = 0; c < chunks; ++c) s += part[(((long long)g * chunks + c) * 2 + which) * D + col];
      tot += w ? w[(long long)g * ld_w + col] * s : s;
    }
  }
  red[gy][cx] = tot;
  __syncthreads();
  if (gy == 0 && col < D) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += red[i][cx];
    out[col] = accumulate ? out[col] + s : s;
  }
}

// ---------------------------------------------------------------------------------------------------
// Column sum of a bf16 [M, N] matrix -> fp32 [N] (bias gradients of qkv / fc1): two deterministic stages.
// stage 1: grid (N/64, row_chunks), 256 threads = 8 row groups x 32 lanes x 2 columns
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
colsum_bf16_stage1(const bf16* __restrict__ a, long long lda, int M, int N, int rows_per_chunk,
                   float* __restrict__ part) {
  const int lane = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int col = blockIdx.x * 64 + lane * 2;
  const int r0 = blockIdx.y * rows_per_chunk;
  const int r1 = min(r0 + rows_per_chunk, M);
  float s0 = 0.f, s1 = 0.f;
  if (col < N) {
    for (int r = r0 + rg; r < r1; r += 8) {
      const float2 v = unpack_bf16(*reinterpret_cast<const uint32_t*>(a + (long long)r * lda + col));
      s0 += v.x;
      s1 += v.y;
    }
  }
  __shared__ float red[8][64];
  red[rg][lane * 2] = s0;
  red[rg][lane * 2 + 1] = s1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    const int c = blockIdx.x * 64 + threadIdx.x;
    if (c < N) part[(long long)blockIdx.y * N + c] = s;
  }
}
__global__ void __launch_bounds__(256)
colsum_stage2(const float* __restrict__ part, int chunks, int N, float* __restrict__ out, int accumulate) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= N) return;
  float s = 0.f;
  for (int k = 0; k < chunks; ++k) s += part[(long long)k * N + c];
  out[c] = accumulate ? out[c] + s : s;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
extern "C" int vaw_ln_fwd(const float* x, const float* shift, const float* scale, long long ld_mod,
                          int rows_per_sample, const float* weight, const float* bias, void* y, float* mean,
                          float* rstd, int M, int D, float eps, cudaStream_t stream) {
  VAW_CHECK_ARG(x && y && mean && rstd && M > 0, "vaw_ln_fwd: bad arguments");
  VAW_CHECK_ARG(D % 4 == 0 && D <= kMaxD, "vaw_ln_fwd: D=%d must be a multiple of 4 and <= %d", D, kMaxD);
  VAW_CHECK_ARG((scale == nullptr) == (shift == nullptr), "vaw_ln_fwd: shift and scale go together");
  VAW_CHECK_ARG((weight == nullptr) == (bias == nullptr), "vaw_ln_fwd: weight and bias go together");
  VAW_CHECK_ARG(!scale || rows_per_sample > 0, "vaw_ln_fwd: rows_per_sample");
#define VAW_LN_FWD(KV)                                                                                        \
  ln_fwd_kernel<KV><<<(M + 7) / 8, 256, 0, stream>>>(x, shift, scale, ld_mod, rows_per_sample > 0 ? rows_per_sample : 1, \
                                                     weight, bias, (bf16*)y, mean, rstd, M, D, eps)
  if (D <= 384) VAW_LN_FWD(3);
  else if (D <= 768) VAW_LN_FWD(6);
  else if (D <= 1152) VAW_LN_FWD(9);
  else VAW_LN_FWD(16);
#undef VAW_LN_FWD
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

// chunks: number of partial rows per group; part must hold groups * chunks * 2 * D floats (may be NULL when
// neither dA nor dB is wanted).  groups * rows_per_group must cover M.
extern "C" int vaw_ln_bwd(const void* dy, const float* x, const float* mean, const float* rstd, const float* scale,
                          long long ld_mod, const float* weight, float* dx_io, int add_into, float* part,
                          int rows_per_group, int groups, int chunks, int M, int D, cudaStream_t stream) {
  VAW_CHECK_ARG(dy && x && mean && rstd && dx_io && M > 0, "vaw_ln_bwd: bad arguments");
  VAW_CHECK_ARG(D % 4 == 0 && D <= kMaxD, "vaw_ln_bwd: D=%d must be a multiple of 4 and <= %d", D, kMaxD);
  VAW_CHECK_ARG(rows_per_group > 0 && groups > 0 && chunks > 0 && (long long)groups * rows_per_group >= M,
                "vaw_ln_bwd: bad grouping");
  const size_t smem = (size_t)8 * 2 * D * sizeof(float);
  static bool configured = false;
  if (!configured) {
    VAW_CUDA_TRY(cudaFuncSetAttribute(ln_bwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * kMaxD * 4));
    VAW_CUDA_TRY(cudaFuncSetAttribute(ln_bwd_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * kMaxD * 4));
    VAW_CUDA_TRY(cudaFuncSetAttribute(ln_bwd_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * kMaxD * 4));
    VAW_CUDA_TRY(cudaFuncSetAttribute(ln_bwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * kMaxD * 4));
    configured = true;
  }
#define VAW_LN_BWD(KV)                                                                                          \
  ln_bwd_kernel<KV><<<dim3(chunks, groups), 256, smem, stream>>>((const bf16*)dy, x, mean, rstd, scale, ld_mod, weight, \
                                                                 dx_io, add_into, part, rows_per_group, chunks, M, D)
  if (D <= 384) VAW_LN_BWD(3);
  else if (D <= 768) VAW_LN_BWD(6);
  else if (D <= 1152) VAW_LN_BWD(9);
  else VAW_LN_BWD(16);
#undef VAW_LN_BWD
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

extern "C" int vaw_gate_bwd(const float* dx, const void* y, const float* gate, long long ld_gate, void* dy,
                            float* part, int rows_per_group, int groups, int chunks, int M, int D,
                            cudaStream_t stream) {
  VAW_CHECK_ARG(dx && dy && part && M > 0, "vaw_gate_bwd: bad arguments");
  VAW_CHECK_ARG(D % 4 == 0 && D <= kMaxD, "vaw_gate_bwd: D=%d must be a multiple of 4 and <= %d", D, kMaxD);
  VAW_CHECK_ARG(rows_per_group > 0 && groups > 0 && chunks > 0 && (long long)groups * rows_per_group >= M,
                "vaw_gate_bwd: bad grouping");
  const size_t smem = (size_t)8 * 2 * D * sizeof(float);
  static bool configured = false;
  if (!configured) {
    VAW_CUDA_TRY(cudaFuncSetAttribute(gate_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * kMaxD * 4));
    configured = true;
  }
  gate_bwd_kernel<<<dim3(chunks, groups), 256, smem, stream>>>(dx, (const bf16*)y, gate, ld_gate, (bf16*)dy, part,
                                                               rows_per_group, chunks, M, D);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

extern "C" int vaw_finish_group(const float* part, int which, int groups, int chunks, int D, float* out,
                                long long ld_out, int accumulate, cudaStream_t stream) {
  VAW_CHECK_ARG(part && out && (which == 0 || which == 1) && groups > 0 && chunks > 0 && D > 0,
                "vaw_finish_group: bad arguments");
  finish_group_kernel<<<dim3((D + 255) / 256, groups), 256, 0, stream>>>(part, which, groups, chunks, D, out, ld_out,
                                                                         accumulate);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

extern "C" int vaw_finish_all(const float* part, int which, int groups, int chunks, int D, const float* w,
                              long long ld_w, float* out, int accumulate, cudaStream_t stream) {
  VAW_CHECK_ARG(part && out && (which == 0 || which == 1) && groups > 0 && chunks > 0 && D > 0,
                "vaw_finish_all: bad arguments");
  finish_all_kernel<<<(D + 31) / 32, 1024, 0, stream>>>(part, which, groups, chunks, D, w, ld_w, out, accumulate);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

// part: scratch of ceil(M / rows_per_chunk) * N floats
extern "C" int vaw_colsum_bf16(const void* a, long long lda, int M, int N, float* part, int rows_per_chunk,
                               float* out, int accumulate, cudaStream_t stream) {
  VAW_CHECK_ARG(a && part && out && M > 0 && N > 0 && N % 2 == 0 && rows_per_chunk > 0, "vaw_colsum_bf16: bad arguments");
  const int chunks = (M + rows_per_chunk - 1) / rows_per_chunk;
  colsum_bf16_stage1<<<dim3((N + 63) / 64, chunks), 256, 0, stream>>>((const bf16*)a, lda, M, N, rows_per_chunk, part);
  VAW_LAUNCH_CHECK();
  colsum_stage2<<<(N + 255) / 256, 256, 0, stream>>>(part, chunks, N, out, accumulate);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

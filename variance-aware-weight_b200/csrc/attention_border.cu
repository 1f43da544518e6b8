// attention_border.cu — sequences a few tokens longer than the tcgen05 attention kernels' 256-token tile (U-ViT:
// 256 patches + time and label tokens = 258, models/uvit.py:221-231; the MoCo-v3 ViT teacher: 256 patches + cls = 257).
//
// The tensor-core kernels (attention_sm100.cu, attention_bwd_sm100.cu) keep the whole key range of a head in TMEM, which
// caps them at T = 256.  Instead of a third, nearly empty 128-row tile per head (+50 % work for 2 tokens) the attention
// matrix is split as
//        main   = queries [0, 256) x keys [0, 256)      -> tensor cores, unchanged kernels on the leading 256 tokens
//        border = the remaining L-shaped strip           -> this file, CUDA cores (nb = T - 256 <= 8 tokens)
// Forward: the main kernel leaves O_A (normalised over its 256 keys) and the log-sum-exp L_A; the border pass folds the
// nb extra keys into every main row by a log-sum-exp merge (L = log2(2^L_A + sum_j 2^s_ij), O = O_A 2^(L_A - L) +
// sum_j 2^(s_ij - L) v_j) and evaluates the nb extra query rows over all T keys.  Backward: the main kernel is run with
// the FINAL L and Delta = rowsum(dO o O), so its P = 2^(s - L) are the true probabilities and its dQ / dK / dV are exact
// partial sums; the border pass adds the strip's contributions (read-modify-write of the rows it owns).
// One CTA per (batch, head), one thread per token; all reductions in a fixed order (deterministic).
#include "vaw_common.cuh"
#include "vaw_internal.h"

namespace {

constexpr int kMain = 256;
constexpr int kMaxBorder = 8;
constexpr int kThreads = 288;   // >= kMain + kMaxBorder, multiple of 32

template <int HD>
__device__ __forceinline__ void load_row(const bf16* __restrict__ p, float (&r)[HD]) {
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p) + i);
    const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
    r[8 * i] = a.x; r[8 * i + 1] = a.y; r[8 * i + 2] = b.x; r[8 * i + 3] = b.y;
    r[8 * i + 4] = c.x; r[8 * i + 5] = c.y; r[8 * i + 6] = d.x; r[8 * i + 7] = d.y;
  }
}

template <int HD>
__device__ __forceinline__ void store_row(bf16* __restrict__ p, const float (&r)[HD]) {
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) {
    uint4 u;
    u.x = pack_bf16(r[8 * i], r[8 * i + 1]); u.y = pack_bf16(r[8 * i + 2], r[8 * i + 3]);
    u.z = pack_bf16(r[8 * i + 4], r[8 * i + 5]); u.w = pack_bf16(r[8 * i + 6], r[8 * i + 7]);
    reinterpret_cast<uint4*>(p)[i] = u;
  }
}

// dot of a register row with a 16-byte aligned shared-memory row (broadcast 128-bit loads, two accumulation chains)
template <int HD>
__device__ __forceinline__ float dot_smem(const float (&r)[HD], const float* __restrict__ s) {
  const float4* s4 = reinterpret_cast<const float4*>(s);
  float a0 = 0.f, a1 = 0.f;
#pragma unroll
  for (int d = 0; d < HD / 4; ++d) {
    const float4 v = s4[d];
    a0 = fmaf(r[4 * d], v.x, a0);
    a1 = fmaf(r[4 * d + 1], v.y, a1);
    a0 = fmaf(r[4 * d + 2], v.z, a0);
    a1 = fmaf(r[4 * d + 3], v.w, a1);
  }
  return a0 + a1;
}

// block-wide sum / max over kThreads values in a fixed tree (warp shuffle, then 9 warp partials in order)
__device__ __forceinline__ float block_reduce(float v, bool is_max, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float w = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, w) : v + w;
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int i = 1; i < kThreads / 32; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
  return r;
}

// A [rows][HD] bf16 matrix staged in shared memory.  Rows are HD * 2 + 16 bytes apart: a warp whose lanes each read
// their own row in 16-byte pieces then spreads over all banks.  Staging is cooperative - 8 (9) consecutive threads fetch
// one 128 (144) byte row - so every global access of these kernels is a fully used line.
template <int HD>
struct Stage {
  static constexpr int kRowBytes = HD * 2 + 16;
  static constexpr int kCh = HD / 8;   // 16-byte chunks per row
  // rows [0, R) of the matrix whose row r starts at g + r * row_stride (elements)
  static constexpr int kIters = ((kMain + kMaxBorder) * kCh + kThreads - 1) / kThreads;
  static __device__ __forceinline__ void fill(uint8_t* sm, const bf16* __restrict__ g, long long row_stride, int R) {
    // all loads of a thread are issued before its first store (the kernel runs few warps per SM: a load -> store ->
    // load chain would pay the memory latency once per chunk)
    uint4 buf[kIters];
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int ch = threadIdx.x + it * kThreads;
      if (ch < R * kCh) {
        const int r = ch / kCh, part = ch - r * kCh;
        buf[it] = __ldg(reinterpret_cast<const uint4*>(g + r * row_stride) + part);
      }
    }
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int ch = threadIdx.x + it * kThreads;
      if (ch < R * kCh) {
        const int r = ch / kCh, part = ch - r * kCh;
        *reinterpret_cast<uint4*>(sm + (size_t)r * kRowBytes + part * 16) = buf[it];
      }
    }
  }
  static __device__ __forceinline__ void row(const uint8_t* sm, int r, float (&out)[HD]) {
    const uint4* src = reinterpret_cast<const uint4*>(sm + (size_t)r * kRowBytes);
#pragma unroll
    for (int i = 0; i < kCh; ++i) {
      const uint4 u = src[i];
      const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
      out[8 * i] = a.x; out[8 * i + 1] = a.y; out[8 * i + 2] = b.x; out[8 * i + 3] = b.y;
      out[8 * i + 4] = c.x; out[8 * i + 5] = c.y; out[8 * i + 6] = d.x; out[8 * i + 7] = d.y;
    }
  }
  static __device__ __forceinline__ float at(const uint8_t* sm, int r, int d) {
    return __bfloat162float(reinterpret_cast<const bf16*>(sm + (size_t)r * kRowBytes)[d]);
  }
};

// rows [0, R) of a global bf16 matrix (row r at g + r * row_stride): row_r = row_r * w0[r] + sum_j w[j][r] * v[j][:]
// (w0 == nullptr: weight 1), 16-byte chunks handed out so that consecutive threads touch consecutive memory
template <int HD, int LD>
__device__ __forceinline__ void rank_update(bf16* __restrict__ g, long long row_stride, int R, const float* __restrict__ w0,
                                            const float (*w)[LD], const float (*v)[HD], int nb) {
  constexpr int kCh = HD / 8;
  constexpr int kIters = ((kMain + kMaxBorder) * kCh + kThreads - 1) / kThreads;
  uint4 buf[kIters];
#pragma unroll
  for (int it = 0; it < kIters; ++it) {   // every load first (see Stage::fill)
    const int ch = threadIdx.x + it * kThreads;
    if (ch < R * kCh) {
      const int r = ch / kCh, part = ch - r * kCh;
      buf[it] = *(reinterpret_cast<const uint4*>(g + r * row_stride) + part);
    }
  }
#pragma unroll
  for (int it = 0; it < kIters; ++it) {
    const int ch = threadIdx.x + it * kThreads;
    if (ch >= R * kCh) continue;
    const int r = ch / kCh, part = ch - r * kCh;
    uint4* p = reinterpret_cast<uint4*>(g + r * row_stride) + part;
    const uint4 u = buf[it];
    float x[8];
    const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
    x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y; x[4] = c.x; x[5] = c.y; x[6] = d.x; x[7] = d.y;
    if (w0) {
      const float s = w0[r];
#pragma unroll
      for (int e = 0; e < 8; ++e) x[e] *= s;
    }
    for (int j = 0; j < nb; ++j) {
      const float wj = w[j][r];
      const float4 v0 = *reinterpret_cast<const float4*>(&v[j][part * 8]);
      const float4 v1 = *reinterpret_cast<const float4*>(&v[j][part * 8 + 4]);
      x[0] = fmaf(wj, v0.x, x[0]); x[1] = fmaf(wj, v0.y, x[1]); x[2] = fmaf(wj, v0.z, x[2]); x[3] = fmaf(wj, v0.w, x[3]);
      x[4] = fmaf(wj, v1.x, x[4]); x[5] = fmaf(wj, v1.y, x[5]); x[6] = fmaf(wj, v1.z, x[6]); x[7] = fmaf(wj, v1.w, x[7]);
    }
    uint4 o;
    o.x = pack_bf16(x[0], x[1]); o.y = pack_bf16(x[2], x[3]); o.z = pack_bf16(x[4], x[5]); o.w = pack_bf16(x[6], x[7]);
    *p = o;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// forward border pass
// ---------------------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(kThreads, 2)
attn_border_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ o, float* __restrict__ lse2, int T, int H,
                       float c /* scale * log2(e) */) {
  const int nb = T - kMain;
  const int h = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
  extern __shared__ __align__(16) uint8_t stage[];     // one [T][HD] bf16 matrix at a time: Q, then K, then V
  __shared__ __align__(16) float sk[kMaxBorder][HD], sv[kMaxBorder][HD], sq[kMaxBorder][HD];
  __shared__ float sp[kMaxBorder][kThreads];     // weights of the extra keys per main row, then the border rows' probabilities
  __shared__ float s_w0[kMain];                  // rescale factor of the main rows' O_A
  __shared__ float red[kThreads / 32];
  const long long row_stride = 3LL * H * HD;
  const bf16* base = qkv + (long long)b * T * row_stride + (long long)h * HD;
  for (int i = t; i < nb * HD; i += kThreads) {
    const int j = i / HD, d = i - j * HD;
    const bf16* r = base + (kMain + j) * row_stride;
    sq[j][d] = __bfloat162float(r[d]);
    sk[j][d] = __bfloat162float(r[(long long)H * HD + d]);
    sv[j][d] = __bfloat162float(r[2LL * H * HD + d]);
  }
  Stage<HD>::fill(stage, base, row_stride, kMain);                                  // Q of the main queries
  __syncthreads();
  float* lrow = lse2 + ((long long)b * H + h) * T;
  // ---- main queries: weights that fold the nb extra keys into the row the tensor-core kernel produced ----
  if (t < kMain) {
    float r[HD];
    Stage<HD>::row(stage, t, r);
    float s[kMaxBorder];
    const float la = lrow[t];
    float m = la;
    for (int j = 0; j < nb; ++j) {
      s[j] = dot_smem<HD>(r, sk[j]) * c;
      m = fmaxf(m, s[j]);
    }
    const float wa = exp2f(la - m);
    float sum = wa;
    for (int j = 0; j < nb; ++j) {
      s[j] = exp2f(s[j] - m);
      sum += s[j];
    }
    const float inv = 1.f / sum;
    s_w0[t] = wa * inv;
    for (int j = 0; j < nb; ++j) sp[j][t] = s[j] * inv;
    lrow[t] = m + log2f(sum);
  }
  __syncthreads();
  rank_update<HD, kThreads>(o + ((long long)b * T * H + h) * HD, (long long)H * HD, kMain, s_w0, sp, sv, nb);
  Stage<HD>::fill(stage, base + (long long)H * HD, row_stride, T);                  // K of every key
  __syncthreads();
  // ---- border queries: full rows over all T keys (thread t = key t) ----
  {
    float kr[HD];
    if (t < T) Stage<HD>::row(stage, t, kr);
    for (int i = 0; i < nb; ++i) {
      const float s = t < T ? dot_smem<HD>(kr, sq[i]) * c : -INFINITY;
      const float m = block_reduce(s, true, red);
      const float p = t < T ? exp2f(s - m) : 0.f;
      const float sum = block_reduce(p, false, red);
      sp[i][t] = p / sum;
      if (t == 0) lrow[kMain + i] = m + log2f(sum);
    }
  }
  __syncthreads();
  Stage<HD>::fill(stage, base + 2LL * H * HD, row_stride, T);                       // V of every key
  __syncthreads();
  // O[i, d] = sum_t p[i, t] v[t, d]: thread (i, d) walks the staged V in key order
  for (int e = t; e < nb * HD; e += kThreads) {
    const int i = e / HD, d = e - i * HD;
    float acc = 0.f;
#pragma unroll 8
    for (int k = 0; k < T; ++k) acc = fmaf(sp[i][k], Stage<HD>::at(stage, k, d), acc);
    o[(((long long)b * T + kMain + i) * H + h) * HD + d] = __float2bfloat16(acc);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// backward border pass (after the tensor-core kernel has written the main block's dq / dk / dv rows [0, 256))
// ---------------------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(kThreads, 2)
attn_border_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ d_o, const float* __restrict__ lse2,
                       const float* __restrict__ delta, bf16* __restrict__ dqkv, int T, int H, float scale, float c) {
  constexpr int kPer = (kMaxBorder * HD + kThreads - 1) / kThreads;   // (border token, column) pairs per thread
  const int nb = T - kMain;
  const int h = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
  extern __shared__ __align__(16) uint8_t smem_dyn[];   // two [T][HD] bf16 matrices: (Q, dO) of the main rows, then (K, V)
  uint8_t* stA = smem_dyn;
  uint8_t* stB = smem_dyn + (size_t)(kMain + kMaxBorder) * Stage<HD>::kRowBytes;
  // border tokens' rows (as keys: k, v; as queries: q, dO); sk / sv are re-used for the border keys' part-B gradients
  __shared__ __align__(16) float sk[kMaxBorder][HD], sv[kMaxBorder][HD], sq[kMaxBorder][HD], sdo[kMaxBorder][HD];
  __shared__ float sX[kMaxBorder][kThreads], sY[kMaxBorder][kThreads];   // per-row weights of the rank-nb updates
  const long long row_stride = 3LL * H * HD;
  const long long boff = (long long)b * T * row_stride + (long long)h * HD;
  const bf16* base = qkv + boff;
  bf16* gbase = dqkv + boff;
  const bf16* dobase = d_o + ((long long)b * T * H + h) * HD;
  const long long do_stride = (long long)H * HD;
  for (int i = t; i < nb * HD; i += kThreads) {
    const int j = i / HD, d = i - j * HD;
    const bf16* r = base + (kMain + j) * row_stride;
    sq[j][d] = __bfloat162float(r[d]);
    sk[j][d] = __bfloat162float(r[(long long)H * HD + d]);
    sv[j][d] = __bfloat162float(r[2LL * H * HD + d]);
    sdo[j][d] = __bfloat162float(dobase[(kMain + j) * do_stride + d]);
  }
  Stage<HD>::fill(stA, base, row_stride, kMain);            // Q of the main queries
  Stage<HD>::fill(stB, dobase, do_stride, kMain);           // dO of the main queries
  __syncthreads();
  const float* lrow = lse2 + ((long long)b * H + h) * T;
  const float* drow = delta + ((long long)b * H + h) * T;

  // ---- part A: main query t x border keys: P_tj (sX) and dS_tj (sY) ----
  if (t < kMain) {
    float r[HD];
    float sdot[kMaxBorder];
    Stage<HD>::row(stA, t, r);
    for (int j = 0; j < nb; ++j) sdot[j] = dot_smem<HD>(r, sk[j]);
    Stage<HD>::row(stB, t, r);
    const float L = lrow[t], Dl = drow[t];
    for (int j = 0; j < nb; ++j) {
      const float p = exp2f(sdot[j] * c - L);
      sX[j][t] = p;
      sY[j][t] = p * (dot_smem<HD>(r, sv[j]) - Dl);
    }
  }
  __syncthreads();
  // border keys, part A share: dK_j[d] = sum_i dS_ij q_i[d], dV_j[d] = sum_i P_ij dO_i[d] over the main queries in order
  float ak[kPer], av[kPer];
#pragma unroll
  for (int u = 0; u < kPer; ++u) {
    const int e = t + u * kThreads;
    ak[u] = av[u] = 0.f;
    if (e < nb * HD) {
      const int j = e / HD, d = e - j * HD;
      float a0 = 0.f, a1 = 0.f;
#pragma unroll 8
      for (int i = 0; i < kMain; ++i) {
        a0 = fmaf(sY[j][i], Stage<HD>::at(stA, i, d), a0);
        a1 = fmaf(sX[j][i], Stage<HD>::at(stB, i, d), a1);
      }
      ak[u] = a0;
      av[u] = a1;
    }
  }
  __syncthreads();   // every (j, d) thread is done reading the unscaled dS
  // dQ_t += scale sum_j dS_tj k_j on the rows the tensor-core kernel wrote (scale folded into the weights)
  if (t < kMain) {
    for (int j = 0; j < nb; ++j) sY[j][t] *= scale;
  }
  __syncthreads();
  rank_update<HD, kThreads>(gbase, row_stride, kMain, nullptr, sY, sk, nb);
  Stage<HD>::fill(stA, base + (long long)H * HD, row_stride, T);     // K of every key
  Stage<HD>::fill(stB, base + 2LL * H * HD, row_stride, T);          // V of every key
  __syncthreads();
  // ---- part B: border queries x key t: P_it (sX), scale * dS_it (sY) ----
  if (t < T) {
    float r[HD];
    float pi[kMaxBorder];
    Stage<HD>::row(stA, t, r);
    for (int i = 0; i < nb; ++i) pi[i] = exp2f(dot_smem<HD>(r, sq[i]) * c - lrow[kMain + i]);
    Stage<HD>::row(stB, t, r);
    for (int i = 0; i < nb; ++i) {
      sX[i][t] = pi[i];
      sY[i][t] = pi[i] * (dot_smem<HD>(r, sdo[i]) - drow[kMain + i]) * scale;
    }
  }
  __syncthreads();
  // main keys: dK_t += sum_i (scale dS_it) q_i, dV_t += sum_i P_it dO_i on the rows the tensor-core kernel wrote
  rank_update<HD, kThreads>(gbase + (long long)H * HD, row_stride, kMain, nullptr, sY, sq, nb);
  rank_update<HD, kThreads>(gbase + 2LL * H * HD, row_stride, kMain, nullptr, sX, sdo, nb);
  // ---- final: border queries' dQ_j = sum_k (scale dS_jk) k_k over ALL keys; border keys' dK / dV = part A's share plus
  //      their part-B share sum_i (scale dS_i,256+j) q_i / sum_i P_i,256+j dO_i ----
#pragma unroll
  for (int u = 0; u < kPer; ++u) {
    const int e = t + u * kThreads;
    if (e < nb * HD) {
      const int j = e / HD, d = e - j * HD;
      float acc = 0.f;
#pragma unroll 8
      for (int k = 0; k < T; ++k) acc = fmaf(sY[j][k], Stage<HD>::at(stA, k, d), acc);
      float bk = ak[u] * scale, bv = av[u];
      for (int i = 0; i < nb; ++i) {
        bk = fmaf(sY[i][kMain + j], sq[i][d], bk);
        bv = fmaf(sX[i][kMain + j], sdo[i][d], bv);
      }
      bf16* row = gbase + (kMain + j) * row_stride;
      row[d] = __float2bfloat16(acc);
      row[(long long)H * HD + d] = __float2bfloat16(bk);
      row[2LL * H * HD + d] = __float2bfloat16(bv);
    }
  }
}

template <int HD>
int border_smem_bytes(int mats) { return mats * (kMain + kMaxBorder) * Stage<HD>::kRowBytes; }

// static + dynamic shared memory exceeds the 48 KB default: opt in once per kernel
template <int HD>
int border_configure() {
  static bool done = false;
  if (!done) {
    VAW_CUDA_TRY(cudaFuncSetAttribute(attn_border_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      border_smem_bytes<HD>(1)));
    VAW_CUDA_TRY(cudaFuncSetAttribute(attn_border_bwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      border_smem_bytes<HD>(2)));
    done = true;
  }
  return VAW_OK;
}

}  // namespace

int vaw_attn_border_supported(int T) { return T > kMain && T - kMain <= kMaxBorder; }

int vaw_attn_border_fwd(const void* qkv, void* o, float* lse2, int B, int T, int H, int head_dim, cudaStream_t stream) {
  const float c = 1.4426950408889634f / sqrtf((float)head_dim);
  if (head_dim == 64) { int rc = border_configure<64>(); if (rc) return rc; }
  if (head_dim == 72) { int rc = border_configure<72>(); if (rc) return rc; }
  if (head_dim == 64)
    attn_border_fwd_kernel<64><<<dim3(H, B), kThreads, border_smem_bytes<64>(1), stream>>>((const bf16*)qkv, (bf16*)o,
                                                                                         lse2, T, H, c);
  else if (head_dim == 72)
    attn_border_fwd_kernel<72><<<dim3(H, B), kThreads, border_smem_bytes<72>(1), stream>>>((const bf16*)qkv, (bf16*)o,
                                                                                         lse2, T, H, c);
  else
    return VAW_ERR_UNSUPPORTED;
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

int vaw_attn_border_bwd(const void* qkv, const void* d_o, const float* lse2, const float* delta, void* dqkv, int B, int T,
                        int H, int head_dim, cudaStream_t stream) {
  const float scale = 1.0f / sqrtf((float)head_dim);
  const float c = scale * 1.4426950408889634f;
  if (head_dim == 64) { int rc = border_configure<64>(); if (rc) return rc; }
  if (head_dim == 72) { int rc = border_configure<72>(); if (rc) return rc; }
  if (head_dim == 64)
    attn_border_bwd_kernel<64><<<dim3(H, B), kThreads, border_smem_bytes<64>(2), stream>>>(
        (const bf16*)qkv, (const bf16*)d_o, lse2, delta, (bf16*)dqkv, T, H, scale, c);
  else if (head_dim == 72)
    attn_border_bwd_kernel<72><<<dim3(H, B), kThreads, border_smem_bytes<72>(2), stream>>>(
        (const bf16*)qkv, (const bf16*)d_o, lse2, delta, (bf16*)dqkv, T, H, scale, c);
  else
    return VAW_ERR_UNSUPPORTED;
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

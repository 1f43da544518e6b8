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

template <int HD>
__device__ __forceinline__ float dot_smem(const float (&r)[HD], const float* __restrict__ s) {
  float acc = 0.f;
#pragma unroll
  for (int d = 0; d < HD; ++d) acc = fmaf(r[d], s[d], acc);
  return acc;
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

// ---------------------------------------------------------------------------------------------------------------
// forward border pass
// ---------------------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(kThreads)
attn_border_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ o, float* __restrict__ lse2, int T, int H,
                       float c /* scale * log2(e) */) {
  const int nb = T - kMain;
  const int h = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
  __shared__ float sk[kMaxBorder][HD], sv[kMaxBorder][HD], sq[kMaxBorder][HD];
  __shared__ float sp[kMaxBorder][kThreads];     // probabilities of the border queries over all keys
  __shared__ float red[kThreads / 32];
  const long long row_stride = 3LL * H * HD;
  const bf16* base = qkv + (long long)b * T * row_stride + (long long)h * HD;
  auto qrow = [&](int tok) { return base + tok * row_stride; };
  auto krow = [&](int tok) { return base + tok * row_stride + (long long)H * HD; };
  auto vrow = [&](int tok) { return base + tok * row_stride + 2LL * H * HD; };
  for (int i = t; i < nb * HD; i += kThreads) {
    const int j = i / HD, d = i - j * HD;
    sk[j][d] = __bfloat162float(krow(kMain + j)[d]);
    sv[j][d] = __bfloat162float(vrow(kMain + j)[d]);
    sq[j][d] = __bfloat162float(qrow(kMain + j)[d]);
  }
  __syncthreads();
  float* lrow = lse2 + ((long long)b * H + h) * T;
  // ---- main queries: fold the nb extra keys into the row the tensor-core kernel produced ----
  if (t < kMain) {
    float r[HD];                       // the query row first, then the output row (one live row keeps registers low)
    load_row<HD>(qrow(t), r);
    float s[kMaxBorder];
    const float la = lrow[t];
    float m = la;
    for (int j = 0; j < nb; ++j) {
      s[j] = dot_smem<HD>(r, sk[j]) * c;
      m = fmaxf(m, s[j]);
    }
    float wa = exp2f(la - m), sum = wa;
    for (int j = 0; j < nb; ++j) {
      s[j] = exp2f(s[j] - m);
      sum += s[j];
    }
    const float inv = 1.f / sum;
    bf16* orow = o + (((long long)b * T + t) * H + h) * HD;
    load_row<HD>(orow, r);
    wa *= inv;
#pragma unroll
    for (int d = 0; d < HD; ++d) r[d] *= wa;
    for (int j = 0; j < nb; ++j) {
      const float w = s[j] * inv;
#pragma unroll
      for (int d = 0; d < HD; ++d) r[d] = fmaf(w, sv[j][d], r[d]);
    }
    store_row<HD>(orow, r);
    lrow[t] = m + log2f(sum);
  }
  // ---- border queries: full rows over all T keys (thread t = key t) ----
  float kr[HD];
  if (t < T) load_row<HD>(krow(t), kr);
  for (int i = 0; i < nb; ++i) {
    const float s = t < T ? dot_smem<HD>(kr, sq[i]) * c : -INFINITY;
    const float m = block_reduce(s, true, red);
    const float p = t < T ? exp2f(s - m) : 0.f;
    const float sum = block_reduce(p, false, red);
    sp[i][t] = p / sum;
    if (t == 0) lrow[kMain + i] = m + log2f(sum);
  }
  __syncthreads();
  // O[i, d] = sum_t p[i, t] v[t, d]: thread (i, d) walks the keys in order
  for (int e = t; e < nb * HD; e += kThreads) {
    const int i = e / HD, d = e - i * HD;
    float acc = 0.f;
    for (int k = 0; k < T; ++k) acc = fmaf(sp[i][k], __bfloat162float(vrow(k)[d]), acc);
    o[(((long long)b * T + kMain + i) * H + h) * HD + d] = __float2bfloat16(acc);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// backward border pass (after the tensor-core kernel has written the main block's dq / dk / dv rows [0, 256))
// ---------------------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(kThreads)
attn_border_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ d_o, const float* __restrict__ lse2,
                       const float* __restrict__ delta, bf16* __restrict__ dqkv, int T, int H, float scale, float c) {
  const int nb = T - kMain;
  const int h = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
  // border tokens' rows (as keys: k, v; as queries: q, dO) and their per-query statistics
  __shared__ float sk[kMaxBorder][HD], sv[kMaxBorder][HD], sq[kMaxBorder][HD], sdo[kMaxBorder][HD];
  __shared__ float sA_p[kMaxBorder][kMain], sA_ds[kMaxBorder][kMain];      // part A: P, dS of (main query, border key)
  __shared__ float sB_ds[kMaxBorder][kThreads];                            // part B: dS of (border query, key)
  const long long row_stride = 3LL * H * HD;
  const long long boff = (long long)b * T * row_stride + (long long)h * HD;
  const bf16* base = qkv + boff;
  bf16* gbase = dqkv + boff;
  auto slot = [&](const bf16* p, int tok, int s) { return p + tok * row_stride + (long long)s * H * HD; };
  auto gslot = [&](int tok, int s) { return gbase + tok * row_stride + (long long)s * H * HD; };
  auto dorow = [&](int tok) { return d_o + (((long long)b * T + tok) * H + h) * HD; };
  for (int i = t; i < nb * HD; i += kThreads) {
    const int j = i / HD, d = i - j * HD;
    sq[j][d] = __bfloat162float(slot(base, kMain + j, 0)[d]);
    sk[j][d] = __bfloat162float(slot(base, kMain + j, 1)[d]);
    sv[j][d] = __bfloat162float(slot(base, kMain + j, 2)[d]);
    sdo[j][d] = __bfloat162float(dorow(kMain + j)[d]);
  }
  __syncthreads();
  const float* lrow = lse2 + ((long long)b * H + h) * T;
  const float* drow = delta + ((long long)b * H + h) * T;

  // ---- part A: main query t x border keys.  dQ_t += scale sum_j dS_tj k_j (read-modify-write of the row the
  //      tensor-core kernel wrote); P_tj and dS_tj go to shared memory for the border keys' dK / dV ----
  if (t < kMain) {
    float r[HD];                       // q row, then dO row, then the dQ row: one live row at a time
    float sdot[kMaxBorder], w[kMaxBorder];
    load_row<HD>(slot(base, t, 0), r);
    for (int j = 0; j < nb; ++j) sdot[j] = dot_smem<HD>(r, sk[j]);
    load_row<HD>(dorow(t), r);
    const float L = lrow[t], Dl = drow[t];
    for (int j = 0; j < nb; ++j) {
      const float p = exp2f(sdot[j] * c - L);
      const float ds = p * (dot_smem<HD>(r, sv[j]) - Dl);
      sA_p[j][t] = p;
      sA_ds[j][t] = ds;
      w[j] = ds * scale;
    }
    load_row<HD>(gslot(t, 0), r);
    for (int j = 0; j < nb; ++j) {
#pragma unroll
      for (int d = 0; d < HD; ++d) r[d] = fmaf(w[j], sk[j][d], r[d]);
    }
    store_row<HD>(gslot(t, 0), r);
  }
  // ---- part B: border queries x key t.  dK_t += scale sum_i dS_it q_i, dV_t += sum_i P_it dO_i on the rows the thread
  //      owns: main keys read-modify-write what the tensor-core kernel wrote; border keys start from zero here and
  //      receive part A's share in the final pass ----
  if (t < T) {
    float r[HD];
    float pi[kMaxBorder], wi[kMaxBorder];
    load_row<HD>(slot(base, t, 1), r);
    for (int i = 0; i < nb; ++i) pi[i] = exp2f(dot_smem<HD>(r, sq[i]) * c - lrow[kMain + i]);
    load_row<HD>(slot(base, t, 2), r);
    for (int i = 0; i < nb; ++i) {
      const float ds = pi[i] * (dot_smem<HD>(r, sdo[i]) - drow[kMain + i]);
      sB_ds[i][t] = ds;
      wi[i] = ds * scale;
    }
    if (t < kMain) {
      load_row<HD>(gslot(t, 1), r);
    } else {
#pragma unroll
      for (int d = 0; d < HD; ++d) r[d] = 0.f;
    }
    for (int i = 0; i < nb; ++i) {
#pragma unroll
      for (int d = 0; d < HD; ++d) r[d] = fmaf(wi[i], sq[i][d], r[d]);
    }
    store_row<HD>(gslot(t, 1), r);
    if (t < kMain) {
      load_row<HD>(gslot(t, 2), r);
    } else {
#pragma unroll
      for (int d = 0; d < HD; ++d) r[d] = 0.f;
    }
    for (int i = 0; i < nb; ++i) {
#pragma unroll
      for (int d = 0; d < HD; ++d) r[d] = fmaf(pi[i], sdo[i][d], r[d]);
    }
    store_row<HD>(gslot(t, 2), r);
  }
  __syncthreads();
  // ---- reductions over the 256 main queries (fixed order): the border keys' dK / dV from part A, and the border
  //      queries' dQ = scale sum_t dS_it k_t over ALL keys ----
  for (int e = t; e < nb * HD; e += kThreads) {
    const int j = e / HD, d = e - j * HD;
    float ak = 0.f, av = 0.f;
    for (int i = 0; i < kMain; ++i) {
      ak = fmaf(sA_ds[j][i], __bfloat162float(slot(base, i, 0)[d]), ak);
      av = fmaf(sA_p[j][i], __bfloat162float(dorow(i)[d]), av);
    }
    bf16* pk = gslot(kMain + j, 1) + d;
    bf16* pv = gslot(kMain + j, 2) + d;
    *pk = __float2bfloat16(__bfloat162float(*pk) + ak * scale);
    *pv = __float2bfloat16(__bfloat162float(*pv) + av);
    float aq = 0.f;
    for (int k = 0; k < T; ++k) aq = fmaf(sB_ds[j][k], __bfloat162float(slot(base, k, 1)[d]), aq);
    gslot(kMain + j, 0)[d] = __float2bfloat16(aq * scale);
  }
}

}  // namespace

int vaw_attn_border_supported(int T) { return T > kMain && T - kMain <= kMaxBorder; }

int vaw_attn_border_fwd(const void* qkv, void* o, float* lse2, int B, int T, int H, int head_dim, cudaStream_t stream) {
  const float c = 1.4426950408889634f / sqrtf((float)head_dim);
  if (head_dim == 64)
    attn_border_fwd_kernel<64><<<dim3(H, B), kThreads, 0, stream>>>((const bf16*)qkv, (bf16*)o, lse2, T, H, c);
  else if (head_dim == 72)
    attn_border_fwd_kernel<72><<<dim3(H, B), kThreads, 0, stream>>>((const bf16*)qkv, (bf16*)o, lse2, T, H, c);
  else
    return VAW_ERR_UNSUPPORTED;
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

int vaw_attn_border_bwd(const void* qkv, const void* d_o, const float* lse2, const float* delta, void* dqkv, int B, int T,
                        int H, int head_dim, cudaStream_t stream) {
  const float scale = 1.0f / sqrtf((float)head_dim);
  const float c = scale * 1.4426950408889634f;
  if (head_dim == 64)
    attn_border_bwd_kernel<64><<<dim3(H, B), kThreads, 0, stream>>>((const bf16*)qkv, (const bf16*)d_o, lse2, delta,
                                                                    (bf16*)dqkv, T, H, scale, c);
  else if (head_dim == 72)
    attn_border_bwd_kernel<72><<<dim3(H, B), kThreads, 0, stream>>>((const bf16*)qkv, (const bf16*)d_o, lse2, delta,
                                                                    (bf16*)dqkv, T, H, scale, c);
  else
    return VAW_ERR_UNSUPPORTED;
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

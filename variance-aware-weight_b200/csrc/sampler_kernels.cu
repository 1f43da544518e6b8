// sampler_kernels.cu — K3: timestep importance sampling on the device, bit-exact with numpy.
//
// Replaces ScheduleSampler.sample (reference tools/resample.py:43-59), LossSecondMomentResampler.weights
// (:142-149) and .update_with_all_losses (:151-159).
//
// numpy arithmetic that is reproduced operation by operation in IEEE fp64 (no FMA contraction):
//   np.mean(h**2, axis=-1)   pairwise-sum leaf for n=10: ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then += r8, += r9, /10
//   np.sqrt                  correctly rounded
//   np.sum(w) over T         numpy pairwise_sum: split n/2 rounded down to a multiple of 8 until n <= 128,
//                            leaf = 8 strided accumulators + sequential tail
//   w /= sum; w *= (1-u); w += u/T
//   p = w / np.sum(w); cdf = cumsum(p) (sequential); cdf /= cdf[-1]
//   idx = searchsorted(cdf, u, side='right'); weight = float32(1 / (T * p[idx]))
// The uniform draws u come from numpy's global MT19937 on the host (np.random.random_sample(B) consumes exactly
// the stream np.random.choice(T, B, p=p) would), so the host RNG state stays interchangeable with the reference.
#include "vaw_common.cuh"

namespace {

constexpr int kMaxT = 2048;  // two fp64 [kMaxT] vectors in static shared memory (32 KB)

// numpy pairwise_sum leaf (n <= 128 or terminal), unrolled by 8
__device__ double np_leaf_sum(const double* a, int n) {
  if (n < 8) {
    double res = 0.;
    for (int i = 0; i < n; ++i) res = __dadd_rn(res, a[i]);
    return res;
  }
  double r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = a[j];
  int i;
  for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a[i + j]);
  }
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                         __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __dadd_rn(res, a[i]);
  return res;
}

// Leaf enumeration of numpy's recursion in DFS order (block-uniform, run by every thread on indices only).
struct LeafList {
  int start[64];
  int len[64];
  int count;
};
__device__ void enum_leaves(int start, int n, LeafList& L) {
  // iterative DFS with an explicit stack: numpy recursion is sum(a, n2) + sum(a + n2, n - n2)
  int st_s[16], st_n[16], sp = 0;
  st_s[sp] = start; st_n[sp] = n; ++sp;
  L.count = 0;
  while (sp > 0) {
    --sp;
    int s = st_s[sp], m = st_n[sp];
    if (m <= 128) {
      L.start[L.count] = s; L.len[L.count] = m; ++L.count;
    } else {
      int n2 = m / 2;
      n2 -= n2 % 8;
      // push right first so that left is processed first (DFS order = left to right)
      st_s[sp] = s + n2; st_n[sp] = m - n2; ++sp;
      st_s[sp] = s; st_n[sp] = n2; ++sp;
    }
  }
}
// combine the leaf sums with the same tree shape: returns sum over [start, start+n), consuming leaves in order
__device__ double combine_leaves(int n, const double* leaf_sums, int& cursor) {
  if (n <= 128) return leaf_sums[cursor++];
  int n2 = n / 2;
  n2 -= n2 % 8;
  double l = combine_leaves(n2, leaf_sums, cursor);
  double r = combine_leaves(n - n2, leaf_sums, cursor);
  return __dadd_rn(l, r);
}

// block-wide numpy-exact sum of a[0..n) living in shared memory; result broadcast through *out_smem
__device__ double block_np_sum(const double* a, int n, double* leaf_sums, double* out_smem) {
  LeafList L;
  enum_leaves(0, n, L);
  __syncthreads();
  if ((int)threadIdx.x < L.count) leaf_sums[threadIdx.x] = np_leaf_sum(a + L.start[threadIdx.x], L.len[threadIdx.x]);
  __syncthreads();
  if (threadIdx.x == 0) {
    int cursor = 0;
    *out_smem = combine_leaves(n, leaf_sums, cursor);
  }
  __syncthreads();
  return *out_smem;
}

// ------------------------------------------------------------------------------------------------
// sample: one CTA of 1024 threads.
//   mode 0: weights are given explicitly (w_in[T], e.g. UniformSampler's ones)
//   mode 1: LossSecondMomentResampler: weights from (history[T,H], counts[T]); uniform until warmed up
// Outputs idx[B] (int64), imp_w[B] (float32) and optionally the weights / p / cdf vectors (debug + parity).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
sampler_sample_kernel(int mode, const double* __restrict__ w_in, const double* __restrict__ history,
                      const int* __restrict__ counts, int T, int H, double uniform_prob,
                      const double* __restrict__ u, long long B, long long* __restrict__ idx,
                      float* __restrict__ imp_w, double* __restrict__ w_out, double* __restrict__ p_out,
                      double* __restrict__ cdf_out) {
  __shared__ double w[kMaxT];
  __shared__ double cdf[kMaxT];
  __shared__ double leaf_sums[64];
  __shared__ double bcast;
  __shared__ int not_warm;
  const int tid = threadIdx.x;

  if (tid == 0) not_warm = 0;
  __syncthreads();
  if (mode == 1) {
    for (int i = tid; i < T; i += blockDim.x)
      if (counts[i] != H) atomicOr(&not_warm, 1);
  }
  __syncthreads();

  if (mode == 0) {
    for (int i = tid; i < T; i += blockDim.x) w[i] = w_in[i];
  } else if (not_warm) {
    for (int i = tid; i < T; i += blockDim.x) w[i] = 1.0;
  } else {
    // weights = sqrt(mean(h**2, axis=-1))
    for (int i = tid; i < T; i += blockDim.x) {
      const double* h = history + (long long)i * H;
      double sq[16];
      double m;
      if (H <= 16) {
        for (int j = 0; j < H; ++j) sq[j] = __dmul_rn(h[j], h[j]);
        m = np_leaf_sum(sq, H);
      } else {
        // generic (H > 16 is outside the reference's configuration; sequential leaf semantics up to 128)
        double res = 0.;
        for (int j = 0; j < H; ++j) res = __dadd_rn(res, __dmul_rn(h[j], h[j]));
        m = res;
      }
      m = __ddiv_rn(m, (double)H);
      w[i] = __dsqrt_rn(m);
    }
    __syncthreads();
    const double s = block_np_sum(w, T, leaf_sums, &bcast);
    const double keep = 1.0 - uniform_prob;          // python float arithmetic: 1 - self.uniform_prob
    const double add = uniform_prob / (double)T;      // self.uniform_prob / len(weights)
    for (int i = tid; i < T; i += blockDim.x) {
      double v = __ddiv_rn(w[i], s);
      v = __dmul_rn(v, keep);
      v = __dadd_rn(v, add);
      w[i] = v;
    }
  }
  __syncthreads();
  if (w_out)
    for (int i = tid; i < T; i += blockDim.x) w_out[i] = w[i];

  // p = w / np.sum(w)
  const double sw = block_np_sum(w, T, leaf_sums, &bcast);
  for (int i = tid; i < T; i += blockDim.x) w[i] = __ddiv_rn(w[i], sw);  // w now holds p
  __syncthreads();
  if (p_out)
    for (int i = tid; i < T; i += blockDim.x) p_out[i] = w[i];

  // cdf = p.cumsum() — sequential by definition; then cdf /= cdf[-1]
  if (tid == 0) {
    double c = 0.;
    for (int i = 0; i < T; ++i) {
      c = (i == 0) ? w[0] : __dadd_rn(c, w[i]);
      cdf[i] = c;
    }
    bcast = c;
  }
  __syncthreads();
  const double last = bcast;
  for (int i = tid; i < T; i += blockDim.x) cdf[i] = __ddiv_rn(cdf[i], last);
  __syncthreads();
  if (cdf_out)
    for (int i = tid; i < T; i += blockDim.x) cdf_out[i] = cdf[i];

  // idx = searchsorted(cdf, u, side='right'): first i with cdf[i] > u
  for (long long b = tid; b < B; b += blockDim.x) {
    const double ub = u[b];
    int lo = 0, hi = T;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (cdf[mid] <= ub) lo = mid + 1; else hi = mid;
    }
    // numpy would return T for u >= cdf[-1]==1.0, which random_sample() in [0,1) never produces; clamp defensively
    int k = lo < T ? lo : T - 1;
    idx[b] = k;
    const double denom = __dmul_rn((double)T, w[k]);
    imp_w[b] = (float)__ddiv_rn(1.0, denom);
  }
}

// ------------------------------------------------------------------------------------------------
// history update: one thread per timestep scans the gathered (t, loss) entries in rank order, so duplicates
// of the same t inside one batch are applied sequentially exactly like the reference loop (:151-159).
// Entries with t < 0 are padding (ragged per-rank batch sizes) and skipped.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sampler_update_kernel(double* __restrict__ history, int* __restrict__ counts, const int* __restrict__ ts,
                      const float* __restrict__ losses, long long n_total, int T, int H) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  double* h = history + (long long)t * H;
  int c = counts[t];
  for (long long i = 0; i < n_total; ++i) {
    if (__ldg(ts + i) != t) continue;
    const double v = (double)__ldg(losses + i);  // exact widening, same as .item()/.tolist()
    if (c == H) {
      for (int j = 0; j + 1 < H; ++j) h[j] = h[j + 1];
      h[H - 1] = v;
    } else {
      h[c] = v;
      ++c;
    }
  }
  counts[t] = c;
}

// pack (int64 t, float loss) -> (int32 t, float loss) pairs for the single all_gather (K7)
__global__ void pack_tloss_kernel(const long long* __restrict__ t, const float* __restrict__ loss, int* __restrict__ t32,
                                  float* __restrict__ l32, long long B, long long Bpad) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Bpad) return;
  t32[i] = i < B ? (int)t[i] : -1;
  l32[i] = i < B ? loss[i] : 0.f;
}

}  // namespace

extern "C" int vaw_sampler_sample(int mode, const double* w_in, const double* history, const int* counts, int T, int H,
                                  double uniform_prob, const double* u, long long B, long long* idx, float* imp_w,
                                  double* w_out, double* p_out, double* cdf_out, cudaStream_t stream) {
  VAW_CHECK_ARG(mode == 0 || mode == 1, "vaw_sampler_sample: mode must be 0 or 1");
  VAW_CHECK_ARG(T > 0 && T <= kMaxT, "vaw_sampler_sample: T=%d out of range (1..%d)", T, kMaxT);
  VAW_CHECK_ARG(mode == 0 ? w_in != nullptr : (history != nullptr && counts != nullptr && H > 0),
                "vaw_sampler_sample: missing weights/history");
  VAW_CHECK_ARG(B >= 0 && (B == 0 || (u && idx && imp_w)), "vaw_sampler_sample: bad batch arguments");
  sampler_sample_kernel<<<1, 1024, 0, stream>>>(mode, w_in, history, counts, T, H, uniform_prob, u, B, idx, imp_w,
                                                w_out, p_out, cdf_out);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

extern "C" int vaw_sampler_update(double* history, int* counts, const int* ts, const float* losses, long long n_total,
                                  int T, int H, cudaStream_t stream) {
  VAW_CHECK_ARG(history && counts && T > 0 && H > 0, "vaw_sampler_update: bad arguments");
  if (n_total == 0) return VAW_OK;
  VAW_CHECK_ARG(ts && losses, "vaw_sampler_update: null entries");
  sampler_update_kernel<<<(T + 255) / 256, 256, 0, stream>>>(history, counts, ts, losses, n_total, T, H);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

extern "C" int vaw_pack_tloss(const long long* t, const float* loss, int* t32, float* l32, long long B, long long Bpad,
                              cudaStream_t stream) {
  VAW_CHECK_ARG(t32 && l32 && Bpad >= B && B >= 0, "vaw_pack_tloss: bad arguments");
  if (Bpad == 0) return VAW_OK;
  VAW_CHECK_ARG(B == 0 || (t && loss), "vaw_pack_tloss: null inputs");
  pack_tloss_kernel<<<(unsigned)((Bpad + 255) / 256), 256, 0, stream>>>(t, loss, t32, l32, B, Bpad);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

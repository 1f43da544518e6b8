// attention_sm100.cu — tcgen05 / TMEM self-attention forward for sequences of up to 256 tokens (every DiT config:
// T = 256, head_dim 64 or 72).  Replaces F.scaled_dot_product_attention inside timm Attention (models/dit.py:126).
//
// One CTA = one (batch, head, 128-query tile); two CTAs are resident per SM (100 KB of shared memory and 256 TMEM
// columns each), so one CTA's softmax overlaps the other's loads and MMAs without any intra-CTA pipeline.
//   1. TMA (4-D tensor maps over the packed qkv activation [B, T, 3*H, hd], out-of-bounds zero fill) brings the Q tile
//      and the whole K and V of the head into shared memory.  head_dim 72 is split 64 + 16: a SWIZZLE_128B box for the
//      first 64 columns and a SWIZZLE_32B box for columns 64..79, of which 72..79 are zero-filled by the TMA unit.
//   2. S = Q K^T  [128 x Nk] fp32 in TMEM columns [0, Nk): 4 (+1) tcgen05.mma (K = 16 each), both operands K-major.
//   3. softmax: two threads per query row (TMEM lane = row), one per 128-key half; ONE pass over the half row with
//      tcgen05.ld (exp2 / sum relative to an estimated row maximum, see the kernel; a CTA that meets a row outside the
//      safe range recomputes S and runs the exact max-then-exp two-pass version); P is written back as packed bf16
//      pairs with tcgen05.st in place of the S columns the same thread has already consumed.
//   4. O = P V  [128 x hd]: A operand read from TMEM (P), B = V from shared memory (MN-major), accumulator in the
//      TMEM columns the softmax freed.
//   5. epilogue: O / rowsum -> bf16 -> out[b, t, h, :]; L2[q] = max*c + log2(rowsum) (log2-domain, see attention.cu).
//      (Direct 16-byte stores: a TMA-store epilogue was measured slower here - it keeps the CTA and its shared memory
//       alive until the store engine has drained the tile, and the next CTA of the SM waits for that.)
#include "vaw_common.cuh"
#include "vaw_tc5.cuh"
#include "vaw_internal.h"

namespace {

using namespace tc5;

// running maximum of one 32-column chunk of a score row (keys >= T are ignored)
__device__ __forceinline__ void row_max_chunk(const uint32_t (&v)[32], int key0, int T, float& m0, float& m1, float& m2,
                                              float& m3) {
  if (key0 + 32 <= T) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      m0 = fmaxf(m0, __uint_as_float(v[j]));
      m1 = fmaxf(m1, __uint_as_float(v[j + 1]));
      m2 = fmaxf(m2, __uint_as_float(v[j + 2]));
      m3 = fmaxf(m3, __uint_as_float(v[j + 3]));
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (key0 + j < T) m0 = fmaxf(m0, __uint_as_float(v[j]));
  }
}
// p = 2^(s*c - m*c) for one chunk: packed bf16 pairs for the P operand, fp32 partial row sums
__device__ __forceinline__ void softmax_chunk(const uint32_t (&v)[32], uint32_t (&pk)[16], int key0, int T, float c,
                                              float mneg, float& s0, float& s1, float& s2, float& s3) {
  if (key0 + 32 <= T) {
    const float2 c2 = make_float2(c, c), m2 = make_float2(mneg, mneg);
    float2 sa = make_float2(s0, s1), sb = make_float2(s2, s3);
#pragma unroll
    for (int j = 0; j < 32; j += 4) {   // packed FFMA2 / FADD2: two elements per issue slot around the MUFU.EX2
      const float2 xa = __ffma2_rn(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), c2, m2);
      const float2 xb = __ffma2_rn(make_float2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])), c2, m2);
      const float2 pa = make_float2(ex2_approx(xa.x), ex2_approx(xa.y));
      const float2 pb = make_float2(ex2_approx(xb.x), ex2_approx(xb.y));
      sa = __fadd2_rn(sa, pa);
      sb = __fadd2_rn(sb, pb);
      pk[j >> 1] = pack_bf16(pa.x, pa.y);
      pk[(j >> 1) + 1] = pack_bf16(pb.x, pb.y);
    }
    s0 = sa.x; s1 = sa.y; s2 = sb.x; s3 = sb.y;
  } else {
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const float p0 = key0 + j < T ? ex2_approx(fmaf(__uint_as_float(v[j]), c, mneg)) : 0.f;
      const float p1 = key0 + j + 1 < T ? ex2_approx(fmaf(__uint_as_float(v[j + 1]), c, mneg)) : 0.f;
      s0 += p0; s1 += p1;
      pk[j >> 1] = pack_bf16(p0, p1);
    }
  }
}

constexpr int kTile = 128;  // query rows per CTA, key rows per TMA box
constexpr int kMaxKeys = 256;

template <int HD>
struct Smem {
  static constexpr bool kTail = HD > 64;
  static constexpr int kQm = 0;                        // Q  main  [128 x 64]  SW128
  static constexpr int kKm = kQm + kTile * 128;        // K  main  [256 x 64]
  static constexpr int kVm = kKm + kMaxKeys * 128;     // V  main  [256 x 64]
  static constexpr int kQt = kVm + kMaxKeys * 128;     // Q  tail  [128 x 16]  SW32
  static constexpr int kKt = kQt + (kTail ? kTile * 32 : 0);
  static constexpr int kVt = kKt + (kTail ? kMaxKeys * 32 : 0);
  static constexpr int kBars = kVt + (kTail ? kMaxKeys * 32 : 0);
  static constexpr int kStat = kBars + 64;              // row max / row sum exchange between the two key halves
  static constexpr int kBytes = kStat + 4 * kTile * 4 + 1024;  // + alignment slack
};

template <int HD>
__global__ void __launch_bounds__(256, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_main, const __grid_constant__ CUtensorMap tm_tail,
                   bf16* __restrict__ o, float* __restrict__ lse2, int T, int Ta, int H, float scale_log2e, int exact) {
  using S = Smem<HD>;
  constexpr bool kTail = S::kTail;
  // TMEM columns after the softmax (S occupied [0, 256)): each key half rewrites its own S columns in place with P
  //   [0, 64) P keys 0..127 | [64, 80) O tail | [128, 192) P keys 128..255 | [192, 256) O main
  constexpr int kOCol = 192, kOTailCol = 64, kPHiCol = 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bar_qk = reinterpret_cast<uint64_t*>(smem + S::kBars);
  uint64_t* bar_v = bar_qk + 1;
  uint64_t* bar_s = bar_qk + 2;
  uint64_t* bar_o = bar_qk + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_qk + 4);
  float* s_max = reinterpret_cast<float*>(smem + S::kStat);
  float* s_sum = s_max + 2 * kTile;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kTile, h = blockIdx.y, b = blockIdx.z;
  const int nk = (T + 15) & ~15;                 // key columns of S (multiple of 16, <= 256)
  const int kboxes = (T + kTile - 1) / kTile;    // 128-row TMA boxes holding keys

  if (threadIdx.x == 0) {
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    fence_mbar_init_cta();
    const uint32_t box_bytes = kTile * 128 + (kTail ? kTile * 32 : 0);
    mbar_expect_tx(bar_qk, box_bytes * (1 + kboxes));
    tma_load_4d(smem + S::kQm, &tm_main, bar_qk, 0, h, q0, b);
    if (kTail) tma_load_4d(smem + S::kQt, &tm_tail, bar_qk, 64, h, q0, b);
    for (int kb = 0; kb < kboxes; ++kb) {
      tma_load_4d(smem + S::kKm + kb * kTile * 128, &tm_main, bar_qk, 0, H + h, kb * kTile, b);
      if (kTail) tma_load_4d(smem + S::kKt + kb * kTile * 32, &tm_tail, bar_qk, 64, H + h, kb * kTile, b);
    }
    mbar_expect_tx(bar_v, box_bytes * kboxes);
    for (int kb = 0; kb < kboxes; ++kb) {
      tma_load_4d(smem + S::kVm + kb * kTile * 128, &tm_main, bar_v, 0, 2 * H + h, kb * kTile, b);
      if (kTail) tma_load_4d(smem + S::kVt + kb * kTile * 32, &tm_tail, bar_v, 64, 2 * H + h, kb * kTile, b);
    }
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  bool first_mma = true;

  // ---- softmax roles: two threads per query row (TMEM lane = threadIdx.x & 127), each owning one half of the keys ----
  const int lane_row = threadIdx.x & 127;        // query row inside the tile == TMEM lane
  const int half = threadIdx.x >> 7;             // 0: key chunks 0..3, 1: key chunks 4..7
  const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const int nchunks = (nk + 31) >> 5;
  const int c_lo = half * 4, c_hi = min(nchunks, half * 4 + 4);
  float mx = 0.f, sum = 0.f;

  // Reading S out of TMEM is the softmax's bottleneck (tcgen05.ld bandwidth), so the row is normally read ONCE: the
  // reference point of the exponentials is the maximum over the first 32-key chunk of each half instead of the row
  // maximum.  softmax is invariant to the reference point, and bf16 / fp32 have the exponent range for
  // 2^(s - ref) up to 2^kSafe; rows whose true maximum lies further above the estimate (never seen with real
  // activations, but inputs are arbitrary) make the CTA recompute S and take the exact two-pass path (attempt 1).
  constexpr float kSafe = 80.f;
  for (int attempt = exact ? 1 : 0;; ++attempt) {
    // ---- S = Q K^T ----
    if (threadIdx.x == 0) {
      if (first_mma) mbar_wait(bar_qk, 0);
      tc_fence_after();
      const uint32_t id = idesc_bf16(nk, 0);
      const uint32_t aq = smem_u32(smem + S::kQm), ak = smem_u32(smem + S::kKm);
#pragma unroll
      for (int k = 0; k < 4; ++k) tc_mma_ss(tmem, desc_sw128(aq + k * 32), desc_sw128(ak + k * 32), id, k ? 1u : 0u);
      if (kTail) tc_mma_ss(tmem, desc_sw32(smem_u32(smem + S::kQt)), desc_sw32(smem_u32(smem + S::kKt)), id, 1u);
      tc_commit(bar_s);
    }
    __syncwarp();
    mbar_wait(bar_s, first_mma ? 0 : 1);
    first_mma = false;
    __syncwarp();
    tc_fence_after();

    uint32_t va[32], vb[32], pk[16];
    float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
    if (c_lo < c_hi) tmem_ld32(trow + c_lo * 32, va);
    if (attempt == 0) {          // estimate: this half's first chunk, which stays in registers for the exp pass
      tmem_ld_wait();
      if (c_lo < c_hi) row_max_chunk(va, c_lo * 32, T, mx0, mx1, mx2, mx3);
    } else {                     // exact: a full pass for the maximum, then the first chunk again
      for (int c = c_lo; c < c_hi; c += 2) {     // chunk c in va, chunk c + 1 in vb; the next load overlaps the math
        tmem_ld_wait();
        if (c + 1 < c_hi) tmem_ld32(trow + (c + 1) * 32, vb);
        row_max_chunk(va, c * 32, T, mx0, mx1, mx2, mx3);
        if (c + 1 < c_hi) {
          tmem_ld_wait();
          if (c + 2 < c_hi) tmem_ld32(trow + (c + 2) * 32, va);
          row_max_chunk(vb, (c + 1) * 32, T, mx0, mx1, mx2, mx3);
        }
      }
      if (c_lo < c_hi) tmem_ld32(trow + c_lo * 32, va);
    }
    mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
    s_max[half * kTile + lane_row] = mx;
    __syncthreads();
    mx = fmaxf(s_max[lane_row], s_max[kTile + lane_row]);
    const float mneg = -mx * scale_log2e;
    float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
    mx0 = mx1 = mx2 = mx3 = -INFINITY;           // attempt 0: the true maximum of this thread's keys, for the range check
    for (int c = c_lo; c < c_hi; c += 2) {
      tmem_ld_wait();
      if (c + 1 < c_hi) tmem_ld32(trow + (c + 1) * 32, vb);
      if (attempt == 0) row_max_chunk(va, c * 32, T, mx0, mx1, mx2, mx3);
      softmax_chunk(va, pk, c * 32, T, scale_log2e, mneg, sum0, sum1, sum2, sum3);
      tmem_st16(trow + half * kPHiCol + (c - c_lo) * 16, pk);
      if (c + 1 < c_hi) {
        tmem_ld_wait();
        if (c + 2 < c_hi) tmem_ld32(trow + (c + 2) * 32, va);
        if (attempt == 0) row_max_chunk(vb, (c + 1) * 32, T, mx0, mx1, mx2, mx3);
        softmax_chunk(vb, pk, (c + 1) * 32, T, scale_log2e, mneg, sum0, sum1, sum2, sum3);
        tmem_st16(trow + half * kPHiCol + (c + 1 - c_lo) * 16, pk);
      }
    }
    s_sum[half * kTile + lane_row] = (sum0 + sum1) + (sum2 + sum3);
    // !(x <= kSafe) also catches NaN scores: they take the exact path like everything unusual
    const float top = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
    const bool bad = attempt == 0 && c_lo < c_hi && !(fmaf(top, scale_log2e, mneg) <= kSafe);
    tmem_st_wait();
    tc_fence_before();
    const int redo = __syncthreads_or(bad);
    sum = s_sum[lane_row] + s_sum[kTile + lane_row];
    if (!redo) break;
  }

  // ---- O = P V ----  (warp 0 walks the loop with warp-uniform descriptors; lane 0 issues)
  if (warp == 0) {
    mbar_wait(bar_v, 0);
    tc_fence_after();
    const uint32_t tm_u = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t idm = idesc_bf16(64, 1), idt = idesc_bf16(16, 1);
    const uint64_t dv = desc_sw128(smem_u32(smem + S::kVm)), dvt = desc_sw32(smem_u32(smem + S::kVt));
    const int ksteps = nk >> 4;
    for (int k = 0; k < ksteps; ++k) {
      const uint32_t pa = tm_u + (k < 8 ? k * 8 : kPHiCol + (k - 8) * 8);
      if (lane == 0) {
        tc_mma_ts(tm_u + kOCol, pa, dv + (uint32_t)(k * 128), idm, k ? 1u : 0u);
        if (kTail) tc_mma_ts(tm_u + kOTailCol, pa, dvt + (uint32_t)(k * 32), idt, k ? 1u : 0u);
      }
    }
    if (lane == 0) tc_commit(bar_o);
  }
  __syncwarp();
  mbar_wait(bar_o, 0);
  __syncwarp();
  tc_fence_after();

  // ---- epilogue: half 0 writes head columns [0, 32) (+ the tail), half 1 columns [32, 64) ----
  const int row = q0 + lane_row;
  const float inv = 1.f / sum;
  if (row < T && half == 0) lse2[((long long)b * H + h) * Ta + row] = mx * scale_log2e + log2f(sum);
  bf16* orow = o + (((long long)b * Ta + row) * H + h) * HD;
  {
    uint32_t v[32];
    tmem_ld32(trow + kOCol + half * 32, v);
    tmem_ld_wait();
    if (row < T) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 w;
        w.x = pack_bf16(__uint_as_float(v[j]) * inv, __uint_as_float(v[j + 1]) * inv);
        w.y = pack_bf16(__uint_as_float(v[j + 2]) * inv, __uint_as_float(v[j + 3]) * inv);
        w.z = pack_bf16(__uint_as_float(v[j + 4]) * inv, __uint_as_float(v[j + 5]) * inv);
        w.w = pack_bf16(__uint_as_float(v[j + 6]) * inv, __uint_as_float(v[j + 7]) * inv);
        *reinterpret_cast<uint4*>(orow + half * 32 + j) = w;
      }
    }
  }
  if (kTail && half == 0) {   // warp-uniform: a half is four whole warps
    uint32_t v[16];
    tmem_ld16(trow + kOTailCol, v);
    tmem_ld_wait();
    if (row < T) {
      uint4 w;
      w.x = pack_bf16(__uint_as_float(v[0]) * inv, __uint_as_float(v[1]) * inv);
      w.y = pack_bf16(__uint_as_float(v[2]) * inv, __uint_as_float(v[3]) * inv);
      w.z = pack_bf16(__uint_as_float(v[4]) * inv, __uint_as_float(v[5]) * inv);
      w.w = pack_bf16(__uint_as_float(v[6]) * inv, __uint_as_float(v[7]) * inv);
      *reinterpret_cast<uint4*>(orow + 64) = w;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// ---- host ------------------------------------------------------------------------------------------------------------
template <int HD>
int launch_fwd_tc(const void* qkv, void* o, float* lse2, int B, int T, int Ta, int H, cudaStream_t stream) {
  using S = Smem<HD>;
  CUtensorMap tm_main, tm_tail;
  int rc = make_head_map(&tm_main, qkv, B, T, Ta, 3 * H, HD, 64, kTile, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = make_head_map(&tm_tail, qkv, B, T, Ta, 3 * H, HD, 16, kTile, CU_TENSOR_MAP_SWIZZLE_32B);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    VAW_CUDA_TRY(cudaFuncSetAttribute(attn_fwd_tc_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kBytes));
    configured = true;
  }
  const float scale = 1.0f / sqrtf((float)HD);
  dim3 grid((T + kTile - 1) / kTile, H, B);
  // VAW_ATTN_EXACT=1: always the two-pass softmax (A/B timing of the single-pass path; results agree either way)
  static const int exact = getenv("VAW_ATTN_EXACT") && atoi(getenv("VAW_ATTN_EXACT")) != 0;
  attn_fwd_tc_kernel<HD><<<grid, 256, S::kBytes, stream>>>(tm_main, tm_tail, (bf16*)o, lse2, T, Ta, H,
                                                            scale * 1.4426950408889634f, exact);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

}  // namespace

// internal entry (vaw_internal.h): returns VAW_ERR_UNSUPPORTED when the shape is outside this kernel's range
int vaw_attn_fwd_sm100(const void* qkv, void* o, float* lse2, int B, int T, int H, int head_dim, cudaStream_t stream) {
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) != 0 || (reinterpret_cast<uintptr_t>(o) & 15) != 0) return VAW_ERR_UNSUPPORTED;
  if (head_dim != 64 && head_dim != 72) return VAW_ERR_UNSUPPORTED;
  if (T > kMaxKeys && !vaw_attn_border_supported(T)) return VAW_ERR_UNSUPPORTED;
  // T in (256, 264] (U-ViT: 258, the ViT teacher: 257): tensor cores on the leading 256 tokens of every sample, then
  // the L-shaped border strip on CUDA cores (attention_border.cu)
  const int Tc = T > kMaxKeys ? kMaxKeys : T;
  int rc = head_dim == 64 ? launch_fwd_tc<64>(qkv, o, lse2, B, Tc, T, H, stream)
                          : launch_fwd_tc<72>(qkv, o, lse2, B, Tc, T, H, stream);
  if (rc == VAW_OK && T > kMaxKeys) rc = vaw_attn_border_fwd(qkv, o, lse2, B, T, H, head_dim, stream);
  return rc;
}

// smallk_gemm.cu — weight gradient of a Linear whose input has one row per SAMPLE: dW[M, N] (+)= A^T B with
// A [K, M] and B [K, N] bf16 and K = batch size (the adaLN-Zero modulation of every DiT block, models/dit.py:118-124:
// mod = silu(c) W^T + b with c [B, D], so dW = dmod^T silu(c) contracts over B = 64 rows only).
//
// Such a product is all epilogue: 1 GFLOP for 32 MB of fp32 output at DiT-XL/2.  The persistent tcgen05 GEMM spent
// 46 us per block on it (one k-block of main loop, then its smem-transposing fp32 epilogue: 0.7 TB/s); this kernel keeps
// the [K, 128] operand panels in shared memory, runs warp-level mma.sync straight out of them (both operands are
// "k-row, feature-contiguous": ldmatrix.trans delivers the A and B fragments directly) and stores the accumulator
// fragments as they are - the store stream is the only cost.  HBM-bound: 4 B per output element.
#include "vaw_common.cuh"
#include "vaw_internal.h"

namespace {

constexpr int kTM = 128, kTN = 128, kKC = 64;   // CTA tile, K chunk staged per pass
constexpr int kLd = kTM + 8;                     // padded row (elements): conflict-free ldmatrix

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// stage rows [k0, k0 + kKC) x columns [c0, c0 + 128) of a [K, ld] bf16 matrix, zero outside
__device__ __forceinline__ void stage_panel(bf16 (*s)[kLd], const bf16* __restrict__ g, long long ld, int K, int k0,
                                            int c0, int cols) {
  for (int ch = threadIdx.x; ch < kKC * (kTM / 8); ch += blockDim.x) {
    const int r = ch / (kTM / 8), part = ch - r * (kTM / 8);
    const int k = k0 + r, c = c0 + part * 8;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (k < K && c + 8 <= cols) {
      v = __ldg(reinterpret_cast<const uint4*>(g + (long long)k * ld + c));
    } else if (k < K && c < cols) {   // ragged right edge
      bf16 t[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) t[e] = c + e < cols ? g[(long long)k * ld + c + e] : __float2bfloat16(0.f);
      v = *reinterpret_cast<uint4*>(t);
    }
    *reinterpret_cast<uint4*>(&s[r][part * 8]) = v;
  }
}

__global__ void __launch_bounds__(256)
wgrad_smallk_kernel(const bf16* __restrict__ A, long long lda, const bf16* __restrict__ Bm, long long ldb,
                    float* __restrict__ out, long long ldo, int M, int N, int K, int accumulate) {
  __shared__ __align__(16) bf16 sA[kKC][kLd];
  __shared__ __align__(16) bf16 sB[kKC][kLd];
  const int m0 = blockIdx.y * kTM, n0 = blockIdx.x * kTN;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 64;   // 4 x 2 warps, 32 x 64 outputs each
  float acc[2][8][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;
  const int mat = lane >> 3, r8 = lane & 7;
  for (int k0 = 0; k0 < K; k0 += kKC) {
    if (k0) __syncthreads();
    stage_panel(sA, A, lda, K, k0, m0, M);
    stage_panel(sB, Bm, ldb, K, k0, n0, N);
    __syncthreads();
    const int kc = min(kKC, (K - k0 + 15) & ~15);
    for (int kk = 0; kk < kc; kk += 16) {
      uint32_t a[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i)   // matrices: (k lo, m lo) (k lo, m hi) (k hi, m lo) (k hi, m hi) -> a0a1 a2a3 a4a5 a6a7
        ldsm_x4_t(a[i], smem_addr(&sA[kk + r8 + (mat >> 1) * 8][wm + i * 16 + (mat & 1) * 8]));
#pragma unroll
      for (int j = 0; j < 4; ++j) {  // two n8 tiles per ldmatrix: (k lo, n0) (k hi, n0) (k lo, n0 + 8) (k hi, n0 + 8)
        uint32_t b[4];
        ldsm_x4_t(b, smem_addr(&sB[kk + r8 + (mat & 1) * 8][wn + j * 16 + (mat >> 1) * 8]));
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          mma16816(acc[i][2 * j], a[i], b[0], b[1]);
          mma16816(acc[i][2 * j + 1], a[i], b[2], b[3]);
        }
      }
    }
  }
  // accumulator fragment: c0 c1 at (row = lane / 4, col = 2 (lane % 4)), c2 c3 eight rows below
#pragma unroll
  for (int i = 0; i < 2; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = n0 + wn + j * 8 + (lane & 3) * 2;
#pragma unroll
      for (int hrow = 0; hrow < 2; ++hrow) {
        const int row = m0 + wm + i * 16 + (lane >> 2) + hrow * 8;
        if (row >= M || col >= N) continue;
        float* p = out + (long long)row * ldo + col;
        float2 v = make_float2(acc[i][j][2 * hrow], acc[i][j][2 * hrow + 1]);
        if (col + 1 < N) {
          if (accumulate) {
            const float2 o = *reinterpret_cast<const float2*>(p);
            v.x += o.x;
            v.y += o.y;
          }
          *reinterpret_cast<float2*>(p) = v;
        } else {
          p[0] = accumulate ? p[0] + v.x : v.x;
        }
      }
    }
  }
}

}  // namespace

// out[M, N] (ldo, fp32) (+)= A^T B; A [K, M] (lda), B [K, N] (ldb) bf16 with the feature index contiguous.  Any K; meant
// for K = batch size (<= a few hundred).  Needs 16-byte aligned operands with lda, ldb multiples of 8 and an 8-byte
// aligned output with even ldo; returns VAW_ERR_UNSUPPORTED otherwise (callers fall back to vaw_gemm_bf16).
extern "C" int vaw_wgrad_smallk(const void* A, long long lda, const void* B, long long ldb, float* out, long long ldo, int M, int N,
                     int K, int accumulate, cudaStream_t stream) {
  if (!A || !B || !out || M <= 0 || N <= 0 || K <= 0) return VAW_ERR_UNSUPPORTED;
  const uintptr_t al = reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B);
  if ((al & 15) || (lda & 7) || (ldb & 7) || (reinterpret_cast<uintptr_t>(out) & 7) || (ldo & 1)) return VAW_ERR_UNSUPPORTED;
  dim3 grid((N + kTN - 1) / kTN, (M + kTM - 1) / kTM);
  wgrad_smallk_kernel<<<grid, 256, 0, stream>>>((const bf16*)A, lda, (const bf16*)B, ldb, out, ldo, M, N, K, accumulate);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

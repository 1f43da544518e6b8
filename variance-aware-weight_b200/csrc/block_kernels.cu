// block_kernels.cu — the memory-bound pieces of the DiT / U-ViT block that cannot live in a GEMM epilogue:
//   LayerNorm (+ adaLN modulate, or affine) forward / backward          (models/dit.py:24-25,122-124,133-137;
//   gate * branch backward (adaLN-Zero) and residual bookkeeping          models/uvit.py:96-121)
//   deterministic column reductions for bias / shift / scale / gate gradients
// All kernels are HBM-bound: one pass over the [rows, D] activation from DRAM, 128-bit accesses, fp32 math.
// The backward kernels give every thread a fixed group of 4 columns and let it walk the rows of its chunk, so the
// per-column sums (d shift / d scale / d gate / bias gradients) live in registers: no shared-memory accumulators,
// no atomics, high occupancy, and a fixed summation order (bit-reproducible gradients).  Per-CTA partials are folded
// by the finish kernels in fixed order.
#include "vaw_common.cuh"
#include "vaw_async.cuh"

namespace {

constexpr int kMaxD = 2048;  // ln_fwd keeps a row in registers: KV = ceil(D / 128) float4 per lane, KV in {3, 6, 9, 16}

// ---------------------------------------------------------------------------------------------------
// LayerNorm forward: y = xhat * A + Bv, xhat = (x - mean) * rstd
//   modulate (DiT):  A = 1 + scale[n, :], Bv = shift[n, :]     (n = row / rows_per_sample; ld_mod = row stride)
//   affine (U-ViT):  A = weight[:],       Bv = bias[:]
//   plain:           A = 1, Bv = 0
// one warp per row at a time, 4 rows per warp, the row lives in registers (two-pass variance like torch).
// ---------------------------------------------------------------------------------------------------
constexpr int kLnRowsPerWarp = 4;

// RES: the row that is normalised is the residual update of the PREVIOUS branch, x_out = x + gate[n, :] * branch, which
// is formed here (and written out for the backward pass / the next residual) instead of in the epilogue of the GEMM
// that produced `branch`: that GEMM becomes a plain bf16-output GEMM and the fp32 residual stream is read once.
struct LnRes {
  const bf16* branch;   // [M, D] output of the previous branch (attention projection / MLP), bias included
  const float* gate;    // [n, :] adaLN-Zero gate of that branch, row stride ld_gate
  long long ld_gate;
  float* x_out;         // [M, D] updated residual stream
};
// Output row layout: row stride ldy >= D; ones_block: the 32 bf16 after the D outputs of every row are set to
// [1, 0, ..., 0] (ldy >= D + 32).  A weight-gradient GEMM that reads y as [M, D + 32] then delivers the bias gradient
// (the row sums of the other operand) as output column D - no separate column-sum pass over that operand.
struct LnOut {
  long long ldy;
  int ones_block;
};

// EXACT: D == KV * 128, every lane owns KV float4 of the row - no bounds predicates.
// RPW rows of a warp are in flight together (their loads are issued before anything is reduced, their reductions
// interleave): a row of D = 384 is 12 registers per lane, and with one such row at a time the kernel ran at 1.3 TB/s -
// one DRAM round trip, two shuffle reductions and the modulation loads per row, back to back, 26 us for ANY D.
// The kernel is issue-limited at D = 1152 (ncu: 578 warp instructions per row, issue slots 43 % busy with 26 warps per SM),
// so the arithmetic is packed two floats per instruction (FADD2 / FMUL2 / FFMA2) and xhat is one fma:
// (v - mean) * rstd = fma(v, rstd, -mean * rstd).
template <int KV, bool RES, bool EXACT, int RPW>
__global__ void __launch_bounds__(256, RPW > 1 ? 2 : (KV <= 9 ? 4 : 2))   // RPW rows per warp in registers: fewer, fatter CTAs
ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ shift, const float* __restrict__ scale,
              long long ld_mod, int rows_per_sample, const float* __restrict__ weight,
              const float* __restrict__ bias, bf16* __restrict__ y, float* __restrict__ mean_out,
              float* __restrict__ rstd_out, int M, int D, float eps, LnRes res, LnOut lo) {
  static_assert(kLnRowsPerWarp % RPW == 0, "rows per warp");
  const int lane = threadIdx.x & 31;
  const int nv = D >> 2;  // float4 per row
  const int row_base = (blockIdx.x * 8 + (threadIdx.x >> 5)) * kLnRowsPerWarp;
  const float2 one2 = make_float2(1.f, 1.f);
#pragma unroll 1
  for (int rr = 0; rr < kLnRowsPerWarp; rr += RPW) {
    const int row0 = row_base + rr;
    if (row0 >= M) return;
    float4 v[RPW][KV];
    float2 s2[RPW];
#pragma unroll
    for (int u = 0; u < RPW; ++u) {
      const int row = row0 + u;
      s2[u] = make_float2(0.f, 0.f);
      if (row < M) {
        const float4* xr = reinterpret_cast<const float4*>(x + (long long)row * D);
#pragma unroll
        for (int i = 0; i < KV; ++i) {
          const int idx = i * 32 + lane;
          if (EXACT || idx < nv) v[u][i] = ldg_stream_f4(xr + idx);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < RPW; ++u) {
      const int row = row0 + u;
      if (row < M) {
#pragma unroll
        for (int i = 0; i < KV; ++i) {
          const int idx = i * 32 + lane;
          if (EXACT || idx < nv) {
            if constexpr (RES) {
              const uint2 br = ldg_stream_u2(reinterpret_cast<const uint2*>(res.branch + (long long)row * D) + idx);
              const float4 g = __ldg(reinterpret_cast<const float4*>(res.gate + (long long)(row / rows_per_sample) * res.ld_gate) + idx);
              const float2 lo2 = __ffma2_rn(make_float2(g.x, g.y), unpack_bf16(br.x), make_float2(v[u][i].x, v[u][i].y));
              const float2 hi2 = __ffma2_rn(make_float2(g.z, g.w), unpack_bf16(br.y), make_float2(v[u][i].z, v[u][i].w));
              v[u][i] = make_float4(lo2.x, lo2.y, hi2.x, hi2.y);
              stg_stream_f4(reinterpret_cast<float4*>(res.x_out + (long long)row * D) + idx, v[u][i]);
            }
            s2[u] = __fadd2_rn(s2[u], __fadd2_rn(make_float2(v[u][i].x, v[u][i].y), make_float2(v[u][i].z, v[u][i].w)));
          }
        }
      }
    }
    float mean[RPW], rstd[RPW];
#pragma unroll
    for (int u = 0; u < RPW; ++u) mean[u] = warp_sum(s2[u].x + s2[u].y) / (float)D;
#pragma unroll
    for (int u = 0; u < RPW; ++u) {
      const float2 nm = make_float2(-mean[u], -mean[u]);
      float2 q2 = make_float2(0.f, 0.f);
      if (row0 + u < M) {
#pragma unroll
        for (int i = 0; i < KV; ++i) {
          const int idx = i * 32 + lane;
          if (EXACT || idx < nv) {
            const float2 a = __fadd2_rn(make_float2(v[u][i].x, v[u][i].y), nm), b = __fadd2_rn(make_float2(v[u][i].z, v[u][i].w), nm);
            q2 = __ffma2_rn(a, a, q2);
            q2 = __ffma2_rn(b, b, q2);
          }
        }
      }
      s2[u] = q2;
    }
#pragma unroll
    for (int u = 0; u < RPW; ++u) rstd[u] = rsqrtf(warp_sum(s2[u].x + s2[u].y) / (float)D + eps);
#pragma unroll
    for (int u = 0; u < RPW; ++u) {
      const int row = row0 + u;
      if (row >= M) continue;
      if (lane == 0) {
        mean_out[row] = mean[u];
        rstd_out[row] = rstd[u];
      }
      const long long mo = scale ? (long long)(row / rows_per_sample) * ld_mod : 0;
      uint2* yr = reinterpret_cast<uint2*>(y + (long long)row * lo.ldy);
      if (lo.ones_block && lane < 4)   // bf16 1.0 = 0x3F80 in the first of 32 elements
        reinterpret_cast<uint4*>(y + (long long)row * lo.ldy + D)[lane] = make_uint4(lane == 0 ? 0x00003F80u : 0u, 0u, 0u, 0u);
      const float2 r2 = make_float2(rstd[u], rstd[u]), nmr = make_float2(-mean[u] * rstd[u], -mean[u] * rstd[u]);
#pragma unroll
      for (int i = 0; i < KV; ++i) {
        const int idx = i * 32 + lane;
        if (EXACT || idx < nv) {
          float2 A0 = one2, A1 = one2, B0 = make_float2(0.f, 0.f), B1 = B0;
          if (scale) {
            const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + mo) + idx);
            const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + mo) + idx);
            A0 = __fadd2_rn(make_float2(sc.x, sc.y), one2);
            A1 = __fadd2_rn(make_float2(sc.z, sc.w), one2);
            B0 = make_float2(sh.x, sh.y);
            B1 = make_float2(sh.z, sh.w);
          } else if (weight) {
            const float4 wv = __ldg(reinterpret_cast<const float4*>(weight) + idx);
            const float4 bv = __ldg(reinterpret_cast<const float4*>(bias) + idx);
            A0 = make_float2(wv.x, wv.y); A1 = make_float2(wv.z, wv.w);
            B0 = make_float2(bv.x, bv.y); B1 = make_float2(bv.z, bv.w);
          }
          const float2 o0 = __ffma2_rn(__ffma2_rn(make_float2(v[u][i].x, v[u][i].y), r2, nmr), A0, B0);
          const float2 o1 = __ffma2_rn(__ffma2_rn(make_float2(v[u][i].z, v[u][i].w), r2, nmr), A1, B1);
          uint2 o;
          o.x = pack_bf16(o0.x, o0.y);
          o.y = pack_bf16(o1.x, o1.y);
          yr[idx] = o;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// LayerNorm backward.  grid = (chunks, groups): group = sample (modulate) or arbitrary row range (affine).
//   g = dy * A ;  dx = rstd * (g - mean(g) - xhat * mean(g * xhat)) ;  dx_io = (add_into ? dx_io : 0) + dx
//   part[group, chunk, 0, :] = sum_rows dy         (d shift / d bias)
//   part[group, chunk, 1, :] = sum_rows dy * xhat  (d scale / d weight)
// phase 1: warps sweep the chunk's rows and leave (mean(g), mean(g*xhat)) per row in shared memory
// phase 2: every thread owns 4 columns and walks the rows (x / dy come back from L1/L2), column sums in registers
// ---------------------------------------------------------------------------------------------------
constexpr int kBwdMaxRows = 64;  // rows per chunk handled by one CTA (shared-memory row statistics)

__global__ void __launch_bounds__(512)
ln_bwd_kernel(const bf16* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean_in,
              const float* __restrict__ rstd_in, const float* __restrict__ scale, long long ld_mod,
              const float* __restrict__ weight, float* __restrict__ dx_io, int add_into, float* __restrict__ part,
              int rows_per_group, int chunks, int M, int D) {
  __shared__ float s_m1[kBwdMaxRows], s_m2[kBwdMaxRows], s_mean[kBwdMaxRows], s_rstd[kBwdMaxRows];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int chunk = blockIdx.x, group = blockIdx.y;
  const int nv = D >> 2;
  const int rows_per_chunk = (rows_per_group + chunks - 1) / chunks;
  const int r_begin = group * rows_per_group + chunk * rows_per_chunk;
  const int r_end = min(min(r_begin + rows_per_chunk, (group + 1) * rows_per_group), M);
  const int nrows = max(r_end - r_begin, 0);
  const float4* Ap = scale ? reinterpret_cast<const float4*>(scale + (long long)group * ld_mod)
                           : reinterpret_cast<const float4*>(weight);
  const float add1 = scale ? 1.f : 0.f;  // A = 1 + scale (modulate) | weight (affine) | 1 (plain)
  const float inv_d = 1.f / (float)D;

  // ---- phase 1: row statistics ----
  for (int r = warp; r < nrows; r += nwarps) {
    const int row = r_begin + r;
    const float4* xr = reinterpret_cast<const float4*>(x + (long long)row * D);
    const uint2* dyr = reinterpret_cast<const uint2*>(dy + (long long)row * D);
    const float mean = mean_in[row], rstd = rstd_in[row];
    float s1 = 0.f, s2 = 0.f;
    for (int idx = lane; idx < nv; idx += 32) {
      const float4 xv = xr[idx];
      const uint2 du = dyr[idx];
      const float2 d0 = unpack_bf16(du.x), d1 = unpack_bf16(du.y);
      float4 A = make_float4(1.f, 1.f, 1.f, 1.f);
      if (Ap) {
        const float4 a = __ldg(Ap + idx);
        A = make_float4(a.x + add1, a.y + add1, a.z + add1, a.w + add1);
      }
      const float g0 = d0.x * A.x, g1 = d0.y * A.y, g2 = d1.x * A.z, g3 = d1.y * A.w;
      s1 += (g0 + g1) + (g2 + g3);
      s2 += (g0 * ((xv.x - mean) * rstd) + g1 * ((xv.y - mean) * rstd)) +
            (g2 * ((xv.z - mean) * rstd) + g3 * ((xv.w - mean) * rstd));
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
      s_m1[r] = s1 * inv_d;
      s_m2[r] = s2 * inv_d;
      s_mean[r] = mean;
      s_rstd[r] = rstd;
    }
  }
  __syncthreads();

  // ---- phase 2: column owners ----
  for (int cg = threadIdx.x; cg < nv; cg += blockDim.x) {
    float4 A = make_float4(1.f, 1.f, 1.f, 1.f);
    if (Ap) {
      const float4 a = __ldg(Ap + cg);
      A = make_float4(a.x + add1, a.y + add1, a.z + add1, a.w + add1);
    }
    float4 accB = make_float4(0.f, 0.f, 0.f, 0.f), accA = accB;
#pragma unroll 4
    for (int r = 0; r < nrows; ++r) {
      const long long o = (long long)(r_begin + r) * D;
      const float4 xv = *(reinterpret_cast<const float4*>(x + o) + cg);
      const uint2 du = *(reinterpret_cast<const uint2*>(dy + o) + cg);
      const float2 d0 = unpack_bf16(du.x), d1 = unpack_bf16(du.y);
      const float mean = s_mean[r], rstd = s_rstd[r], m1 = s_m1[r], m2 = s_m2[r];
      const float h0 = (xv.x - mean) * rstd, h1 = (xv.y - mean) * rstd, h2 = (xv.z - mean) * rstd,
                  h3 = (xv.w - mean) * rstd;
      float4 out;
      out.x = rstd * (d0.x * A.x - m1 - h0 * m2);
      out.y = rstd * (d0.y * A.y - m1 - h1 * m2);
      out.z = rstd * (d1.x * A.z - m1 - h2 * m2);
      out.w = rstd * (d1.y * A.w - m1 - h3 * m2);
      float4* dxp = reinterpret_cast<float4*>(dx_io + o) + cg;
      if (add_into) {
        const float4 p = *dxp;
        out.x += p.x; out.y += p.y; out.z += p.z; out.w += p.w;
      }
      *dxp = out;
      accB.x += d0.x; accB.y += d0.y; accB.z += d1.x; accB.w += d1.y;
      accA.x += d0.x * h0; accA.y += d0.y * h1; accA.z += d1.x * h2; accA.w += d1.y * h3;
    }
    if (part) {
      float* dst = part + ((long long)group * chunks + chunk) * 2 * D;
      *(reinterpret_cast<float4*>(dst) + cg) = accB;
      *(reinterpret_cast<float4*>(dst + D) + cg) = accA;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// LayerNorm backward, shared-memory staged (the production path; the kernel above is the fallback for D % 8 != 0).
// A CTA walks its chunk in sub-chunks of R rows.  The rows of a sub-chunk are contiguous in memory, so x (fp32) and
// dy (bf16) arrive with two 1-D bulk copies (TMA engine, mbarrier completion) and are read twice from shared memory
// instead of twice from L2/DRAM (the two-pass version re-fetched ~45 % of its bytes from DRAM, ncu).  Two CTAs per
// SM: one computes while the other's copies are in flight.
// Threads form `halves` x nvp column owners: thread (h, cg) owns float4 column cg for rows h, h+halves, ...; column
// sums stay in registers across sub-chunks and are folded over h in fixed order at the end (deterministic).
// ---------------------------------------------------------------------------------------------------
constexpr int kStagedBudget = 110592;  // bytes of staged rows per CTA (2 stages of x fp32 + dy bf16): 2 x 8 rows at D = 1152

// FUSE: the residual-branch backward of the NEXT branch rides along (vaw_ln_bwd_gate): with dx' the updated residual
// gradient, dy_next = bf16(dx' * gate_next), part_gate = (sum dx', sum dx' * y_next) - exactly gate_bwd_kernel's outputs,
// without re-reading dx' (75 MB per call at DiT-XL/2 B=64).  That variant runs fewer threads with more registers.
struct GateFuse {
  const bf16* y;       // [M, D] output of the next branch (saved in forward), null: plain residual
  const float* gate;   // [groups, ld_gate] gate rows, null: plain residual
  long long ld_gate;
  bf16* dy;            // [M, D] out
  float* part;         // [groups, chunks, 2, D] out
};
template <int MAXT, bool FUSE>
__global__ void __launch_bounds__(MAXT, 2)
ln_bwd_staged_kernel(const bf16* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean_in,
                     const float* __restrict__ rstd_in, const float* __restrict__ scale, long long ld_mod,
                     const float* __restrict__ weight, float* __restrict__ dx_io, int add_into,
                     float* __restrict__ part, int rows_per_group, int chunks, int M, int D, int R, int nvp,
                     GateFuse gf) {
  extern __shared__ __align__(128) uint8_t ln_smem[];
  // stage s: x [R, D] fp32 | dy [R, D] bf16 ; then per stage [6, R] floats (two half-row sums of g and g*xhat, mean,
  // rstd) ; then 2 mbarriers
  const size_t stage_bytes = (size_t)R * D * 6;
  float* s_stat_all = reinterpret_cast<float*>(ln_smem + 2 * stage_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ln_smem + 2 * stage_bytes + (size_t)R * 48);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int halves = blockDim.x / nvp;
  const int half = threadIdx.x / nvp, cg = threadIdx.x % nvp;
  const int chunk = blockIdx.x, group = blockIdx.y;
  const int nv = D >> 2;
  const bool owner = cg < nv;
  const int rows_per_chunk = (rows_per_group + chunks - 1) / chunks;
  const int c_begin = group * rows_per_group + chunk * rows_per_chunk;
  const int c_end = min(min(c_begin + rows_per_chunk, (group + 1) * rows_per_group), M);
  const int nsub = c_end > c_begin ? (c_end - c_begin + R - 1) / R : 0;
  const float4* Ap = scale ? reinterpret_cast<const float4*>(scale + (long long)group * ld_mod)
                           : reinterpret_cast<const float4*>(weight);
  const float add1 = scale ? 1.f : 0.f;
  const float inv_d = 1.f / (float)D;

  auto issue = [&](int sub) {   // thread 0: bulk copies of sub-chunk `sub` into stage sub & 1
    const int r0 = c_begin + sub * R;
    const int nrows = min(R, c_end - r0);
    uint8_t* st = ln_smem + (size_t)(sub & 1) * stage_bytes;
    const uint32_t bx = (uint32_t)nrows * D * 4, bd = (uint32_t)nrows * D * 2;
    mbar_expect_tx(&bars[sub & 1], bx + bd);
    bulk_g2s(st, x + (long long)r0 * D, bx, &bars[sub & 1]);
    bulk_g2s(st + (size_t)R * D * 4, dy + (long long)r0 * D, bd, &bars[sub & 1]);
  };

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init_cta();
    if (nsub > 0) issue(0);
    if (nsub > 1) issue(1);
  }
  __syncthreads();

  float4 A = make_float4(1.f, 1.f, 1.f, 1.f);
  if (Ap && owner) {
    const float4 a = __ldg(Ap + cg);
    A = make_float4(a.x + add1, a.y + add1, a.z + add1, a.w + add1);
  }
  float4 accB = make_float4(0.f, 0.f, 0.f, 0.f), accA = accB;
  float4 accS = accB, accG = accB, gt = make_float4(1.f, 1.f, 1.f, 1.f);
  if (FUSE && gf.gate && owner) gt = __ldg(reinterpret_cast<const float4*>(gf.gate + (long long)group * gf.ld_gate) + cg);

  for (int sub = 0; sub < nsub; ++sub) {
    const int r0 = c_begin + sub * R;
    const int nrows = min(R, c_end - r0);
    const int stg = sub & 1;
    const float* sx = reinterpret_cast<const float*>(ln_smem + (size_t)stg * stage_bytes);
    const bf16* sdy = reinterpret_cast<const bf16*>(ln_smem + (size_t)stg * stage_bytes + (size_t)R * D * 4);
    float* s_h1 = s_stat_all + stg * 6 * R, *s_h2 = s_h1 + 2 * R, *s_mean = s_h1 + 4 * R, *s_rstd = s_h1 + 5 * R;
    if (threadIdx.x < nrows) {   // row statistics of the forward pass ride along
      s_mean[threadIdx.x] = mean_in[r0 + threadIdx.x];
      s_rstd[threadIdx.x] = rstd_in[r0 + threadIdx.x];
    }
    mbar_wait(&bars[stg], (uint32_t)(sub >> 1) & 1u);
    __syncthreads();

    // ---- row statistics: sum(g), sum(g * xhat) with g = dy * A; two warps per row (each half of the columns), so that
    //      16 of the 18 warps work; the two halves meet in shared memory in fixed order ----
    {
      const int nv_half = (nv + 1) >> 1;
      for (int job = warp; job < 2 * nrows; job += nwarps) {
        const int r = job >> 1, hf = job & 1;
        const float4* xr = reinterpret_cast<const float4*>(sx + (size_t)r * D);
        const uint2* dyr = reinterpret_cast<const uint2*>(sdy + (size_t)r * D);
        const float rstd = s_rstd[r], nmr = -s_mean[r] * rstd;   // xhat = fma(x, rstd, -mean * rstd)
        const float2 rs2 = make_float2(rstd, rstd), nm2 = make_float2(nmr, nmr);
        float2 s1v = make_float2(0.f, 0.f), s2v = s1v;
        float s1, s2;
        const int i_end = min(nv, (hf + 1) * nv_half);
        for (int idx = hf * nv_half + lane; idx < i_end; idx += 32) {
          const float4 xv = xr[idx];
          const uint2 du = dyr[idx];
          const float2 d0 = unpack_bf16(du.x), d1 = unpack_bf16(du.y);
          float4 Aw = make_float4(1.f, 1.f, 1.f, 1.f);
          if (Ap) {
            const float4 a = __ldg(Ap + idx);
            Aw = make_float4(a.x + add1, a.y + add1, a.z + add1, a.w + add1);
          }
          // packed fp32 (FFMA2 / FMUL2 / FADD2): the kernel is issue-bound, two elements per instruction
          const float2 ga = __fmul2_rn(d0, make_float2(Aw.x, Aw.y)), gb = __fmul2_rn(d1, make_float2(Aw.z, Aw.w));
          s1v = __fadd2_rn(s1v, __fadd2_rn(ga, gb));
          s2v = __ffma2_rn(ga, __ffma2_rn(make_float2(xv.x, xv.y), rs2, nm2), s2v);
          s2v = __ffma2_rn(gb, __ffma2_rn(make_float2(xv.z, xv.w), rs2, nm2), s2v);
        }
        s1 = warp_sum(s1v.x + s1v.y);
        s2 = warp_sum(s2v.x + s2v.y);
        if (lane == 0) {
          s_h1[hf * R + r] = s1;
          s_h2[hf * R + r] = s2;
        }
      }
    }
    __syncthreads();

    // ---- column owners: dx and the column sums, 4 rows per step (dx_io loads issued before the math).
    //      dx = rstd * (A dy - m1 - xhat m2) is evaluated as P dy + Q x + Rr with the per-row constants
    //      Q = -rstd^2 m2, Rr = -rstd m1 - Q mean and P = rstd A (per row and column): 2 FMA + 1 MUL per element ----
    if (owner) {
      for (int rb = half; rb < nrows; rb += 4 * halves) {
        float4 prev[4];
        uint2 yv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = rb + j * halves;
          prev[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          yv[j] = make_uint2(0u, 0u);
          if (add_into && r < nrows)
            prev[j] = ldg_stream_f4(reinterpret_cast<const float4*>(dx_io + (long long)(r0 + r) * D) + cg);
          if (FUSE && gf.y && r < nrows)
            yv[j] = ldg_stream_u2(reinterpret_cast<const uint2*>(gf.y + (long long)(r0 + r) * D) + cg);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = rb + j * halves;
          if (r < nrows) {
            const float4 xv = *(reinterpret_cast<const float4*>(sx + (size_t)r * D) + cg);
            const uint2 du = *(reinterpret_cast<const uint2*>(sdy + (size_t)r * D) + cg);
            const float2 d0 = unpack_bf16(du.x), d1 = unpack_bf16(du.y);
            const float mean = s_mean[r], rstd = s_rstd[r];
            const float m1 = (s_h1[r] + s_h1[R + r]) * inv_d, m2 = (s_h2[r] + s_h2[R + r]) * inv_d;
            const float nmr = -mean * rstd;
            const float Q = -rstd * rstd * m2, Rr = -rstd * m1 - Q * mean;
            const float2 rs2 = make_float2(rstd, rstd), nm2 = make_float2(nmr, nmr), Q2 = make_float2(Q, Q),
                         R2 = make_float2(Rr, Rr);
            const float2 xa = make_float2(xv.x, xv.y), xb = make_float2(xv.z, xv.w);
            const float2 oa = __ffma2_rn(__fmul2_rn(rs2, make_float2(A.x, A.y)), d0,
                                         __ffma2_rn(Q2, xa, __fadd2_rn(R2, make_float2(prev[j].x, prev[j].y))));
            const float2 ob = __ffma2_rn(__fmul2_rn(rs2, make_float2(A.z, A.w)), d1,
                                         __ffma2_rn(Q2, xb, __fadd2_rn(R2, make_float2(prev[j].z, prev[j].w))));
            const float4 out = make_float4(oa.x, oa.y, ob.x, ob.y);
            *(reinterpret_cast<float4*>(dx_io + (long long)(r0 + r) * D) + cg) = out;
            {
              const float2 ba = __fadd2_rn(make_float2(accB.x, accB.y), d0), bb = __fadd2_rn(make_float2(accB.z, accB.w), d1);
              accB = make_float4(ba.x, ba.y, bb.x, bb.y);
              const float2 aa = __ffma2_rn(d0, __ffma2_rn(xa, rs2, nm2), make_float2(accA.x, accA.y));
              const float2 ab = __ffma2_rn(d1, __ffma2_rn(xb, rs2, nm2), make_float2(accA.z, accA.w));
              accA = make_float4(aa.x, aa.y, ab.x, ab.y);
            }
            if (FUSE) {
              const float2 ga = __fmul2_rn(oa, make_float2(gt.x, gt.y)), gb = __fmul2_rn(ob, make_float2(gt.z, gt.w));
              uint2 ov;
              ov.x = pack_bf16(ga.x, ga.y);
              ov.y = pack_bf16(gb.x, gb.y);
              *(reinterpret_cast<uint2*>(gf.dy + (long long)(r0 + r) * D) + cg) = ov;
              const float2 y0 = unpack_bf16(yv[j].x), y1 = unpack_bf16(yv[j].y);
              const float2 sa = __fadd2_rn(make_float2(accS.x, accS.y), oa), sb = __fadd2_rn(make_float2(accS.z, accS.w), ob);
              accS = make_float4(sa.x, sa.y, sb.x, sb.y);
              const float2 qa = __ffma2_rn(oa, y0, make_float2(accG.x, accG.y));
              const float2 qb = __ffma2_rn(ob, y1, make_float2(accG.z, accG.w));
              accG = make_float4(qa.x, qa.y, qb.x, qb.y);
            }
          }
        }
      }
    }
    __syncthreads();   // every read of this stage is done: refill it with the sub-chunk after next
    if (threadIdx.x == 0 && sub + 2 < nsub) issue(sub + 2);
  }

  if (!part && !FUSE) return;
  // fold the halves in fixed order through shared memory (the staged rows are dead now)
  constexpr int NACC = FUSE ? 4 : 2;
  float4* red = reinterpret_cast<float4*>(ln_smem);   // [halves - 1, NACC, nvp] float4
  if (half > 0 && owner) {
    red[((half - 1) * NACC + 0) * nvp + cg] = accB;
    red[((half - 1) * NACC + 1) * nvp + cg] = accA;
    if (FUSE) {
      red[((half - 1) * NACC + 2) * nvp + cg] = accS;
      red[((half - 1) * NACC + 3) * nvp + cg] = accG;
    }
  }
  __syncthreads();
  if (half == 0 && owner) {
    for (int h = 1; h < halves; ++h) {
      const float4 b = red[((h - 1) * NACC + 0) * nvp + cg], a = red[((h - 1) * NACC + 1) * nvp + cg];
      accB.x += b.x; accB.y += b.y; accB.z += b.z; accB.w += b.w;
      accA.x += a.x; accA.y += a.y; accA.z += a.z; accA.w += a.w;
      if (FUSE) {
        const float4 c = red[((h - 1) * NACC + 2) * nvp + cg], g = red[((h - 1) * NACC + 3) * nvp + cg];
        accS.x += c.x; accS.y += c.y; accS.z += c.z; accS.w += c.w;
        accG.x += g.x; accG.y += g.y; accG.z += g.z; accG.w += g.w;
      }
    }
    if (part) {
      float* dst = part + ((long long)group * chunks + chunk) * 2 * D;
      *(reinterpret_cast<float4*>(dst) + cg) = accB;
      *(reinterpret_cast<float4*>(dst + D) + cg) = accA;
    }
    if (FUSE) {
      float* dst = gf.part + ((long long)group * chunks + chunk) * 2 * D;
      *(reinterpret_cast<float4*>(dst) + cg) = accS;
      *(reinterpret_cast<float4*>(dst + D) + cg) = accG;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Residual-branch backward: dx (fp32 grad of the block output x_out = x_in + gate * y) ->
//   dy[row, :]   = bf16(dx * gate[n, :])           (operand of the branch's dgrad / wgrad GEMMs)
//   part[n, chunk, 0, :] = sum_rows dx             (-> bias gradient: db = sum_n gate[n] * s[n])
//   part[n, chunk, 1, :] = sum_rows dx * y         (-> dgate[n])
// gate == null: plain residual (U-ViT): dy = bf16(dx), only the column sum is produced.
// Column-owner threads, register accumulators (see the header comment).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
gate_bwd_kernel(const float* __restrict__ dx, const bf16* __restrict__ y, const float* __restrict__ gate,
                long long ld_gate, bf16* __restrict__ dy, float* __restrict__ part, int rows_per_group, int chunks,
                int M, int D) {
  const int chunk = blockIdx.x, group = blockIdx.y;
  const int nv = D >> 2;
  const int rows_per_chunk = (rows_per_group + chunks - 1) / chunks;
  const int r_begin = group * rows_per_group + chunk * rows_per_chunk;
  const int r_end = min(min(r_begin + rows_per_chunk, (group + 1) * rows_per_group), M);
  const float4* gp = gate ? reinterpret_cast<const float4*>(gate + (long long)group * ld_gate) : nullptr;
  for (int cg = threadIdx.x; cg < nv; cg += blockDim.x) {
    float4 gt = make_float4(1.f, 1.f, 1.f, 1.f);
    if (gp) gt = __ldg(gp + cg);
    float4 accS = make_float4(0.f, 0.f, 0.f, 0.f), accG = accS;
#pragma unroll 4
    for (int row = r_begin; row < r_end; ++row) {
      const long long o = (long long)row * D;
      const float4 d = ldg_stream_f4(reinterpret_cast<const float4*>(dx + o) + cg);
      uint2 ov;
      ov.x = pack_bf16(d.x * gt.x, d.y * gt.y);
      ov.y = pack_bf16(d.z * gt.z, d.w * gt.w);
      *(reinterpret_cast<uint2*>(dy + o) + cg) = ov;
      accS.x += d.x; accS.y += d.y; accS.z += d.z; accS.w += d.w;
      if (y) {
        const uint2 yu = ldg_stream_u2(reinterpret_cast<const uint2*>(y + o) + cg);
        const float2 y0 = unpack_bf16(yu.x), y1 = unpack_bf16(yu.y);
        accG.x += d.x * y0.x; accG.y += d.y * y0.y; accG.z += d.z * y1.x; accG.w += d.w * y1.y;
      }
    }
    float* dst = part + ((long long)group * chunks + chunk) * 2 * D;
    *(reinterpret_cast<float4*>(dst) + cg) = accS;
    *(reinterpret_cast<float4*>(dst + D) + cg) = accG;
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
// One launch that finishes every partial-sum buffer of a DiT block's backward (replaces 6 finish_group, 2 finish_all,
// the fp32->bf16 cast of d mod and the adaLN-bias column sum: ten 3-7 us launches per block).
//   pA = gate_bwd of the MLP branch   [B, ch, 2, D]: 0 = sum dx       1 = sum dx * y
//   pB = ln_bwd  of the MLP LayerNorm               : 0 = sum dy       1 = sum dy * xhat
//   pC = gate_bwd of the attention branch, pD = ln_bwd of the attention LayerNorm
//   d mod[n, :] = [d shift_msa, d scale_msa, d gate_msa, d shift_mlp, d scale_mlp, d gate_mlp]  (fp32 + bf16 copy)
//   g_fc2_b[col]  (+)= sum_n gate_mlp[n, col] * pA0[n, col]      g_proj_b[col] (+)= sum_n gate_msa[n, col] * pC0[n, col]
//   g_ada_b[slot * D + col] (+)= sum_n d mod[n, slot, col]        (bias of adaLN_modulation.1, dit.py:139-141)
// block = 32 columns x 16 sample lanes; every sum runs in a fixed order (deterministic).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
dit_block_finish_kernel(const float* __restrict__ pA, const float* __restrict__ pB, const float* __restrict__ pC,
                        const float* __restrict__ pD, int B, int ch, int D, const float* __restrict__ mod,
                        long long ldm, float* __restrict__ dmod, bf16* __restrict__ dmod_b,
                        float* __restrict__ g_fc2_b, float* __restrict__ g_proj_b, float* __restrict__ g_ada_b,
                        int accumulate) {
  // blockIdx.y selects the partial buffer; each buffer feeds two of the eight outputs:
  //   y = 0: pA -> fc2.bias (gate-weighted q0), d gate_mlp (slot 5, q1)      y = 1: pB -> slots 3, 4
  //   y = 2: pC -> proj.bias (gate-weighted q0), d gate_msa (slot 2, q1)     y = 3: pD -> slots 0, 1
  __shared__ float red[16][2][33];
  const int cx = threadIdx.x & 31, ny = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + cx;
  const int which = blockIdx.y;
  const float* part = which == 0 ? pA : which == 1 ? pB : which == 2 ? pC : pD;
  const int slot0 = which == 1 ? 3 : which == 3 ? 0 : -1;            // d mod slot fed by q0 (-1: q0 feeds a bias grad)
  const int slot1 = which == 0 ? 5 : which == 1 ? 4 : which == 2 ? 2 : 1;
  const int gate_slot = which == 0 ? 5 : 2;
  float t0 = 0.f, t1 = 0.f;
  if (col < D) {
    // The kernel is latency-bound (21 MB in 72-element strides, few warps): every thread owns up to four samples per
    // pass and issues the loads of four chunks of ALL of them (32 independent loads) before folding any, which cuts the
    // dependent memory round trips per thread from 12 to 3 at B = 64, ch = 9.  Each sample's sum still runs in chunk
    // order, each thread's sample order is unchanged: results are bit-identical to the one-sample-at-a-time loop.
    constexpr int S = 4;
    for (int n0 = ny; n0 < B; n0 += 16 * S) {
      float q0[S], q1[S];
      const float* pn[S];
#pragma unroll
      for (int j = 0; j < S; ++j) {
        q0[j] = 0.f;
        q1[j] = 0.f;
        const int n = n0 + 16 * j;
        pn[j] = part + ((long long)(n < B ? n : n0) * ch * 2) * D + col;
      }
      int c = 0;
      for (; c + 4 <= ch; c += 4) {
        float a[S][4], b[S][4];
#pragma unroll
        for (int j = 0; j < S; ++j) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            a[j][u] = __ldg(pn[j] + (long long)(c + u) * 2 * D);
            b[j][u] = __ldg(pn[j] + (long long)(c + u) * 2 * D + D);
          }
        }
#pragma unroll
        for (int j = 0; j < S; ++j) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            q0[j] += a[j][u];
            q1[j] += b[j][u];
          }
        }
      }
      for (; c < ch; ++c) {
        float a[S], b[S];
#pragma unroll
        for (int j = 0; j < S; ++j) {
          a[j] = __ldg(pn[j] + (long long)c * 2 * D);
          b[j] = __ldg(pn[j] + (long long)c * 2 * D + D);
        }
#pragma unroll
        for (int j = 0; j < S; ++j) {
          q0[j] += a[j];
          q1[j] += b[j];
        }
      }
#pragma unroll
      for (int j = 0; j < S; ++j) {
        const int n = n0 + 16 * j;
        if (n >= B) break;
        const long long mo = (long long)n * ldm + col;
        dmod[mo + (long long)slot1 * D] = q1[j];
        dmod_b[mo + (long long)slot1 * D] = __float2bfloat16(q1[j]);
        t1 += q1[j];
        if (slot0 >= 0) {
          dmod[mo + (long long)slot0 * D] = q0[j];
          dmod_b[mo + (long long)slot0 * D] = __float2bfloat16(q0[j]);
          t0 += q0[j];
        } else {
          t0 += mod[mo + (long long)gate_slot * D] * q0[j];
        }
      }
    }
  }
  red[ny][0][cx] = t0;
  red[ny][1][cx] = t1;
  __syncthreads();
  if (ny < 2 && col < D) {   // sample-lane ny folds quantity ny over the 16 lanes, in order
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += red[i][ny][cx];
    float* dst;
    if (ny == 1) dst = g_ada_b + (long long)slot1 * D + col;
    else if (slot0 >= 0) dst = g_ada_b + (long long)slot0 * D + col;
    else dst = (which == 0 ? g_fc2_b : g_proj_b) + col;
    *dst = accumulate ? *dst + s : s;
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
#pragma unroll 4
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
// 16-byte loads: a warp covers 4 rows x 64 columns per instruction (lane = row l/8, 8 columns (l%8)*8); the four row
// sub-lanes are folded with shuffles, the eight warps through shared memory, both in fixed order.
__global__ void __launch_bounds__(256)
colsum_bf16_stage1_v8(const bf16* __restrict__ a, long long lda, int M, int N, int rows_per_chunk,
                      float* __restrict__ part) {
  const int lane = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int sub = lane >> 3, cl = (lane & 7) * 8;
  const int col = blockIdx.x * 64 + cl;
  const int r0 = blockIdx.y * rows_per_chunk;
  const int r1 = min(r0 + rows_per_chunk, M);
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < N) {
#pragma unroll 4
    for (int r = r0 + rg * 4 + sub; r < r1; r += 32) {
      const uint4 v = ldg_stream_u4(reinterpret_cast<const uint4*>(a + (long long)r * lda + col));
      const float2 f0 = unpack_bf16(v.x), f1 = unpack_bf16(v.y), f2 = unpack_bf16(v.z), f3 = unpack_bf16(v.w);
      s[0] += f0.x; s[1] += f0.y; s[2] += f1.x; s[3] += f1.y;
      s[4] += f2.x; s[5] += f2.y; s[6] += f3.x; s[7] += f3.y;
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    s[k] += __shfl_xor_sync(0xffffffffu, s[k], 8);
    s[k] += __shfl_xor_sync(0xffffffffu, s[k], 16);
  }
  __shared__ float red[8][64];
  if (sub == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) red[rg][cl + k] = s[k];
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    const int c = blockIdx.x * 64 + threadIdx.x;
    if (c < N) part[(long long)blockIdx.y * N + c] = t;
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

inline int threads_for_columns(int D) {
  int t = ((D / 4) + 31) / 32 * 32;
  if (t > 512) t = 512;
  if (t < 64) t = 64;
  return t;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
namespace {
int launch_ln_fwd(const float* x, const LnRes& res, const float* shift, const float* scale, long long ld_mod,
                  int rows_per_sample, const float* weight, const float* bias, void* y, const LnOut& lo, float* mean,
                  float* rstd, int M, int D, float eps, cudaStream_t stream) {
  VAW_CHECK_ARG(x && y && mean && rstd && M > 0, "vaw_ln_fwd: bad arguments");
  VAW_CHECK_ARG(D % 4 == 0 && D <= kMaxD, "vaw_ln_fwd: D=%d must be a multiple of 4 and <= %d", D, kMaxD);
  VAW_CHECK_ARG((scale == nullptr) == (shift == nullptr), "vaw_ln_fwd: shift and scale go together");
  VAW_CHECK_ARG((weight == nullptr) == (bias == nullptr), "vaw_ln_fwd: weight and bias go together");
  VAW_CHECK_ARG(!scale || rows_per_sample > 0, "vaw_ln_fwd: rows_per_sample");
  VAW_CHECK_ARG(lo.ldy >= D + (lo.ones_block ? 32 : 0) && lo.ldy % 4 == 0 && (!lo.ones_block || (lo.ldy % 8 == 0 && D % 8 == 0)),
                "vaw_ln_fwd: ldy=%lld too small or misaligned for D=%d", lo.ldy, D);
  const int rps = rows_per_sample > 0 ? rows_per_sample : 1;
  if (res.branch) {
    VAW_CHECK_ARG(res.gate && res.x_out && scale && rows_per_sample > 0 && res.ld_gate % 4 == 0 && ld_mod % 4 == 0,
                  "vaw_ln_fwd_res: bad arguments");
    VAW_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(res.x_out) |
                    reinterpret_cast<uintptr_t>(res.gate) | reinterpret_cast<uintptr_t>(shift) |
                    reinterpret_cast<uintptr_t>(scale)) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(res.branch) & 7) == 0 && (reinterpret_cast<uintptr_t>(y) & 7) == 0,
                  "vaw_ln_fwd_res: operands must be 16-byte aligned (bf16 tensors 8-byte)");
  }
  VAW_CHECK_ARG(!lo.ones_block || (reinterpret_cast<uintptr_t>(y) & 15) == 0, "vaw_ln_fwd: y must be 16-byte aligned");
  const int rows_per_cta = 8 * kLnRowsPerWarp;
  const unsigned grid = (unsigned)((M + rows_per_cta - 1) / rows_per_cta);
#define VAW_LN_FWD_K(KV, RES_, EX_)                                                                                \
  ln_fwd_kernel<KV, RES_, EX_, (KV <= 3 ? 4 : KV <= 6 ? 2 : 1)><<<grid, 256, 0, stream>>>(                              \
      x, shift, scale, ld_mod, rps, weight, bias, (bf16*)y, mean, rstd, M, D, eps, res, lo)
#define VAW_LN_FWD(KV)                                                     \
  do {                                                                     \
    const bool exact = D == KV * 128;                                      \
    if (res.branch) {                                                      \
      if (exact) VAW_LN_FWD_K(KV, true, true); else VAW_LN_FWD_K(KV, true, false);   \
    } else {                                                               \
      if (exact) VAW_LN_FWD_K(KV, false, true); else VAW_LN_FWD_K(KV, false, false); \
    }                                                                      \
  } while (0)
  if (D <= 384) VAW_LN_FWD(3);
  else if (D <= 768) VAW_LN_FWD(6);
  else if (D <= 1152) VAW_LN_FWD(9);
  else VAW_LN_FWD(16);
#undef VAW_LN_FWD
#undef VAW_LN_FWD_K
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}
}  // namespace

extern "C" int vaw_ln_fwd(const float* x, const float* shift, const float* scale, long long ld_mod,
                          int rows_per_sample, const float* weight, const float* bias, void* y, float* mean,
                          float* rstd, int M, int D, float eps, cudaStream_t stream) {
  return launch_ln_fwd(x, LnRes{}, shift, scale, ld_mod, rows_per_sample, weight, bias, y, LnOut{D, 0}, mean, rstd, M, D,
                       eps, stream);
}

// Residual update of the previous branch + LayerNorm + adaLN modulate of the next one in ONE pass over the row
// (models/dit.py:133-137: x = x + gate.unsqueeze(1) * branch(...), then modulate(norm(x), shift, scale)):
//   x_out[r, :] = x[r, :] + gate[n, :] * branch[r, :];   y[r, :] = LN(x_out[r, :]) * (1 + scale[n, :]) + shift[n, :]
extern "C" int vaw_ln_fwd_res(const float* x, const void* branch, const float* gate, long long ld_gate, float* x_out,
                              const float* shift, const float* scale, long long ld_mod, int rows_per_sample, void* y,
                              float* mean, float* rstd, int M, int D, float eps, cudaStream_t stream) {
  VAW_CHECK_ARG(branch, "vaw_ln_fwd_res: bad arguments");
  return launch_ln_fwd(x, LnRes{(const bf16*)branch, gate, ld_gate, x_out}, shift, scale, ld_mod, rows_per_sample,
                       nullptr, nullptr, y, LnOut{D, 0}, mean, rstd, M, D, eps, stream);
}

// The general form: optional folded-in branch (branch == NULL: none), modulate or affine, output row stride ldy and the
// optional ones block after each row (see LnOut).
extern "C" int vaw_ln_fwd_ex(const float* x, const void* branch, const float* gate, long long ld_gate, float* x_out,
                             const float* shift, const float* scale, long long ld_mod, int rows_per_sample,
                             const float* weight, const float* bias, void* y, long long ldy, int ones_block, float* mean,
                             float* rstd, int M, int D, float eps, cudaStream_t stream) {
  return launch_ln_fwd(x, LnRes{(const bf16*)branch, gate, ld_gate, x_out}, shift, scale, ld_mod, rows_per_sample, weight,
                       bias, y, LnOut{ldy ? ldy : D, ones_block}, mean, rstd, M, D, eps, stream);
}

// chunks: number of partial rows per group (ceil(rows_per_group / chunks) must be <= 64); part must hold
// groups * chunks * 2 * D floats (may be NULL when neither dA nor dB is wanted).  groups * rows_per_group covers M.
namespace {
// common launcher: fuse == nullptr -> plain LayerNorm backward
int launch_ln_bwd(const void* dy, const float* x, const float* mean, const float* rstd, const float* scale,
                  long long ld_mod, const float* weight, float* dx_io, int add_into, float* part, int rows_per_group,
                  int groups, int chunks, int M, int D, const GateFuse* fuse, cudaStream_t stream) {
  VAW_CHECK_ARG(dy && x && mean && rstd && dx_io && M > 0, "vaw_ln_bwd: bad arguments");
  VAW_CHECK_ARG(D % 4 == 0, "vaw_ln_bwd: D=%d must be a multiple of 4", D);
  VAW_CHECK_ARG(rows_per_group > 0 && groups > 0 && chunks > 0 && (long long)groups * rows_per_group >= M,
                "vaw_ln_bwd: bad grouping");
  const int nvp = ((D / 4) + 31) / 32 * 32;
  int R = kStagedBudget / (12 * D);   // rows per stage, two stages
  if (R > 32) R = 32;
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy)) & 15) == 0;
  const int max_threads = fuse ? 320 : 576;
  if (D % 8 == 0 && R >= 4 && nvp <= max_threads && aligned) {
    int halves = 1;
    while (halves * 2 * nvp <= max_threads && halves * 2 * 4 <= R) halves *= 2;
    const size_t smem = 2 * (size_t)R * D * 6 + (size_t)R * 48 + 16;
    static bool configured = false;
    if (!configured) {
      VAW_CUDA_TRY(cudaFuncSetAttribute(ln_bwd_staged_kernel<576, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        kStagedBudget + 32 * 48 + 16));
      VAW_CUDA_TRY(cudaFuncSetAttribute(ln_bwd_staged_kernel<320, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        kStagedBudget + 32 * 48 + 16));
      configured = true;
    }
    if (fuse)
      ln_bwd_staged_kernel<320, true><<<dim3(chunks, groups), halves * nvp, smem, stream>>>(
          (const bf16*)dy, x, mean, rstd, scale, ld_mod, weight, dx_io, add_into, part, rows_per_group, chunks, M, D, R,
          nvp, *fuse);
    else
      ln_bwd_staged_kernel<576, false><<<dim3(chunks, groups), halves * nvp, smem, stream>>>(
          (const bf16*)dy, x, mean, rstd, scale, ld_mod, weight, dx_io, add_into, part, rows_per_group, chunks, M, D, R,
          nvp, GateFuse{});
    VAW_LAUNCH_CHECK();
    return VAW_OK;
  }
  VAW_CHECK_ARG((rows_per_group + chunks - 1) / chunks <= kBwdMaxRows,
                "vaw_ln_bwd: more than %d rows per chunk (rows_per_group=%d chunks=%d)", kBwdMaxRows, rows_per_group,
                chunks);
  ln_bwd_kernel<<<dim3(chunks, groups), threads_for_columns(D), 0, stream>>>(
      (const bf16*)dy, x, mean, rstd, scale, ld_mod, weight, dx_io, add_into, part, rows_per_group, chunks, M, D);
  VAW_LAUNCH_CHECK();
  if (fuse) {   // shapes outside the fused kernel's range: the two-kernel sequence it replaces
    gate_bwd_kernel<<<dim3(chunks, groups), threads_for_columns(D), 0, stream>>>(dx_io, fuse->y, fuse->gate, fuse->ld_gate,
                                                                                fuse->dy, fuse->part, rows_per_group,
                                                                                chunks, M, D);
    VAW_LAUNCH_CHECK();
  }
  return VAW_OK;
}
}  // namespace

extern "C" int vaw_ln_bwd(const void* dy, const float* x, const float* mean, const float* rstd, const float* scale,
                          long long ld_mod, const float* weight, float* dx_io, int add_into, float* part,
                          int rows_per_group, int groups, int chunks, int M, int D, cudaStream_t stream) {
  return launch_ln_bwd(dy, x, mean, rstd, scale, ld_mod, weight, dx_io, add_into, part, rows_per_group, groups, chunks, M,
                       D, nullptr, stream);
}

// LayerNorm backward fused with the residual-branch backward of the branch that follows in the backward pass:
// afterwards dx_io holds the updated residual gradient dx', dy_next = bf16(dx' * gate_next[group]) and part_gate holds
// (sum dx', sum dx' * y_next) per chunk - what vaw_ln_bwd followed by vaw_gate_bwd(dx_io, y_next, gate_next, ...) give.
extern "C" int vaw_ln_bwd_gate(const void* dy, const float* x, const float* mean, const float* rstd, const float* scale,
                               long long ld_mod, const float* weight, float* dx_io, int add_into, float* part,
                               const void* y_next, const float* gate_next, long long ld_gate, void* dy_next,
                               float* part_gate, int rows_per_group, int groups, int chunks, int M, int D,
                               cudaStream_t stream) {
  VAW_CHECK_ARG(dy_next && part_gate, "vaw_ln_bwd_gate: dy_next and part_gate are required");
  GateFuse gf;
  gf.y = reinterpret_cast<const bf16*>(y_next);
  gf.gate = gate_next;
  gf.ld_gate = ld_gate;
  gf.dy = reinterpret_cast<bf16*>(dy_next);
  gf.part = part_gate;
  return launch_ln_bwd(dy, x, mean, rstd, scale, ld_mod, weight, dx_io, add_into, part, rows_per_group, groups, chunks, M,
                       D, &gf, stream);
}

extern "C" int vaw_gate_bwd(const float* dx, const void* y, const float* gate, long long ld_gate, void* dy,
                            float* part, int rows_per_group, int groups, int chunks, int M, int D,
                            cudaStream_t stream) {
  VAW_CHECK_ARG(dx && dy && part && M > 0, "vaw_gate_bwd: bad arguments");
  VAW_CHECK_ARG(D % 4 == 0, "vaw_gate_bwd: D=%d must be a multiple of 4", D);
  VAW_CHECK_ARG(rows_per_group > 0 && groups > 0 && chunks > 0 && (long long)groups * rows_per_group >= M,
                "vaw_gate_bwd: bad grouping");
  gate_bwd_kernel<<<dim3(chunks, groups), threads_for_columns(D), 0, stream>>>(dx, (const bf16*)y, gate, ld_gate,
                                                                              (bf16*)dy, part, rows_per_group, chunks,
                                                                              M, D);
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

extern "C" int vaw_dit_block_finish(const float* pA, const float* pB, const float* pC, const float* pD, int B, int chunks,
                                    int D, const float* mod, long long ldm, float* dmod, void* dmod_b, float* g_fc2_b,
                                    float* g_proj_b, float* g_ada_b, int accumulate, cudaStream_t stream) {
  VAW_CHECK_ARG(pA && pB && pC && pD && mod && dmod && dmod_b && g_fc2_b && g_proj_b && g_ada_b && B > 0 && chunks > 0 &&
                    D > 0,
                "vaw_dit_block_finish: bad arguments");
  dit_block_finish_kernel<<<dim3((D + 31) / 32, 4), 512, 0, stream>>>(pA, pB, pC, pD, B, chunks, D, mod, ldm, dmod, (bf16*)dmod_b,
                                                            g_fc2_b, g_proj_b, g_ada_b, accumulate);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

// part: scratch of ceil(M / rows_per_chunk) * N floats
extern "C" int vaw_colsum_bf16(const void* a, long long lda, int M, int N, float* part, int rows_per_chunk,
                               float* out, int accumulate, cudaStream_t stream) {
  VAW_CHECK_ARG(a && part && out && M > 0 && N > 0 && N % 2 == 0 && rows_per_chunk > 0, "vaw_colsum_bf16: bad arguments");
  const int chunks = (M + rows_per_chunk - 1) / rows_per_chunk;
  if (N % 8 == 0 && lda % 8 == 0 && (reinterpret_cast<uintptr_t>(a) & 15) == 0)
    colsum_bf16_stage1_v8<<<dim3((N + 63) / 64, chunks), 256, 0, stream>>>((const bf16*)a, lda, M, N, rows_per_chunk, part);
  else
    colsum_bf16_stage1<<<dim3((N + 63) / 64, chunks), 256, 0, stream>>>((const bf16*)a, lda, M, N, rows_per_chunk, part);
  VAW_LAUNCH_CHECK();
  colsum_stage2<<<(N + 255) / 256, 256, 0, stream>>>(part, chunks, N, out, accumulate);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

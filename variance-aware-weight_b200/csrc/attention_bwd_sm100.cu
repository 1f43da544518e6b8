// attention_bwd_sm100.cu — tcgen05 / TMEM self-attention backward for sequences of up to 256 tokens (every DiT
// config).  Autograd backward of F.scaled_dot_product_attention inside timm Attention (models/dit.py:126).
//
// One CTA = one (batch, head).  Q, K, V, dO of the head are TMA-loaded once (64-row boxes, head_dim 72 = a
// SWIZZLE_128B block of 64 columns + a SWIZZLE_32B block of 16 zero-padded columns, see attention_sm100.cu) and every
// product runs on the tensor cores with the key index on the TMEM lanes, so the transposed probabilities feed the
// dV / dK products straight from TMEM:
//   per key tile kt (128 keys) and query block qs (64 queries):
//     S^T  = K[kt] Q[qs]^T        [128 x 64] fp32  TMEM          (SS, both operands K-major)
//     dP^T = V[kt] dO[qs]^T       [128 x 64] fp32  TMEM
//     elementwise (8 warps, 2 threads per key row):  P^T = 2^(S^T c - L2[q]),  dS^T = P^T (dP^T - Delta[q])
//         -> bf16 pairs back to TMEM (A operands) and dS^T also to shared memory (MN-major A operand of dQ)
//     dV[kt] += P^T  dO[qs]       [128 x hd]       TMEM          (TS: A from TMEM, B = dO rows as MN-major)
//     dK[kt] += dS^T Q[qs]
//     every second block:  dQ[128 queries] += dS K[kt]           (SS: A = dS^T tile in smem read M-major)
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warps 2..9 = elementwise + epilogue.
// The MMA warp runs one block ahead: S^T / dP^T of block s+1 are issued as soon as the elementwise warps have pulled
// block s into registers, so the tensor pipe and the exp2 / FMA pipes overlap.
// TMEM (512 columns): S^T 64 | dP^T 64 | P^T 32 | dS^T 32 | dV 80 | dK 80 | dQ 2 x 80.
#include "vaw_tc5.cuh"
#include "vaw_internal.h"

namespace {
using namespace tc5;

constexpr int kMaxT = 256;
constexpr int kBwdThreads = 320;
constexpr int cS = 0, cDP = 64, cPt = 128, cDSt = 160, cDV = 192, cDK = 272, cDQ = 352;  // TMEM columns

template <int HD>
struct BSmem {
  static constexpr bool kTail = HD > 64;
  static constexpr int kMain = kMaxT * 128;            // [256 rows x 64 cols] bf16, SWIZZLE_128B
  static constexpr int kTl = kTail ? kMaxT * 32 : 0;   // [256 rows x 16 cols] bf16, SWIZZLE_32B
  static constexpr int kQ = 0, kK = kMain, kV = 2 * kMain, kDO = 3 * kMain;
  static constexpr int kQt = 4 * kMain, kKt = kQt + kTl, kVt = kKt + kTl, kDOt = kVt + kTl;
  static constexpr int kDS = kDOt + kTl;               // dS^T [2 blocks of 64 queries][128 keys x 128 B]
  static constexpr int kL2 = kDS + 2 * 128 * 128;      // -L2[q]  fp32 [256]
  static constexpr int kDelta = kL2 + kMaxT * 4;       // Delta[q] fp32 [256]
  static constexpr int kBars = kDelta + kMaxT * 4;
  static constexpr int kBytes = kBars + 128 + 1024;
};
// (An all-SWIZZLE_32B layout - five 16-column blocks per tensor, so that N = 80 goes into one MMA - was measured:
//  35 % fewer tcgen05.mma but 2.5x more TMA requests of 32 B each; the kernel got 20 % slower.  Kept: 64 + 16.)

__device__ __forceinline__ float dot8(uint4 a, uint4 b) {
  const float2 a0 = unpack_bf16(a.x), a1 = unpack_bf16(a.y), a2 = unpack_bf16(a.z), a3 = unpack_bf16(a.w);
  const float2 b0 = unpack_bf16(b.x), b1 = unpack_bf16(b.y), b2 = unpack_bf16(b.z), b3 = unpack_bf16(b.w);
  return (a0.x * b0.x + a0.y * b0.y) + (a1.x * b1.x + a1.y * b1.y) + (a2.x * b2.x + a2.y * b2.y) +
         (a3.x * b3.x + a3.y * b3.y);
}

// 32 fp32 accumulator columns -> bf16 -> 64 contiguous bytes of global memory
__device__ __forceinline__ void store_row32(bf16* dst, const uint32_t (&v)[32], float mul) {
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    uint4 w;
    w.x = pack_bf16(__uint_as_float(v[j]) * mul, __uint_as_float(v[j + 1]) * mul);
    w.y = pack_bf16(__uint_as_float(v[j + 2]) * mul, __uint_as_float(v[j + 3]) * mul);
    w.z = pack_bf16(__uint_as_float(v[j + 4]) * mul, __uint_as_float(v[j + 5]) * mul);
    w.w = pack_bf16(__uint_as_float(v[j + 6]) * mul, __uint_as_float(v[j + 7]) * mul);
    *reinterpret_cast<uint4*>(dst + j) = w;
  }
}
__device__ __forceinline__ void store_row8(bf16* dst, const uint32_t (&v)[16], float mul) {
  uint4 w;
  w.x = pack_bf16(__uint_as_float(v[0]) * mul, __uint_as_float(v[1]) * mul);
  w.y = pack_bf16(__uint_as_float(v[2]) * mul, __uint_as_float(v[3]) * mul);
  w.z = pack_bf16(__uint_as_float(v[4]) * mul, __uint_as_float(v[5]) * mul);
  w.w = pack_bf16(__uint_as_float(v[6]) * mul, __uint_as_float(v[7]) * mul);
  *reinterpret_cast<uint4*>(dst) = w;
}

template <int HD>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_qkv_t,
                   const __grid_constant__ CUtensorMap tm_do, const __grid_constant__ CUtensorMap tm_do_t,
                   const __grid_constant__ CUtensorMap tm_dq, const __grid_constant__ CUtensorMap tm_dq_t,
                   const bf16* __restrict__ o, const bf16* __restrict__ d_o, const float* __restrict__ lse2,
                   bf16* __restrict__ dqkv, int T, int Ta, int H, float scale, float scale_log2e,
                   const float* __restrict__ delta_in, unsigned long long* __restrict__ trace) {
  using S = BSmem<HD>;
  constexpr bool kTail = S::kTail;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* s_nl2 = reinterpret_cast<float*>(smem + S::kL2);
  float* s_delta = reinterpret_cast<float*>(smem + S::kDelta);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBars);
  uint64_t* kv_full = bars;          // [2]  TMA: K, V rows of key tile kt
  uint64_t* qdo_full = bars + 2;     // [4]  TMA: Q, dO rows of query block qs
  uint64_t* sdp_full = bars + 6;     // MMA  -> EW : S^T / dP^T of block s are in TMEM
  uint64_t* sdp_free = bars + 7;     // EW   -> MMA: block s is in registers (8 warp arrivals)
  uint64_t* p_full = bars + 8;       // EW   -> MMA: P^T / dS^T of block s written (TMEM + smem)
  uint64_t* p_free = bars + 9;       // MMA  -> EW : dV / dK products of block s done (TMEM P^T / dS^T reusable)
  uint64_t* ds_free = bars + 10;     // MMA  -> EW : dQ product done (smem dS^T reusable)
  uint64_t* acc_full = bars + 11;    // MMA  -> EW : dV / dK of a key tile (and finally dQ) complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, b = blockIdx.y;
  const int nkt = (T + 127) >> 7, nqs = (T + 63) >> 6;
  const int nblocks = nkt * nqs;
  const int halves_per_kt = (nqs + 1) >> 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 6; ++i) mbar_init(&bars[i], 1);
    mbar_init(sdp_full, 1);
    mbar_init(sdp_free, 8);
    mbar_init(p_full, 8);
    mbar_init(p_free, 1);
    mbar_init(ds_free, 1);
    mbar_init(acc_full, 1);
    fence_mbar_init_cta();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_v = *tmem_slot;
  const uint32_t tmem = tmem_v;
  // debug timeline (vaw_attn_set_trace): CTA (h = 0, b = gridDim.y / 2) stamps globaltimer at its hand-off points;
  // slot layout: [role 0 = elementwise warp 2, role 1 = MMA warp][64 stamps]
  const bool tracing = trace != nullptr && lane == 0 && h == 0 && b == (int)gridDim.y / 2;
  int tpos = 0;
  auto stamp = [&](int role) {
    if (tracing && tpos < 64) trace[role * 64 + tpos++] = globaltimer_ns();
  };

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      const uint32_t blk_bytes = 64 * 128 + (kTail ? 64 * 32 : 0);
      auto load = [&](const CUtensorMap* mm, const CUtensorMap* mt, int slot, int rb, int main_off, int tail_off,
                      uint64_t* bar) {
        tma_load_4d(smem + main_off + rb * 64 * 128, mm, bar, 0, slot, rb * 64, b);
        if (kTail) tma_load_4d(smem + tail_off + rb * 64 * 32, mt, bar, 64, slot, rb * 64, b);
      };
      auto load_kv = [&](int kt) {
        mbar_expect_tx(&kv_full[kt], 4 * blk_bytes);
        for (int rb = 2 * kt; rb < 2 * kt + 2; ++rb) {
          load(&tm_qkv, &tm_qkv_t, H + h, rb, S::kK, S::kKt, &kv_full[kt]);
          load(&tm_qkv, &tm_qkv_t, 2 * H + h, rb, S::kV, S::kVt, &kv_full[kt]);
        }
      };
      load_kv(0);
      for (int qs = 0; qs < nqs; ++qs) {
        mbar_expect_tx(&qdo_full[qs], 2 * blk_bytes);
        load(&tm_qkv, &tm_qkv_t, h, qs, S::kQ, S::kQt, &qdo_full[qs]);
        load(&tm_do, &tm_do_t, h, qs, S::kDO, S::kDOt, &qdo_full[qs]);
      }
      if (nkt > 1) load_kv(1);
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    // The whole warp walks the loop so that every descriptor is warp-uniform (uniform datapath, no per-MMA address
    // arithmetic in vector registers); only lane 0 issues.  Descriptors are base + (byte offset >> 4) in the low word.
    const bool leader = lane == 0;
    const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_v, 0);   // provably warp-uniform: stays in uniform registers
    const uint64_t dQ = desc_sw128(smem_u32(smem + S::kQ)), dK = desc_sw128(smem_u32(smem + S::kK)),
                   dV = desc_sw128(smem_u32(smem + S::kV)), dDO = desc_sw128(smem_u32(smem + S::kDO)),
                   dQt = desc_sw32(smem_u32(smem + S::kQt)), dKt = desc_sw32(smem_u32(smem + S::kKt)),
                   dVt = desc_sw32(smem_u32(smem + S::kVt)), dDOt = desc_sw32(smem_u32(smem + S::kDOt)),
                   dDS = desc_sw128_mn(smem_u32(smem + S::kDS), 16384);
    const uint32_t id_kk = idesc_bf16(64, 0, 0);      // S^T, dP^T : N = 64 queries, K-major x K-major
    const uint32_t id_m = idesc_bf16(64, 1, 0);       // dV, dK main columns: B MN-major
    const uint32_t id_t = idesc_bf16(16, 1, 0);       // dV, dK tail columns
    const uint32_t id_qm = idesc_bf16(64, 1, 1);      // dQ: A and B MN-major
    const uint32_t id_qt = idesc_bf16(16, 1, 1);
    auto issue_sdp = [&](int s) {
      const int kt = s / nqs, qs = s - kt * nqs;
      if (qs == 0) mbar_wait(&kv_full[kt], 0);
      if (kt == 0) mbar_wait(&qdo_full[qs], 0);
      tc_fence_after();
      const uint64_t ak = dK + (uint32_t)(kt * (16384 >> 4)), av = dV + (uint32_t)(kt * (16384 >> 4));
      const uint64_t bq = dQ + (uint32_t)(qs * (8192 >> 4)), bo = dDO + (uint32_t)(qs * (8192 >> 4));
      if (leader) {
        // consecutive MMAs alternate between the two accumulators.  (Measured: no faster than accumulator-by-accumulator
        // order - the ~55 ns per small MMA is a per-instruction cost, not accumulate latency.)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          tc_mma_ss(tmem + cS, ak + 2 * k, bq + 2 * k, id_kk, k ? 1u : 0u);
          tc_mma_ss(tmem + cDP, av + 2 * k, bo + 2 * k, id_kk, k ? 1u : 0u);
        }
        if (kTail) {
          tc_mma_ss(tmem + cS, dKt + (uint32_t)(kt * (4096 >> 4)), dQt + (uint32_t)(qs * (2048 >> 4)), id_kk, 1u);
          tc_mma_ss(tmem + cDP, dVt + (uint32_t)(kt * (4096 >> 4)), dDOt + (uint32_t)(qs * (2048 >> 4)), id_kk, 1u);
        }
        tc_commit(sdp_full);
      }
      __syncwarp();
    };
    stamp(1);
    issue_sdp(0);
    stamp(1);
    for (int s = 0; s < nblocks; ++s) {
      const int kt = s / nqs, qs = s - kt * nqs;
      mbar_wait(sdp_free, (uint32_t)s & 1u);
      stamp(1);
      if (s + 1 < nblocks) issue_sdp(s + 1);
      stamp(1);
      mbar_wait(p_full, (uint32_t)s & 1u);
      stamp(1);
      tc_fence_after();
      const uint64_t bo = dDO + (uint32_t)(qs * (8192 >> 4)), bot = dDOt + (uint32_t)(qs * (2048 >> 4));
      const uint64_t bq = dQ + (uint32_t)(qs * (8192 >> 4)), bqt = dQt + (uint32_t)(qs * (2048 >> 4));
      const uint32_t acc0 = qs ? 1u : 0u;
      if (leader) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {   // dV[kt] += P^T dO[qs], dK[kt] += dS^T Q[qs]: four accumulators round-robin
          tc_mma_ts(tmem + cDV, tmem + cPt + k * 8, bo + 128 * k, id_m, k ? 1u : acc0);
          tc_mma_ts(tmem + cDK, tmem + cDSt + k * 8, bq + 128 * k, id_m, k ? 1u : acc0);
          if (kTail) {
            tc_mma_ts(tmem + cDV + 64, tmem + cPt + k * 8, bot + 32 * k, id_t, k ? 1u : acc0);
            tc_mma_ts(tmem + cDK + 64, tmem + cDSt + k * 8, bqt + 32 * k, id_t, k ? 1u : acc0);
          }
        }
        tc_commit(p_free);
      }
      if ((qs & 1) || qs == nqs - 1) {   // a 128-query half is complete: dQ[half] += dS K[kt]
        const uint32_t dq = tmem + cDQ + (qs >> 1) * 80;
        const uint64_t bk = dK + (uint32_t)(kt * (16384 >> 4)), bkt = dKt + (uint32_t)(kt * (4096 >> 4));
        const uint32_t accq = kt ? 1u : 0u;
        if (leader) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            tc_mma_ss(dq, dDS + 128 * k, bk + 128 * k, id_qm, k ? 1u : accq);
            if (kTail) tc_mma_ss(dq + 64, dDS + 128 * k, bkt + 32 * k, id_qt, k ? 1u : accq);
          }
          tc_commit(ds_free);
        }
      }
      if (leader && qs == nqs - 1) tc_commit(acc_full);
      __syncwarp();
      stamp(1);
    }
  } else {
    // =========================== elementwise + epilogue warps ===========================
    const int g = warp & 3;              // TMEM lane group this warp may touch
    const int hq = (warp - 2) >> 2;      // which 32 of the block's 64 queries
    const int j = g * 32 + lane;         // key row inside the tile == TMEM lane
    const uint32_t trow = tmem + ((uint32_t)(g * 32) << 16);
    const bool tr_ew = warp == 2;
    if (tr_ew) stamp(0);
    {  // per-query constants: -L2[q] and Delta[q] = sum_d dO[q, d] O[q, d]
      const int q = (warp - 2) * 32 + lane;
      float nl2 = -INFINITY, delta = 0.f;
      if (q < T) {
        if (delta_in) {   // precomputed by attn_delta_kernel (coalesced, full bandwidth)
          delta = delta_in[((long long)b * H + h) * Ta + q];
        } else {
          const long long off = (((long long)b * Ta + q) * H + h) * HD;
          const uint4* po = reinterpret_cast<const uint4*>(o + off);
          const uint4* pd = reinterpret_cast<const uint4*>(d_o + off);
#pragma unroll
          for (int i = 0; i < HD / 8; ++i) delta += dot8(__ldg(po + i), __ldg(pd + i));
        }
        nl2 = -lse2[((long long)b * H + h) * Ta + q];
      }
      s_nl2[q] = nl2;
      s_delta[q] = delta;
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    if (tr_ew) stamp(0);
    // accumulator drains (lanes = rows of the output): dV / dK of key tile kt, then dQ
    // Accumulator drains.  The gradients leave through shared memory and the TMA store engine: a thread's row is
    // written into the (dead) operand tile of the same geometry - dK over K[kt], dV over V[kt], dQ over Q - in the
    // swizzled box layout, then one thread issues bulk tensor stores.  (Per-thread 16-byte global stores of rows that
    // are 6.9 KB apart cost 2 us per drain in the LSU; the staged version is off the critical path.)
    auto stage_row = [&](int main_off, int tail_off, int row, const uint32_t (&v)[32], const uint32_t (&w)[16],
                         float mul) {
      uint8_t* rm = smem + main_off + row * 128;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 x;
        x.x = pack_bf16(__uint_as_float(v[8 * c]) * mul, __uint_as_float(v[8 * c + 1]) * mul);
        x.y = pack_bf16(__uint_as_float(v[8 * c + 2]) * mul, __uint_as_float(v[8 * c + 3]) * mul);
        x.z = pack_bf16(__uint_as_float(v[8 * c + 4]) * mul, __uint_as_float(v[8 * c + 5]) * mul);
        x.w = pack_bf16(__uint_as_float(v[8 * c + 6]) * mul, __uint_as_float(v[8 * c + 7]) * mul);
        *reinterpret_cast<uint4*>(rm + (((hq * 4 + c) ^ (row & 7)) * 16)) = x;
      }
      if (kTail && hq == 0) {
        uint8_t* rt = smem + tail_off + row * 32;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint4 x;
          x.x = pack_bf16(__uint_as_float(w[8 * c]) * mul, __uint_as_float(w[8 * c + 1]) * mul);
          x.y = pack_bf16(__uint_as_float(w[8 * c + 2]) * mul, __uint_as_float(w[8 * c + 3]) * mul);
          x.z = pack_bf16(__uint_as_float(w[8 * c + 4]) * mul, __uint_as_float(w[8 * c + 5]) * mul);
          x.w = pack_bf16(__uint_as_float(w[8 * c + 6]) * mul, __uint_as_float(w[8 * c + 7]) * mul);
          *reinterpret_cast<uint4*>(rt + ((c ^ ((row >> 2) & 1)) * 16)) = x;
        }
      }
    };
    auto store_rows = [&](int main_off, int tail_off, int slot, int rb) {   // one thread: rows [64 rb, 64 rb + 64)
      tma_store_4d(&tm_dq, smem + main_off + rb * 64 * 128, 0, slot, rb * 64, b);
      if (kTail) tma_store_4d(&tm_dq_t, smem + tail_off + rb * 64 * 32, 64, slot, rb * 64, b);
    };
    auto drain_dvdk = [&](int kt) {
      mbar_wait(acc_full, (uint32_t)kt & 1u);
      __syncwarp();
      tc_fence_after();
      uint32_t v[32], u[32], wv[16], wk[16];
      tmem_ld32(trow + cDV + hq * 32, v);
      tmem_ld32(trow + cDK + hq * 32, u);
      if (kTail && hq == 0) {
        tmem_ld16(trow + cDV + 64, wv);
        tmem_ld16(trow + cDK + 64, wk);
      }
      tmem_ld_wait();
      stage_row(S::kV, S::kVt, kt * 128 + j, v, wv, 1.f);
      stage_row(S::kK, S::kKt, kt * 128 + j, u, wk, scale);
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (warp == 2 && lane == 0) {
        for (int rb = 2 * kt; rb < 2 * kt + 2; ++rb) {
          if (rb * 64 < T) {
            store_rows(S::kV, S::kVt, 2 * H + h, rb);
            store_rows(S::kK, S::kKt, H + h, rb);
          }
        }
        bulk_commit_group();
      }
    };

    for (int s = 0; s < nblocks; ++s) {
      const int kt = s / nqs, qs = s - kt * nqs;
      mbar_wait(sdp_full, (uint32_t)s & 1u);
      __syncwarp();
      if (tr_ew) stamp(0);
      tc_fence_after();
      uint32_t sv[32], dv[32];
      tmem_ld32(trow + cS + hq * 32, sv);
      tmem_ld32(trow + cDP + hq * 32, dv);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sdp_free);

      uint32_t pp[16], dd[16];
      const float4* nl = reinterpret_cast<const float4*>(s_nl2 + qs * 64 + hq * 32);
      const float4* dl = reinterpret_cast<const float4*>(s_delta + qs * 64 + hq * 32);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 L = nl[i], D = dl[i];
        const float p0 = ex2_approx(fmaf(__uint_as_float(sv[4 * i]), scale_log2e, L.x));
        const float p1 = ex2_approx(fmaf(__uint_as_float(sv[4 * i + 1]), scale_log2e, L.y));
        const float p2 = ex2_approx(fmaf(__uint_as_float(sv[4 * i + 2]), scale_log2e, L.z));
        const float p3 = ex2_approx(fmaf(__uint_as_float(sv[4 * i + 3]), scale_log2e, L.w));
        pp[2 * i] = pack_bf16(p0, p1);
        pp[2 * i + 1] = pack_bf16(p2, p3);
        dd[2 * i] = pack_bf16(p0 * (__uint_as_float(dv[4 * i]) - D.x), p1 * (__uint_as_float(dv[4 * i + 1]) - D.y));
        dd[2 * i + 1] =
            pack_bf16(p2 * (__uint_as_float(dv[4 * i + 2]) - D.z), p3 * (__uint_as_float(dv[4 * i + 3]) - D.w));
      }
      if (tr_ew) stamp(0);
      // previous key tile's dV / dK leave TMEM here: after this block's math (the MMA warp has finished them by now)
      // and before p_full lets the first dV / dK product of this tile overwrite them
      if (qs == 0 && kt > 0) drain_dvdk(kt - 1);
      if (s > 0) {
        mbar_wait(p_free, (uint32_t)(s - 1) & 1u);
        __syncwarp();
        tc_fence_after();
      }
      if (tr_ew) stamp(0);
      tmem_st16(trow + cPt + hq * 16, pp);
      tmem_st16(trow + cDSt + hq * 16, dd);
      const int half_idx = kt * halves_per_kt + (qs >> 1);
      if (half_idx > 0) mbar_wait(ds_free, (uint32_t)(half_idx - 1) & 1u);
      if (tr_ew) stamp(0);
      {  // dS^T row j, queries [hq * 32, hq * 32 + 32) of block (qs & 1): four swizzled 16-byte chunks
        uint8_t* rowp = smem + S::kDS + (qs & 1) * 16384 + j * 128;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int chunk = (hq * 4 + c) ^ (j & 7);
          *reinterpret_cast<uint4*>(rowp + chunk * 16) = make_uint4(dd[4 * c], dd[4 * c + 1], dd[4 * c + 2], dd[4 * c + 3]);
        }
      }
      fence_proxy_async_smem();
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (tr_ew) stamp(0);
    }
    drain_dvdk(nkt - 1);
    if (tr_ew) stamp(0);   // the last acc_full also covers every dQ product
    {
      const int nqh = (T + 127) >> 7;
      uint32_t v[2][32], w[2][16];
#pragma unroll
      for (int qh = 0; qh < 2; ++qh) {
        if (qh < nqh) {
          tmem_ld32(trow + cDQ + qh * 80 + hq * 32, v[qh]);
          if (kTail && hq == 0) tmem_ld16(trow + cDQ + qh * 80 + 64, w[qh]);
        }
      }
      tmem_ld_wait();
#pragma unroll
      for (int qh = 0; qh < 2; ++qh)
        if (qh < nqh) stage_row(S::kQ, S::kQt, qh * 128 + j, v[qh], w[qh], scale);
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (warp == 2 && lane == 0) {
        for (int rb = 0; rb * 64 < T; ++rb) store_rows(S::kQ, S::kQt, h, rb);
        bulk_commit_group();
        bulk_wait_group_read0();   // shared memory must stay valid until the store engine has read it
      }
    }
  }
  if (warp == 2) stamp(0);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// Delta[b, h, t] = sum_d dO[b, t, h, d] * O[b, t, h, d]; one thread per (token, head), consecutive threads read
// consecutive 2*HD-byte segments (coalesced across the warp).
template <int HD>
__global__ void __launch_bounds__(256)
attn_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ d_o, float* __restrict__ delta, int T, int H,
                  long long n) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const uint4* po = reinterpret_cast<const uint4*>(o + i * HD);
  const uint4* pd = reinterpret_cast<const uint4*>(d_o + i * HD);
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < HD / 8; ++k) acc += dot8(__ldg(po + k), __ldg(pd + k));
  const int h = (int)(i % H);
  const long long bt = i / H;
  const long long b = bt / T;
  const int t = (int)(bt - b * T);
  delta[(b * H + h) * T + t] = acc;
}

// Coalesced version for 256 % H == 0: a CTA owns 256 consecutive (token, head) segments = a contiguous range of
// 256 * HD / 8 16-byte chunks.  Thread i takes chunks i, i + 256, ... of O and dO (fully coalesced), leaves their partial
// dot products in shared memory, then thread j folds the HD / 8 partials of segment j (stride HD / 8 is odd for hd 72:
// conflict-free) and the results are written head-major, 256 / H consecutive tokens per head.
template <int HD>
__global__ void __launch_bounds__(256)
attn_delta_coalesced_kernel(const bf16* __restrict__ o, const bf16* __restrict__ d_o, float* __restrict__ delta, int T,
                            int H, long long n_seg) {
  constexpr int CPS = HD / 8;   // 16-byte chunks per segment
  __shared__ float partial[256 * CPS];
  __shared__ float seg[256];
  const long long seg0 = (long long)blockIdx.x * 256;
  const long long nseg_here = min((long long)256, n_seg - seg0);
  const uint4* po = reinterpret_cast<const uint4*>(o) + seg0 * CPS;
  const uint4* pd = reinterpret_cast<const uint4*>(d_o) + seg0 * CPS;
  const int nchunks = (int)nseg_here * CPS;
#pragma unroll
  for (int k = 0; k < CPS; ++k) {
    const int i = threadIdx.x + 256 * k;
    if (i < nchunks) partial[i] = dot8(__ldg(po + i), __ldg(pd + i));
  }
  __syncthreads();
  if (threadIdx.x < nseg_here) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < CPS; ++k) acc += partial[threadIdx.x * CPS + k];
    seg[threadIdx.x] = acc;
  }
  __syncthreads();
  // segment s = token * H + head (within the CTA: tokens_per_cta = 256 / H whole tokens); write head-major
  const int tpc = 256 / H;
  const int hh = threadIdx.x / tpc, tl = threadIdx.x % tpc;
  const long long token = seg0 / H + tl;          // global token index b * T + t
  if (tl * H + hh < nseg_here) {
    const long long b = token / T;
    const int t = (int)(token - b * T);
    delta[(b * H + hh) * T + t] = seg[tl * H + hh];
  }
}

unsigned long long* g_attn_trace = nullptr;   // debug: device buffer of 128 stamps (vaw_attn_set_trace)

template <int HD>
int launch_bwd_tc(const void* qkv, const void* o, const void* d_o, const float* lse2, void* dqkv, float* delta_ws, int B,
                  int T, int Ta, int H, cudaStream_t stream) {
  using S = BSmem<HD>;
  CUtensorMap tq, tqt, td, tdt;
  int rc = make_head_map(&tq, qkv, B, T, Ta, 3 * H, HD, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B);
  if (!rc) rc = make_head_map(&tqt, qkv, B, T, Ta, 3 * H, HD, 16, 64, CU_TENSOR_MAP_SWIZZLE_32B);
  if (!rc) rc = make_head_map(&td, d_o, B, T, Ta, H, HD, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B);
  if (!rc) rc = make_head_map(&tdt, d_o, B, T, Ta, H, HD, 16, 64, CU_TENSOR_MAP_SWIZZLE_32B);
  CUtensorMap tg, tgt;
  if (!rc) rc = make_head_map(&tg, dqkv, B, T, Ta, 3 * H, HD, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B);
  if (!rc) rc = make_head_map(&tgt, dqkv, B, T, Ta, 3 * H, HD, 16, 64, CU_TENSOR_MAP_SWIZZLE_32B);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    VAW_CUDA_TRY(cudaFuncSetAttribute(attn_bwd_tc_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kBytes));
    configured = true;
  }
  if (delta_ws) {
    const long long n = (long long)B * Ta * H;   // every token of every sample, the border rows included
    if (256 % H == 0)
      attn_delta_coalesced_kernel<HD><<<(unsigned)((n + 255) / 256), 256, 0, stream>>>((const bf16*)o, (const bf16*)d_o,
                                                                                     delta_ws, Ta, H, n);
    else
      attn_delta_kernel<HD><<<(unsigned)((n + 255) / 256), 256, 0, stream>>>((const bf16*)o, (const bf16*)d_o, delta_ws, Ta,
                                                                           H, n);
    VAW_LAUNCH_CHECK();
  }
  const float scale = 1.0f / sqrtf((float)HD);
  attn_bwd_tc_kernel<HD><<<dim3(H, B), kBwdThreads, S::kBytes, stream>>>(
      tq, tqt, td, tdt, tg, tgt, (const bf16*)o, (const bf16*)d_o, lse2, (bf16*)dqkv, T, Ta, H, scale, scale * 1.4426950408889634f,
      delta_ws, g_attn_trace);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

}  // namespace

// debug knob: device buffer [2][64] of globaltimer stamps written by one CTA of the backward kernel (null = off)
extern "C" int vaw_attn_set_trace(void* device_buf) {
  g_attn_trace = reinterpret_cast<unsigned long long*>(device_buf);
  return VAW_OK;
}

int vaw_attn_bwd_sm100(const void* qkv, const void* o, const void* d_o, const float* lse2, void* dqkv, float* delta_ws,
                       int B, int T, int H, int head_dim, cudaStream_t stream) {
  const uintptr_t al = reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(o) |
                       reinterpret_cast<uintptr_t>(d_o) | reinterpret_cast<uintptr_t>(dqkv);
  if ((al & 15) != 0 || (head_dim != 64 && head_dim != 72)) return VAW_ERR_UNSUPPORTED;
  // T in (256, 264]: the tensor-core kernel on the leading 256 tokens with the FINAL log-sum-exp and Delta (so its
  // probabilities are the true ones), then the border strip (attention_border.cu).  Needs the Delta scratch.
  if (T > kMaxT && !(vaw_attn_border_supported(T) && delta_ws)) return VAW_ERR_UNSUPPORTED;
  const int Tc = T > kMaxT ? kMaxT : T;
  int rc = head_dim == 64 ? launch_bwd_tc<64>(qkv, o, d_o, lse2, dqkv, delta_ws, B, Tc, T, H, stream)
                          : launch_bwd_tc<72>(qkv, o, d_o, lse2, dqkv, delta_ws, B, Tc, T, H, stream);
  if (rc == VAW_OK && T > kMaxT) rc = vaw_attn_border_bwd(qkv, d_o, lse2, delta_ws, dqkv, B, T, H, head_dim, stream);
  return rc;
}

// gemm_sm100.cu — K4: bf16 GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM,
// operands staged by TMA with 128-byte swizzle), with the DiT / U-ViT epilogues fused in.
//
//   D[M,N] = A[M,K] * B[N,K]^T   (fp32 accumulate)
//
// Replaces the cuBLAS calls behind nn.Linear in the reference's blocks (models/dit.py:118-137 via timm
// Attention/Mlp, models/uvit.py:96-121) for forward, dgrad and wgrad:
//   forward  Y  = X  W^T          A = X  [M,K]  K-major      B = W  [N,K]   K-major
//   dgrad    dX = dY W            A = dY [M,Nout] K-major    B = W  [Nout,Kin] used as MN-major (no transposed copy)
//   wgrad    dW = dY^T X          A = dY [tokens,Nout] MN-major,  B = X [tokens,Kin] MN-major
//
// Kernel shape (one persistent CTA per SM, 192 threads):
//   warp 0      TMA producer   (cp.async.bulk.tensor -> STAGES x {A 128x64, B BNx64} ring, mbarrier full/empty)
//   warp 1      MMA issuer     (one lane issues tcgen05.mma 128xBNx16, commits to the ring / to the epilogue)
//   warps 2-5   epilogue       (tcgen05.ld 32 lanes x 32 columns -> registers -> fused epilogue -> global)
// Two accumulator stages in TMEM (2 x BN columns) let the epilogue of tile i overlap the MMAs of tile i+1.
#include "vaw_common.cuh"
#include "vaw_internal.h"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kThreads = 192;
constexpr int kABytes = BM * BK * 2;  // 16 KB per stage

// epilogue selectors (mirrored in include/vaw_b200.h)
enum : int {
  EPI_BF16 = 0,        // out_bf16 = acc + bias
  EPI_F32 = 1,         // out_f32  = acc + bias (+ out_f32 when accumulate)
  EPI_GELU_TANH = 2,   // out_bf16 = pre = acc + bias ; out2_bf16 = gelu_tanh(pre)
  EPI_GELU_ERF = 3,    // same with exact GELU (U-ViT)
  EPI_GATE_RES = 4,    // out_bf16 = y = acc + bias ; out2_f32 = resid + gate[row / rows_per_sample, col] * y  (adaLN-Zero)
  EPI_RES = 5,         // out2_f32 = resid + (acc + bias)                                                    (U-ViT)
  EPI_DGELU_TANH = 6,  // out_bf16 = acc * gelu_tanh'(aux)
  EPI_DGELU_ERF = 7,   // out_bf16 = acc * gelu_erf'(aux)
  EPI_SILU = 8,        // out_bf16 = pre ; out2_bf16 = silu(pre)           (REPA projector)
  EPI_DSILU = 9,       // out_bf16 = acc * silu'(aux)
  EPI_COUNT = 10
};

struct EpiParams {
  void* out;
  void* out2;
  const float* bias;
  const float* resid;
  const float* gate;
  const bf16* aux;
  long long ldo;    // leading dimension (elements) of out/out2/resid/aux
  long long ldg;    // leading dimension of gate
  int rows_per_sample;
  int resid_mod;    // > 0: the residual is a [resid_mod, N] table indexed by row % resid_mod (pos_embed)
  int accumulate;
  int M, N, K;
  int a_mn, b_mn;   // operand majorness: 0 = K-major, 1 = MN-major
  int k_splits;     // > 1: split-K; work item = (tile, split); raw fp32 partials go to out + split * split_stride
  int kb_per_split;
  long long split_stride;
};

// ---------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a pipeline bug must surface as a trapped kernel (launch failure), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = globaltimer_ns();
  while (!mbar_try_wait(bar, parity)) {
    if (globaltimer_ns() - t0 > 4000000000ull) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor, SWIZZLE_128B (cute/arch/mma_sm100_desc.hpp field layout):
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout (2 = 128B swizzle)
// K-major tile  : rows of 128 B (64 bf16 along K), 8-row groups 1024 B apart  -> SBO = 1024, LBO unused (=1)
// MN-major tile : 64(MN) x 8(K) atoms of 1024 B; next 8 k's at +1024 (SBO); next 64 MN at +8192 (LBO)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, int mn_major) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(mn_major ? 512u : 1u) << 16;
  d |= (uint64_t)64u << 32;
  d |= (uint64_t)1u << 46;
  d |= (uint64_t)2u << 61;
  return d;
}

// ---------------------------------------------------------------------------------------------------
// fused epilogue on one 32-column chunk of one row (also reused by the tail-split fix-up kernel)
// ---------------------------------------------------------------------------------------------------
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const EpiParams& p, int row, int col0, const uint32_t (&acc)[32],
                                               long long slab = 0) {
  if (row >= p.M) return;
  const long long ro = (long long)row * p.ldo;
  const long long rr = (long long)(p.resid_mod > 0 ? row % p.resid_mod : row) * p.ldo;
#pragma unroll
  for (int g = 0; g < 4; ++g) {  // 4 groups of 8 columns
    const int c = col0 + g * 8;
    if (c >= p.N) break;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(acc[g * 8 + j]);
    if (p.bias) {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + c));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + c) + 1);
      v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
      v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
    }
    if constexpr (EPI == EPI_F32) {
      float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + slab + ro + c);
      if (p.accumulate) {
        const float4 o0 = o[0], o1 = o[1];
        v[0] += o0.x; v[1] += o0.y; v[2] += o0.z; v[3] += o0.w;
        v[4] += o1.x; v[5] += o1.y; v[6] += o1.z; v[7] += o1.w;
      }
      o[0] = make_float4(v[0], v[1], v[2], v[3]);
      o[1] = make_float4(v[4], v[5], v[6], v[7]);
    } else if constexpr (EPI == EPI_RES) {
      const float4* r = reinterpret_cast<const float4*>(p.resid + rr + c);
      const float4 r0 = r[0], r1 = r[1];
      float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out2) + ro + c);
      // the linear output is a bf16 tensor in the reference's autocast path: round before the residual add
      float y[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = __bfloat162float(__float2bfloat16_rn(v[j]));
      o[0] = make_float4(r0.x + y[0], r0.y + y[1], r0.z + y[2], r0.w + y[3]);
      o[1] = make_float4(r1.x + y[4], r1.y + y[5], r1.z + y[6], r1.w + y[7]);
    } else {
      // every other epilogue writes a bf16 primary output
      if constexpr (EPI == EPI_DGELU_TANH || EPI == EPI_DGELU_ERF || EPI == EPI_DSILU) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(p.aux + ro + c));
        const float2 a0 = unpack_bf16(a.x), a1 = unpack_bf16(a.y), a2 = unpack_bf16(a.z), a3 = unpack_bf16(a.w);
        const float h[8] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, a3.x, a3.y};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if constexpr (EPI == EPI_DGELU_TANH) v[j] *= gelu_tanh_grad_f(h[j]);
          else if constexpr (EPI == EPI_DGELU_ERF) v[j] *= gelu_erf_grad_f(h[j]);
          else v[j] *= silu_grad_f(h[j]);
        }
      }
      uint4 pk;
      pk.x = pack_bf16(v[0], v[1]);
      pk.y = pack_bf16(v[2], v[3]);
      pk.z = pack_bf16(v[4], v[5]);
      pk.w = pack_bf16(v[6], v[7]);
      *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + ro + c) = pk;
      if constexpr (EPI == EPI_GELU_TANH || EPI == EPI_GELU_ERF || EPI == EPI_SILU) {
        // activation of the bf16-rounded pre-activation (what the next Linear sees in the reference)
        const float2 q0 = unpack_bf16(pk.x), q1 = unpack_bf16(pk.y), q2 = unpack_bf16(pk.z), q3 = unpack_bf16(pk.w);
        const float h[8] = {q0.x, q0.y, q1.x, q1.y, q2.x, q2.y, q3.x, q3.y};
        float a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if constexpr (EPI == EPI_GELU_TANH) a[j] = gelu_tanh_f(h[j]);
          else if constexpr (EPI == EPI_GELU_ERF) a[j] = gelu_erf_f(h[j]);
          else a[j] = silu_f(h[j]);
        }
        uint4 ak;
        ak.x = pack_bf16(a[0], a[1]);
        ak.y = pack_bf16(a[2], a[3]);
        ak.z = pack_bf16(a[4], a[5]);
        ak.w = pack_bf16(a[6], a[7]);
        *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out2) + ro + c) = ak;
      }
      if constexpr (EPI == EPI_GATE_RES) {
        const float2 q0 = unpack_bf16(pk.x), q1 = unpack_bf16(pk.y), q2 = unpack_bf16(pk.z), q3 = unpack_bf16(pk.w);
        const float y[8] = {q0.x, q0.y, q1.x, q1.y, q2.x, q2.y, q3.x, q3.y};
        const float* gp = p.gate + (long long)(row / p.rows_per_sample) * p.ldg + c;
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gp));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gp) + 1);
        const float4* r = reinterpret_cast<const float4*>(p.resid + rr + c);
        const float4 r0 = r[0], r1 = r[1];
        float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out2) + ro + c);
        o[0] = make_float4(fmaf(g0.x, y[0], r0.x), fmaf(g0.y, y[1], r0.y), fmaf(g0.z, y[2], r0.z), fmaf(g0.w, y[3], r0.w));
        o[1] = make_float4(fmaf(g1.x, y[4], r1.x), fmaf(g1.y, y[5], r1.y), fmaf(g1.z, y[6], r1.z), fmaf(g1.w, y[7], r1.w));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// the GEMM kernel
// ---------------------------------------------------------------------------------------------------
template <int BN>
struct Cfg {
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 192 ? 5 : 6);
  static constexpr int kTmemCols = (2 * BN <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + 256 /*barriers*/ + 1024 /*alignment slack*/;
};

template <int BN, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const EpiParams p) {
  using C = Cfg<BN>;
  constexpr int STAGES = C::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * C::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (p.M + BM - 1) / BM;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb_total = (p.K + BK - 1) / BK;
  const int num_work = num_tiles * p.k_splits;  // work item w: tile = w % num_tiles, split = w / num_tiles

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 4);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int work = blockIdx.x; work < num_work; work += gridDim.x) {
        const int tile = work % num_tiles, split = work / num_tiles;
        const int m0 = (tile / n_tiles) * BM;
        const int n0 = (tile % n_tiles) * BN;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(num_kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1u);
          mbar_expect_tx(&full[stage], C::kStageBytes);
          uint8_t* a_dst = sA + stage * kABytes;
          uint8_t* b_dst = sB + stage * C::kBBytes;
          const int k0 = kb * BK;
          if (!p.a_mn) {
            tma_load_2d(a_dst, &tmA, &full[stage], k0, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d(a_dst + j * 8192, &tmA, &full[stage], m0 + j * 64, k0);
          }
          if (!p.b_mn) {
            tma_load_2d(b_dst, &tmB, &full[stage], k0, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(b_dst + j * 8192, &tmB, &full[stage], n0 + j * 64, k0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      // instruction descriptor: D=f32, A=B=bf16, majorness bits, N>>3, M>>4
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.a_mn ? 1 : 0) << 15) |
                             ((uint32_t)(p.b_mn ? 1 : 0) << 16) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(BM >> 4) << 24);
      const uint32_t a_kstep = p.a_mn ? 2048u : 32u;  // bytes per UMMA_K=16 step
      const uint32_t b_kstep = p.b_mn ? 2048u : 32u;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int work = blockIdx.x; work < num_work; work += gridDim.x) {
        const int split = work / num_tiles;
        const int kb0 = split * p.kb_per_split;
        const int num_kb = min(num_kb_total, kb0 + p.kb_per_split) - kb0;
        mbar_wait(&tempty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA + stage * kABytes);
          const uint32_t b_base = smem_u32(sB + stage * C::kBBytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t ad = umma_desc(a_base + k * a_kstep, p.a_mn);
            const uint64_t bd = umma_desc(b_base + k * b_kstep, p.b_mn);
            tc_mma_f16(d_tmem, ad, bd, idesc, (kb | k) ? 1u : 0u);
          }
          tc_commit(&empty[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit(&tfull[acc]);  // accumulator complete -> epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ===================== epilogue warps (2..5) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int work = blockIdx.x; work < num_work; work += gridDim.x) {
      const int tile = work % num_tiles, split = work / num_tiles;
      const int m0 = (tile / n_tiles) * BM;
      const int n0 = (tile % n_tiles) * BN;
      const long long slab = (long long)split * p.split_stride;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(t_row + (uint32_t)(c * 32), v);
        tmem_ld_wait();
        epilogue_chunk<EPI>(p, row, n0 + c * 32, v, slab);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------
// host side: tensor maps + dispatch
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

// 2-D bf16 tensor map over a row-major [rows, cols] matrix with leading dimension ld (elements);
// box = {64 contiguous elements (128 B, swizzled), box_rows rows}
int make_tmap(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) {
    vaw_set_error("cuTensorMapEncodeTiled entry point not available");
    return VAW_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    vaw_set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld box_rows=%d base=%p", (int)r, rows,
                  cols, ld, box_rows, base);
    return VAW_ERR_CUDA;
  }
  return VAW_OK;
}

// split-K finish: out[r, c] = (accumulate ? out : 0) + bias[c] + sum_s ws[s][r, c]   (fixed order -> deterministic)
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ ws, int splits, long long stride, float* __restrict__ out,
                     const float* __restrict__ bias, int M, int N, long long ldo, int accumulate) {
  const long long n4 = (long long)M * (N >> 2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / (N >> 2)), c = (int)(i % (N >> 2)) * 4;
    const long long o = (long long)r * ldo + c;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < splits; ++s) {
      const float4 v = *reinterpret_cast<const float4*>(ws + (long long)s * stride + o);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    if (bias) {
      const float4 b = *reinterpret_cast<const float4*>(bias + c);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    float4* dst = reinterpret_cast<float4*>(out + o);
    if (accumulate) {
      const float4 d = *dst;
      a.x += d.x; a.y += d.y; a.z += d.z; a.w += d.w;
    }
    *dst = a;
  }
}

template <int BN, int EPI>
int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const EpiParams& p, cudaStream_t stream) {
  using C = Cfg<BN>;
  static bool configured = false;
  auto kern = gemm_bf16_tcgen05_kernel<BN, EPI>;
  if (!configured) {
    VAW_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    configured = true;
  }
  const int m_tiles = (p.M + BM - 1) / BM, n_tiles = (p.N + BN - 1) / BN;
  int grid = m_tiles * n_tiles * (p.k_splits > 0 ? p.k_splits : 1);
  const int sms = vaw_num_sms();
  if (grid > sms) grid = sms;
  kern<<<grid, kThreads, C::kSmemBytes, stream>>>(tmA, tmB, p);
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

template <int BN>
int dispatch_epi(int epi, const CUtensorMap& tmA, const CUtensorMap& tmB, const EpiParams& p, cudaStream_t s) {
  switch (epi) {
    case EPI_BF16: return launch_gemm<BN, EPI_BF16>(tmA, tmB, p, s);
    case EPI_F32: return launch_gemm<BN, EPI_F32>(tmA, tmB, p, s);
    case EPI_GELU_TANH: return launch_gemm<BN, EPI_GELU_TANH>(tmA, tmB, p, s);
    case EPI_GELU_ERF: return launch_gemm<BN, EPI_GELU_ERF>(tmA, tmB, p, s);
    case EPI_GATE_RES: return launch_gemm<BN, EPI_GATE_RES>(tmA, tmB, p, s);
    case EPI_RES: return launch_gemm<BN, EPI_RES>(tmA, tmB, p, s);
    case EPI_DGELU_TANH: return launch_gemm<BN, EPI_DGELU_TANH>(tmA, tmB, p, s);
    case EPI_DGELU_ERF: return launch_gemm<BN, EPI_DGELU_ERF>(tmA, tmB, p, s);
    case EPI_SILU: return launch_gemm<BN, EPI_SILU>(tmA, tmB, p, s);
    case EPI_DSILU: return launch_gemm<BN, EPI_DSILU>(tmA, tmB, p, s);
  }
  vaw_set_error("vaw_gemm_bf16: unknown epilogue %d", epi);
  return VAW_ERR_INVALID;
}

}  // namespace

extern "C" int vaw_gemm_bf16(const vaw_gemm_args* a, cudaStream_t stream) {
  VAW_CHECK_ARG(a && a->A && a->B, "vaw_gemm_bf16: null operand");
  VAW_CHECK_ARG(a->M > 0 && a->N > 0 && a->K > 0, "vaw_gemm_bf16: bad shape M=%d N=%d K=%d", a->M, a->N, a->K);
  VAW_CHECK_ARG(a->N % 8 == 0, "vaw_gemm_bf16: N=%d must be a multiple of 8", a->N);
  VAW_CHECK_ARG(a->lda % 8 == 0 && a->ldb % 8 == 0, "vaw_gemm_bf16: lda/ldb must be multiples of 8 elements");
  VAW_CHECK_ARG((((uintptr_t)a->A | (uintptr_t)a->B) & 15) == 0, "vaw_gemm_bf16: operands must be 16-byte aligned");
  VAW_CHECK_ARG(a->epilogue >= 0 && a->epilogue < EPI_COUNT, "vaw_gemm_bf16: unknown epilogue %d", a->epilogue);
  const int epi = a->epilogue;
  const long long ldo = a->ldo ? a->ldo : a->N;
  VAW_CHECK_ARG(ldo % 8 == 0, "vaw_gemm_bf16: ldo must be a multiple of 8");
  const bool needs_out = (epi != EPI_RES);
  VAW_CHECK_ARG(!needs_out || a->out, "vaw_gemm_bf16: missing out");
  const bool needs_out2 = (epi == EPI_GELU_TANH || epi == EPI_GELU_ERF || epi == EPI_GATE_RES || epi == EPI_RES ||
                           epi == EPI_SILU);
  VAW_CHECK_ARG(!needs_out2 || a->out2, "vaw_gemm_bf16: missing out2");
  VAW_CHECK_ARG(!(epi == EPI_GATE_RES || epi == EPI_RES) || a->resid, "vaw_gemm_bf16: missing resid");
  VAW_CHECK_ARG(epi != EPI_GATE_RES || (a->gate && a->rows_per_sample > 0), "vaw_gemm_bf16: missing gate");
  VAW_CHECK_ARG(!(epi == EPI_DGELU_TANH || epi == EPI_DGELU_ERF || epi == EPI_DSILU) || a->aux,
                "vaw_gemm_bf16: missing aux");

  int bn = a->tile_n;
  if (bn == 0) {
    if (a->N % 192 == 0) bn = 192;
    else if (a->N % 256 == 0) bn = 256;
    else if (a->N % 128 == 0) bn = 128;
    else bn = (a->N > 128) ? 192 : 128;
  }
  VAW_CHECK_ARG(bn == 128 || bn == 192 || bn == 256, "vaw_gemm_bf16: tile_n must be 128, 192 or 256");

  CUtensorMap tmA, tmB;
  int rc;
  if (!a->a_mn) rc = make_tmap(&tmA, a->A, a->M, a->K, a->lda, BM);
  else rc = make_tmap(&tmA, a->A, a->K, a->M, a->lda, BK);
  if (rc) return rc;
  if (!a->b_mn) rc = make_tmap(&tmB, a->B, a->N, a->K, a->ldb, bn);
  else rc = make_tmap(&tmB, a->B, a->K, a->N, a->ldb, BK);
  if (rc) return rc;

  EpiParams p;
  p.out = a->out;
  p.out2 = a->out2;
  p.bias = a->bias;
  p.resid = a->resid;
  p.gate = a->gate;
  p.aux = reinterpret_cast<const bf16*>(a->aux);
  p.ldo = ldo;
  p.ldg = a->ldg ? a->ldg : a->N;
  p.rows_per_sample = a->rows_per_sample > 0 ? a->rows_per_sample : 1;
  p.accumulate = a->accumulate;
  p.resid_mod = a->resid_mod;
  p.M = a->M;
  p.N = a->N;
  p.K = a->K;
  p.a_mn = a->a_mn ? 1 : 0;
  p.b_mn = a->b_mn ? 1 : 0;
  // split-K (EPI_F32 only): partial slabs in split_ws, then a fixed-order reduction that applies bias / accumulate
  const int num_kb = (a->K + BK - 1) / BK;
  int splits = a->k_splits > 1 ? a->k_splits : 1;
  if (splits > num_kb) splits = num_kb;
  int kb_per = (num_kb + splits - 1) / splits;
  splits = (num_kb + kb_per - 1) / kb_per;
  p.k_splits = splits;
  p.kb_per_split = kb_per;
  p.split_stride = (long long)a->M * ldo;
  if (splits > 1) {
    VAW_CHECK_ARG(epi == EPI_F32 && a->split_ws, "vaw_gemm_bf16: split-K needs the F32 epilogue and split_ws");
    p.out = a->split_ws;
    p.bias = nullptr;
    p.accumulate = 0;
    int rc2;
    switch (bn) {
      case 128: rc2 = dispatch_epi<128>(epi, tmA, tmB, p, stream); break;
      case 192: rc2 = dispatch_epi<192>(epi, tmA, tmB, p, stream); break;
      default: rc2 = dispatch_epi<256>(epi, tmA, tmB, p, stream); break;
    }
    if (rc2) return rc2;
    const long long n4 = (long long)a->M * (a->N / 4);
    long long blocks = (n4 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    VAW_CHECK_ARG(a->N % 4 == 0, "vaw_gemm_bf16: split-K needs N %% 4 == 0");
    splitk_reduce_kernel<<<(unsigned)blocks, 256, 0, stream>>>(a->split_ws, splits, p.split_stride, (float*)a->out,
                                                              a->bias, a->M, a->N, ldo, a->accumulate);
    VAW_LAUNCH_CHECK();
    return VAW_OK;
  }

  switch (bn) {
    case 128: return dispatch_epi<128>(epi, tmA, tmB, p, stream);
    case 192: return dispatch_epi<192>(epi, tmA, tmB, p, stream);
    default: return dispatch_epi<256>(epi, tmA, tmB, p, stream);
  }
}

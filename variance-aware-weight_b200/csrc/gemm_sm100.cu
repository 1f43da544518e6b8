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
// Kernel shape (persistent, 320 threads per CTA):
//   warp 0      TMA producer   (cp.async.bulk.tensor -> STAGES x {A, B} ring, mbarrier full/empty)
//   warp 1      MMA issuer     (the warp walks the loop with uniform descriptors, lane 0 issues tcgen05.mma and commits)
//   warps 2-9   epilogue       (tcgen05.ld -> registers -> fused math -> swizzled staging boxes -> TMA store for the
//                               bf16 / residual epilogues, with TMA loads of their tile inputs; the fp32
//                               weight-gradient and split-K paths transpose through smem for coalesced stores)
// Two accumulator stages in TMEM (2 x BN columns) let the epilogue of tile i overlap the MMAs of tile i+1.
// Sustained runs are power-capped (scripts/dev_gemm_power.py): the epilogues are written for few instructions and
// little data movement, because under the cap kernel time follows energy, not the critical path.
//
// PAIR = true: two CTAs of one cluster (an SM pair) work on a 256 x BN tile with tcgen05.mma.cta_group::2:
// each CTA stages its own 128 rows of A and HALF of the B tile, the leader CTA issues the MMAs for both, and the
// accumulator rows are split over the two CTAs' TMEM.  Per SM this halves the B traffic through shared memory and
// L2, which is what bounds the single-CTA kernel (128x256x16 per 128 cycles needs 96 B/cycle of operand reads plus
// the same again of TMA writes against a 128 B/cycle shared-memory port).
#include "vaw_common.cuh"
#include "vaw_async.cuh"
#include "vaw_internal.h"
#include <stdlib.h>

namespace {

constexpr int BK = 64;
constexpr int kEpiWarps = 8;
constexpr int kEpiParts = kEpiWarps / 4;      // warps sharing one TMEM lane quarter; chunk c belongs to part c % kEpiParts
constexpr int kThreads = 64 + 32 * kEpiWarps;  // TMA warp + MMA warp + epilogue warps
constexpr int kABytes = 128 * BK * 2;          // 16 KB per stage per CTA

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
  EPI_ALIGN_MSE = 10,  // out_bf16 = zs = acc + bias ; out2_f32[(row / 32) * ceil(N / 32) + col / 32] = sum over the
                       // 32 x 32 block of (zs - aux)^2: the REPA alignment loss accumulated where zs is produced
  EPI_COUNT = 11
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
  int n_out;        // EPI_F32: columns [0, n_out) go to out; column n_out (if < N) is a row sum that goes to out2[row]
                    // (B = [X | 1 0 ... 0]: weight gradient and bias gradient from one GEMM), columns above are dropped
  int a_mn, b_mn;   // operand majorness: 0 = K-major, 1 = MN-major
  // work decomposition (see decode_work): full tiles, then the remaining tiles split along K
  int num_work, full_tiles, tail_splits, kb_per_split;
  float* split_ws;
  int tma_res;  // residual epilogues: resid / out2 / out move through TMA (set by the host when the layout allows)
  int dbg;  // debug knobs (env VAW_DBG): 1 = ring of 2 stages, 2 = skip MMA issue, 4 = skip TMA issue,
            // 8 = epilogue drains TMEM only (| 64: plus (dbg >> 8) x 32 FFMAs of pure ALU work per chunk),
            // 32 = bf16 epilogues do all their math but skip the global stores
};

// ---------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // shared::cluster address of the same offset in the pair's leader CTA
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}

template <bool PAIR>
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  if constexpr (!PAIR) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
  } else {
    // data lands in this CTA's shared memory; the transaction bytes are credited to the leader CTA's barrier
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
        "[%2];" ::"r"(smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
        : "memory");
  }
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

template <bool PAIR>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_smem, uint32_t ncols) {
  if constexpr (!PAIR) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <bool PAIR>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if constexpr (!PAIR)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// MMA completion -> mbarrier.  PAIR: the arrive is multicast to the barrier at this offset in both CTAs of the pair.
template <bool PAIR>
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  if constexpr (!PAIR) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
  } else {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"((uint16_t)3)
        : "memory");
  }
}
template <bool PAIR>
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  if constexpr (!PAIR) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
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
// fused epilogue on 4 consecutive columns of one row.  The kernel transposes each 32x32 accumulator chunk through
// shared memory first, so a warp touches 4 rows x 128 contiguous bytes per instruction: every global load / store of
// the epilogue (residual, gate, saved pre-activation, outputs) is fully coalesced.
// ---------------------------------------------------------------------------------------------------
struct EpiPre {  // global operands of the epilogue, fetched for a whole chunk before any dependent math / store
  float4 r;      // residual (RES / GATE_RES) or the previous output (F32 accumulate)
  float4 g;      // gate (GATE_RES)
  uint2 a;       // saved pre-activation (D-activation epilogues)
};
template <int EPI>
__device__ __forceinline__ void epilogue_load(const EpiParams& p, int row, int col, long long gate_off, bool interior,
                                              EpiPre& pre) {
  if (!interior && (row >= p.M || col >= p.N)) return;
  const long long o = (long long)row * p.ldo + col;
  if constexpr (EPI == EPI_F32) {
    if (col >= p.n_out) {   // the row-sum column (and the padding after it)
      if (p.accumulate && col == p.n_out && row < p.M) pre.r.x = reinterpret_cast<const float*>(p.out2)[row];
      return;
    }
    if (p.accumulate) pre.r = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.out) + o);
  }
  if constexpr (EPI == EPI_RES || EPI == EPI_GATE_RES) {
    const long long rr = (long long)(p.resid_mod > 0 ? row % p.resid_mod : row) * p.ldo + col;
    pre.r = *reinterpret_cast<const float4*>(p.resid + rr);
  }
  if constexpr (EPI == EPI_GATE_RES) pre.g = __ldg(reinterpret_cast<const float4*>(p.gate + gate_off + col));
  if constexpr (EPI == EPI_DGELU_TANH || EPI == EPI_DGELU_ERF || EPI == EPI_DSILU)
    pre.a = __ldg(reinterpret_cast<const uint2*>(p.aux + o));
}

template <int EPI>
// (the bias is added by the caller: a lane's four columns are the same for all eight rows of a chunk)
__device__ __forceinline__ void epilogue_vec4(const EpiParams& p, int row, int col, float4 v, bool interior,
                                              const EpiPre& pre) {
  if (!interior && (row >= p.M || col >= p.N)) return;
  const long long o = (long long)row * p.ldo + col;
  if constexpr (EPI == EPI_F32) {
    if (col >= p.n_out) {
      if (col == p.n_out && row < p.M) reinterpret_cast<float*>(p.out2)[row] = p.accumulate ? pre.r.x + v.x : v.x;
      return;
    }
    float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + o);
    if (p.accumulate) {
      const float4 d = pre.r;
      v.x += d.x; v.y += d.y; v.z += d.z; v.w += d.w;
    }
    *dst = v;
  } else if constexpr (EPI == EPI_RES) {
    const float4 r = pre.r;
    // the linear output is a bf16 tensor in the reference's autocast path: round before the residual add
    const float2 y0 = unpack_bf16(pack_bf16(v.x, v.y)), y1 = unpack_bf16(pack_bf16(v.z, v.w));
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out2) + o) = make_float4(r.x + y0.x, r.y + y0.y, r.z + y1.x, r.w + y1.y);
  } else {
    if constexpr (EPI == EPI_DGELU_TANH || EPI == EPI_DGELU_ERF || EPI == EPI_DSILU) {
      const float2 h0 = unpack_bf16(pre.a.x), h1 = unpack_bf16(pre.a.y);
      if constexpr (EPI == EPI_DGELU_TANH) {
        const float2 g0 = gelu_tanh_grad_f2(h0), g1 = gelu_tanh_grad_f2(h1);
        const float2 r0 = __fmul2_rn(make_float2(v.x, v.y), g0), r1 = __fmul2_rn(make_float2(v.z, v.w), g1);
        v = make_float4(r0.x, r0.y, r1.x, r1.y);
      } else if constexpr (EPI == EPI_DGELU_ERF) {
        v.x *= gelu_erf_grad_f(h0.x); v.y *= gelu_erf_grad_f(h0.y); v.z *= gelu_erf_grad_f(h1.x); v.w *= gelu_erf_grad_f(h1.y);
      } else {
        v.x *= silu_grad_f(h0.x); v.y *= silu_grad_f(h0.y); v.z *= silu_grad_f(h1.x); v.w *= silu_grad_f(h1.y);
      }
    }
    uint2 pk;
    pk.x = pack_bf16(v.x, v.y);
    pk.y = pack_bf16(v.z, v.w);
    const bool dry = (p.dbg & 32) != 0;   // experiment: all the math, no global stores
    // p.out may be null for the epilogues with a second output (forward-only use: nobody reads the saved pre-activation /
    // branch output)
    if (p.out && (!dry || pk.x == 0x7fc17fc2u)) *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(p.out) + o) = pk;
    if constexpr (EPI == EPI_GELU_TANH || EPI == EPI_GELU_ERF || EPI == EPI_SILU) {
      // activation of the bf16-rounded pre-activation (what the next Linear sees in the reference)
      const float2 h0 = unpack_bf16(pk.x), h1 = unpack_bf16(pk.y);
      uint2 ak;
      if constexpr (EPI == EPI_GELU_TANH) {
        const float2 a0 = gelu_tanh_f2(h0), a1 = gelu_tanh_f2(h1);
        ak.x = pack_bf16(a0.x, a0.y); ak.y = pack_bf16(a1.x, a1.y);
      } else if constexpr (EPI == EPI_GELU_ERF) {
        ak.x = pack_bf16(gelu_erf_f(h0.x), gelu_erf_f(h0.y)); ak.y = pack_bf16(gelu_erf_f(h1.x), gelu_erf_f(h1.y));
      } else {
        ak.x = pack_bf16(silu_f(h0.x), silu_f(h0.y)); ak.y = pack_bf16(silu_f(h1.x), silu_f(h1.y));
      }
      if (!dry || ak.x == 0x7fc17fc2u) *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(p.out2) + o) = ak;
    }
    if constexpr (EPI == EPI_GATE_RES) {
      const float2 y0 = unpack_bf16(pk.x), y1 = unpack_bf16(pk.y);
      const float4 g = pre.g;
      const float4 r = pre.r;
      const float2 o0 = __ffma2_rn(make_float2(g.x, g.y), y0, make_float2(r.x, r.y));
      const float2 o1 = __ffma2_rn(make_float2(g.z, g.w), y1, make_float2(r.z, r.w));
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out2) + o) = make_float4(o0.x, o0.y, o1.x, o1.y);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// the GEMM kernel
// ---------------------------------------------------------------------------------------------------
template <int BN, bool PAIR>
struct Cfg {
  static constexpr int kTileM = PAIR ? 256 : 128;
  static constexpr int kBRows = PAIR ? BN / 2 : BN;       // rows (K-major) / columns (MN-major) of B staged per CTA
  static constexpr int kBBoxes = (kBRows + 63) / 64;      // MN-major B: 64-wide TMA boxes per stage
  static constexpr int kBBytes = kBBoxes * 8192;          // smem reserved per stage for B (>= kBRows * 128)
  static constexpr int kBBytesKMajor = kBRows * BK * 2;   // bytes a K-major B load actually transfers
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingBytes = kEpiWarps * 4096;  // one 32x32 fp32 transpose tile per epilogue warp
  static constexpr int kBudget = 232448 - 1024 - 256 - kStagingBytes;
  static constexpr int kStagesRaw = kBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTmemCols = (2 * BN <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + 256 /*barriers*/ + kStagingBytes + 1024 /*alignment slack*/;
};

// One unit of work of the persistent loop.  The first `full_tiles` items are whole output tiles; the remaining
// tiles (the partial last wave) are each split into `tail_splits` K-ranges whose raw fp32 partial accumulators go
// to a scratch slab and are folded by splitk_fixup_kernel in fixed order (deterministic).
struct Work {
  int m0, n0, kb0, kb1, partial, slab;
};
template <int TM, int BN>
__device__ __forceinline__ Work decode_work(int w, const EpiParams& p, int n_tiles, int num_kb_total) {
  Work k;
  int tile;
  if (w < p.full_tiles) {
    tile = w;
    k.kb0 = 0;
    k.kb1 = num_kb_total;
    k.partial = 0;
    k.slab = 0;
  } else {
    const int r = w - p.full_tiles;
    tile = p.full_tiles + r / p.tail_splits;
    const int sp = r % p.tail_splits;
    k.kb0 = sp * p.kb_per_split;
    k.kb1 = min(num_kb_total, k.kb0 + p.kb_per_split);
    k.partial = 1;
    k.slab = r;
  }
  k.m0 = (tile / n_tiles) * TM;
  k.n0 = (tile % n_tiles) * BN;
  return k;
}

// Tensor maps of the bf16 outputs for the TMA-store epilogue ([32 rows x 32 columns] boxes, SWIZZLE_64B).
struct OutMaps {
  CUtensorMap c, c2, c3;   // bf16 out | bf16 out2 or aux, fp32 out2 of the residual epilogues | fp32 resid
};
// Epilogues whose only per-element input is the accumulator (plus a per-column bias) and whose outputs are bf16 take the
// TMA-store path: the math runs in the TMEM row layout (thread = row, 32 consecutive columns), the bf16 results go
// into a swizzled shared-memory box and leave through the bulk-store engine - no fp32 transpose, no per-thread global
// stores or address arithmetic.  (Sustained runs are power-capped, so the epilogue's instruction and data-movement
// energy, not its latency, is what it costs: scripts/dev_gemm_power.py.)
template <int EPI>
struct TmaEpi {
  // aux: the d-activation epilogues also READ a bf16 tile (the saved pre-activation); it arrives by TMA in the same box
  // geometry, one chunk ahead of the math (om.c2 is then the map of that input)
  static constexpr bool aux = (EPI == EPI_DGELU_TANH || EPI == EPI_DGELU_ERF || EPI == EPI_DSILU ||
                               EPI == EPI_ALIGN_MSE);
  static constexpr bool value =
      (EPI == EPI_BF16 || EPI == EPI_GELU_TANH || EPI == EPI_GELU_ERF || EPI == EPI_SILU) || aux;
  static constexpr bool two = (EPI == EPI_GELU_TANH || EPI == EPI_GELU_ERF || EPI == EPI_SILU);
};
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}

// Columns of a tile that exist (the last tile of a row of tiles may be ragged: N = 1152 = 4.5 x 256).  The MMAs of such
// a tile run with N = tile_cols() instead of BN, and the epilogue skips the missing 32-column chunks: a tenth of the
// tensor-core work (and energy) of every N = 1152 GEMM was spent on zero padding.  Multiple of 32 so that each CTA of a
// pair holds a multiple of 16 columns of B.
__device__ __forceinline__ int tile_cols(int N, int n0, int BN_) {
  const int left = (N - n0 + 31) & ~31;
  return left < BN_ ? left : BN_;
}

template <int BN, int EPI, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ OutMaps om, const EpiParams p) {
  using C = Cfg<BN, PAIR>;
  constexpr int STAGES = C::kStages;
  constexpr int TM = C::kTileM;
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
  uint64_t* abar = bars + 2 * STAGES + 5;   // [kEpiWarps] per-warp barriers of the TMA-loaded aux tiles

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;      // 0 = leader of the CTA pair
  const int unit = PAIR ? (blockIdx.x >> 1) : blockIdx.x;   // persistent work is dealt to CTAs / CTA pairs
  const int num_units = PAIR ? (gridDim.x >> 1) : gridDim.x;

  const int n_tiles = (p.N + BN - 1) / BN;
  const int num_kb_total = (p.K + BK - 1) / BK;
  const int num_work = p.num_work;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      // one arrival: the (leader's) producer arrive.expect_tx.  In PAIR mode it expects the bytes of BOTH CTAs; the
      // peer's TMA credits them to this barrier directly.  (A release-arrive from the peer after its TMA issue
      // waited for those loads to land and serialised the ring to depth 1: 1 us per k-block, measured.)
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], PAIR ? 2 * kEpiWarps : kEpiWarps);
    }
    for (int w = 0; w < kEpiWarps; ++w) mbar_init(&abar[w], 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tmem_alloc<PAIR>(tmem_slot, C::kTmemCols);
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer (every CTA stages its own rows of A and its share of B) ===============
      int stage = 0;
      uint32_t phase = 0;
      const int ring = (p.dbg & 1) ? 2 : STAGES;
      const bool skip_tma = (p.dbg & 4) != 0;
      const uint32_t bbytes = p.b_mn ? (uint32_t)C::kBBytes : (uint32_t)C::kBBytesKMajor;
      for (int work = unit; work < num_work; work += num_units) {
        const Work wk = decode_work<TM, BN>(work, p, n_tiles, num_kb_total);
        const int am0 = wk.m0 + (int)rank * 128;
        // a CTA pair splits the tile's columns of B in two; on a ragged tile the split point moves with it
        const int bn0 = wk.n0 + (PAIR ? (int)rank * (tile_cols(p.N, wk.n0, BN) >> 1) : 0);
        for (int kb = wk.kb0; kb < wk.kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1u);
          if (rank == 0) mbar_expect_tx(&full[stage], skip_tma ? 0u : (PAIR ? 2u : 1u) * (kABytes + bbytes));
          uint8_t* a_dst = sA + stage * kABytes;
          uint8_t* b_dst = sB + stage * C::kBBytes;
          const int k0 = kb * BK;
          if (skip_tma) {
          } else if (!p.a_mn) {
            tma_load_2d<PAIR>(a_dst, &tmA, &full[stage], k0, am0);
          } else {
#pragma unroll
            for (int j = 0; j < 2; ++j) tma_load_2d<PAIR>(a_dst + j * 8192, &tmA, &full[stage], am0 + j * 64, k0);
          }
          if (skip_tma) {
          } else if (!p.b_mn) {
            tma_load_2d<PAIR>(b_dst, &tmB, &full[stage], k0, bn0);
          } else {
#pragma unroll
            for (int j = 0; j < C::kBBoxes; ++j) tma_load_2d<PAIR>(b_dst + j * 8192, &tmB, &full[stage], bn0 + j * 64, k0);
          }
          if (++stage == ring) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ===================== MMA issuer (leader CTA only) =====================
      // The whole warp walks the loop so that barriers, descriptors and TMEM addresses stay warp-uniform (uniform
      // datapath); only lane 0 issues.  With everything under `if (lane == 0)` every MMA dragged ~20 vector instructions
      // of descriptor arithmetic along.
      // instruction descriptor: D=f32, A=B=bf16, majorness bits, N>>3, M>>4
      const bool leader = lane == 0;
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.a_mn ? 1 : 0) << 15) |
                             ((uint32_t)(p.b_mn ? 1 : 0) << 16) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(TM >> 4) << 24);
      const uint32_t a_kstep = (p.a_mn ? 2048u : 32u) >> 4;  // descriptor units (16 B) per UMMA_K=16 step
      const uint32_t b_kstep = (p.b_mn ? 2048u : 32u) >> 4;
      const uint32_t tm_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const int ring = (p.dbg & 1) ? 2 : STAGES;
      const bool skip_mma = (p.dbg & 2) != 0;
      for (int work = unit; work < num_work; work += num_units) {
        const Work wk = decode_work<TM, BN>(work, p, n_tiles, num_kb_total);
        const int num_kb = wk.kb1 - wk.kb0;
        mbar_wait(&tempty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tm_u + (uint32_t)(acc * BN);
        const uint32_t idesc_t = (idesc & ~(0x3Fu << 17)) | ((uint32_t)(tile_cols(p.N, wk.n0, BN) >> 3) << 17);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t ad0 = umma_desc(smem_u32(sA + stage * kABytes), p.a_mn);
          const uint64_t bd0 = umma_desc(smem_u32(sB + stage * C::kBBytes), p.b_mn);
          if (leader) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              if (!skip_mma) tc_mma_f16<PAIR>(d_tmem, ad0 + k * a_kstep, bd0 + k * b_kstep, idesc_t, (kb | k) ? 1u : 0u);
            tc_commit<PAIR>(&empty[stage]);  // frees the smem slot (in both CTAs) once these MMAs have read it
          }
          __syncwarp();
          if (++stage == ring) { stage = 0; phase ^= 1u; }
        }
        if (leader) tc_commit<PAIR>(&tfull[acc]);  // accumulator complete -> epilogue warps (of both CTAs)
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ===================== epilogue warps =====================
    // kEpiParts warps per TMEM lane quarter; 32-column chunk c of the tile is drained by part c % kEpiParts.
    // (Sixteen epilogue warps were measured too: no better than eight once the epilogue math was trimmed, and the
    //  register cap then spills the residual epilogues.)
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int part = (warp - 2) >> 2;
    constexpr int NCH_MAX = BN / 32;
    float4* stg = reinterpret_cast<float4*>(smem + STAGES * C::kStageBytes + 256) + (warp - 2) * 256;
    int acc = 0;
    uint32_t acc_phase = 0, aux_phase = 0;
    // staging slot of the alternating single-output TMA-store epilogue.  It lives across work items: `wait_group.read 1`
    // only proves that the store issued two chunks ago is done with its slot, so a tile with an odd number of chunks
    // per warp must be followed by a tile that starts on the OTHER slot (with a short K loop the next tile's first
    // chunk otherwise overwrites a slot the store engine is still reading).
    int slot = 0;
    // after the transpose this lane owns rows {4i + lane/8} (i = 0..7) and column group lane%8 of every chunk
    const int sub_r = lane >> 3, sub_c = lane & 7;
    for (int work = unit; work < num_work; work += num_units) {
      const Work wk = decode_work<TM, BN>(work, p, n_tiles, num_kb_total);
      const int NCH = tile_cols(p.N, wk.n0, BN) >> 5;              // 32-column chunks that exist in this tile
      const int row0 = wk.m0 + (int)rank * 128 + q * 32 + sub_r;  // first global row of this lane (then +4 per i)
      if constexpr (TmaEpi<EPI>::aux) {
        // the first aux tile of this work item does not depend on the accumulator: fetch it before waiting for the MMAs
        // (slot 0 is free: the previous tile's last aux tile was consumed into registers)
        if (lane == 0 && part < NCH) {
          mbar_expect_tx(&abar[warp - 2], 2048);
          tma_load_2d<false>(reinterpret_cast<uint8_t*>(stg), &om.c2, &abar[warp - 2], wk.n0 + part * 32,
                             wk.m0 + (int)rank * 128 + q * 32);
        }
        __syncwarp();
      }
      constexpr bool kResEpi = (EPI == EPI_GATE_RES || EPI == EPI_RES);
      if constexpr (kResEpi) {
        if (p.tma_res && lane == 0 && part < NCH) {   // first 16-column residual tile of this work item
          mbar_expect_tx(&abar[warp - 2], 2048);
          tma_load_2d<false>(reinterpret_cast<uint8_t*>(stg), &om.c3, &abar[warp - 2], wk.n0 + part * 32,
                             wk.m0 + (int)rank * 128 + q * 32);
        }
        __syncwarp();
      }
      mbar_wait_sleep(&tfull[acc], acc_phase, 128);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      const bool interior = wk.m0 + TM <= p.M && wk.n0 + BN <= p.N;   // no bounds checks inside the tile
      if constexpr (TmaEpi<EPI>::value) {
        // ---- TMA-store epilogue: thread = accumulator row, one 32-column chunk at a time ----
        uint8_t* stg_b = reinterpret_cast<uint8_t*>(stg);          // 2 slots of [32 rows x 64 B], SWIZZLE_64B
        const int grow = wk.m0 + (int)rank * 128 + q * 32;            // first global row of this warp's 32 rows
        // SWIZZLE_64B is a function of the absolute shared-memory address (bits 4-5 ^= bits 7-8); the staging area is
        // only 256-byte aligned, so the row's swizzle term comes from its address, not from its index.  Both slots are
        // 2048 B apart, i.e. share the term.
        const int sw = (int)(((smem_u32(stg_b) + (uint32_t)lane * 64u) >> 7) & 3u);   // 16-byte chunk c lives at c ^ sw
        // aux variant: slot 0 receives the aux tile (TMA load, per-warp mbarrier), slot 1 stages the output
        uint64_t* my_bar = &abar[warp - 2];
        auto aux_load = [&](int c) {   // lane 0
          mbar_expect_tx(my_bar, 2048);
          tma_load_2d<false>(stg_b, &om.c2, my_bar, wk.n0 + c * 32, grow);
        };
#pragma unroll 1
        for (int c = part; c < NCH; c += kEpiParts) {
          uint32_t ax[16];
          if constexpr (TmaEpi<EPI>::aux) {
            mbar_wait(my_bar, aux_phase);
            aux_phase ^= 1u;
            const uint8_t* a0 = stg_b + lane * 64;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint4 t = *reinterpret_cast<const uint4*>(a0 + ((k ^ sw) * 16));
              ax[4 * k] = t.x; ax[4 * k + 1] = t.y; ax[4 * k + 2] = t.z; ax[4 * k + 3] = t.w;
            }
            fence_proxy_async_smem();   // order this lane's generic reads before the async-proxy (TMA) refill
            __syncwarp();   // every lane has its row: the next chunk's tile may overwrite the slot
            if (lane == 0 && c + kEpiParts < NCH) aux_load(c + kEpiParts);
          }
          // the slot(s) written below must have been read by the store engine (at most one older group may be in flight
          // for the alternating single-output epilogue, none otherwise)
          if (lane == 0) {
            if (TmaEpi<EPI>::two || TmaEpi<EPI>::aux) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          }
          __syncwarp();   // reconverge before the warp-collective TMEM load
          uint32_t v[32];
          tmem_ld32(t_row + (uint32_t)(c * 32), v);
          tmem_ld_wait();
          const int col0 = wk.n0 + c * 32;
          uint32_t pk[16], ak[16];
          [[maybe_unused]] float align_sq = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias && col0 + 4 * j < p.N) b = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + j);  // warp-uniform
            float2 lo = __fadd2_rn(make_float2(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])),
                                   make_float2(b.x, b.y));
            float2 hi = __fadd2_rn(make_float2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])),
                                   make_float2(b.z, b.w));
            if constexpr (EPI == EPI_ALIGN_MSE) {
              // zs leaves as bf16; the loss sees the rounded value, like F.mse_loss on the bf16 projector output
              const float2 f0 = unpack_bf16(ax[2 * j]), f1 = unpack_bf16(ax[2 * j + 1]);
              const float2 z0 = unpack_bf16(pack_bf16(lo.x, lo.y)), z1 = unpack_bf16(pack_bf16(hi.x, hi.y));
              if (col0 + 4 * j < p.N) {   // N % 4 == 0; columns past N hold bias-free zero padding on both sides anyway
                const float d0 = z0.x - f0.x, d1 = z0.y - f0.y, d2 = z1.x - f1.x, d3 = z1.y - f1.y;
                align_sq += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
              }
            } else if constexpr (TmaEpi<EPI>::aux) {   // dX = dY * act'(saved pre-activation)
              const float2 h0 = unpack_bf16(ax[2 * j]), h1 = unpack_bf16(ax[2 * j + 1]);
              if constexpr (EPI == EPI_DGELU_TANH) {
                lo = __fmul2_rn(lo, gelu_tanh_grad_f2(h0));
                hi = __fmul2_rn(hi, gelu_tanh_grad_f2(h1));
              } else if constexpr (EPI == EPI_DGELU_ERF) {
                lo = __fmul2_rn(lo, make_float2(gelu_erf_grad_f(h0.x), gelu_erf_grad_f(h0.y)));
                hi = __fmul2_rn(hi, make_float2(gelu_erf_grad_f(h1.x), gelu_erf_grad_f(h1.y)));
              } else {
                lo = __fmul2_rn(lo, make_float2(silu_grad_f(h0.x), silu_grad_f(h0.y)));
                hi = __fmul2_rn(hi, make_float2(silu_grad_f(h1.x), silu_grad_f(h1.y)));
              }
            }
            pk[2 * j] = pack_bf16(lo.x, lo.y);
            pk[2 * j + 1] = pack_bf16(hi.x, hi.y);
            if constexpr (TmaEpi<EPI>::two) {
              // activation of the bf16-rounded pre-activation (what the next Linear sees in the reference)
              const float2 h0 = unpack_bf16(pk[2 * j]), h1 = unpack_bf16(pk[2 * j + 1]);
              if constexpr (EPI == EPI_GELU_TANH) {
                const float2 a0 = gelu_tanh_f2(h0), a1 = gelu_tanh_f2(h1);
                ak[2 * j] = pack_bf16(a0.x, a0.y);
                ak[2 * j + 1] = pack_bf16(a1.x, a1.y);
              } else if constexpr (EPI == EPI_GELU_ERF) {
                ak[2 * j] = pack_bf16(gelu_erf_f(h0.x), gelu_erf_f(h0.y));
                ak[2 * j + 1] = pack_bf16(gelu_erf_f(h1.x), gelu_erf_f(h1.y));
              } else {
                ak[2 * j] = pack_bf16(silu_f(h0.x), silu_f(h0.y));
                ak[2 * j + 1] = pack_bf16(silu_f(h1.x), silu_f(h1.y));
              }
            }
          }
          if constexpr (EPI == EPI_ALIGN_MSE) {
            // one partial per (32-row group, 32-column chunk), written by exactly one warp: the fold order of the loss
            // does not depend on the tile configuration (deterministic)
            if (grow + lane >= p.M) align_sq = 0.f;
            align_sq = warp_sum(align_sq);
            if (lane == 0 && grow < p.M)
              reinterpret_cast<float*>(p.out2)[(long long)(grow >> 5) * ((p.N + 31) >> 5) + (col0 >> 5)] = align_sq;
          }
          if constexpr (TmaEpi<EPI>::aux) slot = 1;
          uint8_t* s0 = stg_b + (TmaEpi<EPI>::two ? 0 : slot * 2048) + lane * 64;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            *reinterpret_cast<uint4*>(s0 + ((k ^ sw) * 16)) = make_uint4(pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
          if constexpr (TmaEpi<EPI>::two) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              *reinterpret_cast<uint4*>(s0 + 2048 + ((k ^ sw) * 16)) =
                  make_uint4(ak[4 * k], ak[4 * k + 1], ak[4 * k + 2], ak[4 * k + 3]);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (!TmaEpi<EPI>::two || p.out) tma_store_2d(&om.c, stg_b + (TmaEpi<EPI>::two ? 0 : slot * 2048), col0, grow);
            if (TmaEpi<EPI>::two) tma_store_2d(&om.c2, stg_b + 2048, col0, grow);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          slot ^= 1;
        }
      } else if (kResEpi && p.tma_res) {
        // ---- residual epilogues through TMA: out2 = resid + gate * y (fp32), y = bf16(acc + bias) ----
        // 16-column halves (64-byte fp32 rows): slot 0 receives the residual tile one half ahead (per-warp mbarrier),
        // slot 1 stages out2 of each half and finally the bf16 y tile of the whole 32-column chunk.
        uint8_t* stg_b = reinterpret_cast<uint8_t*>(stg);
        const int grow = wk.m0 + (int)rank * 128 + q * 32;
        const int sw = (int)(((smem_u32(stg_b) + (uint32_t)lane * 64u) >> 7) & 3u);
        uint64_t* my_bar = &abar[warp - 2];
        const int my_row = min(grow + lane, p.M - 1);
        const float* gate_row = EPI == EPI_GATE_RES ? p.gate + (long long)(my_row / p.rows_per_sample) * p.ldg : nullptr;
        auto wait_slot1 = [&]() {
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
        };
#pragma unroll 1
        for (int c = part; c < NCH; c += kEpiParts) {
          const int col0 = wk.n0 + c * 32;
          uint32_t v[32], ypk[16];
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            mbar_wait(my_bar, aux_phase);
            aux_phase ^= 1u;
            float4 r[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) r[k] = *reinterpret_cast<const float4*>(stg_b + lane * 64 + ((k ^ sw) * 16));
            fence_proxy_async_smem();   // order this lane's generic reads before the async-proxy (TMA) refill
            __syncwarp();   // slot 0 consumed: fetch the next residual tile (other half, or the next chunk's first half)
            if (lane == 0) {
              const int nc = hf == 0 ? c : c + kEpiParts;
              if (nc < NCH) {
                mbar_expect_tx(my_bar, 2048);
                tma_load_2d<false>(stg_b, &om.c3, my_bar, wk.n0 + nc * 32 + (hf == 0 ? 16 : 0), grow);
              }
            }
            __syncwarp();
            if (hf == 0) {
              tmem_ld32(t_row + (uint32_t)(c * 32), v);
              tmem_ld_wait();
            }
            float4 o4[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int col = col0 + hf * 16 + 4 * k;
              float4 b = make_float4(0.f, 0.f, 0.f, 0.f), g = make_float4(1.f, 1.f, 1.f, 1.f);
              if (col < p.N) {
                if (p.bias) b = __ldg(reinterpret_cast<const float4*>(p.bias + col));
                if (EPI == EPI_GATE_RES) g = __ldg(reinterpret_cast<const float4*>(gate_row + col));
              }
              const int j = hf * 4 + k;
              const float2 lo = __fadd2_rn(make_float2(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])),
                                           make_float2(b.x, b.y));
              const float2 hi = __fadd2_rn(make_float2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])),
                                           make_float2(b.z, b.w));
              // the linear output is a bf16 tensor in the reference's autocast path: round before the residual add
              ypk[2 * j] = pack_bf16(lo.x, lo.y);
              ypk[2 * j + 1] = pack_bf16(hi.x, hi.y);
              const float2 y0 = unpack_bf16(ypk[2 * j]), y1 = unpack_bf16(ypk[2 * j + 1]);
              const float2 a0 = __ffma2_rn(make_float2(g.x, g.y), y0, make_float2(r[k].x, r[k].y));
              const float2 a1 = __ffma2_rn(make_float2(g.z, g.w), y1, make_float2(r[k].z, r[k].w));
              o4[k] = make_float4(a0.x, a0.y, a1.x, a1.y);
            }
            wait_slot1();
#pragma unroll
            for (int k = 0; k < 4; ++k) *reinterpret_cast<float4*>(stg_b + 2048 + lane * 64 + ((k ^ sw) * 16)) = o4[k];
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&om.c2, stg_b + 2048, col0 + hf * 16, grow);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          }
          if (p.out) {   // y (bf16) of the whole chunk, needed by the backward pass of the gated branch
            wait_slot1();
#pragma unroll
            for (int k = 0; k < 4; ++k)
              *reinterpret_cast<uint4*>(stg_b + 2048 + lane * 64 + ((k ^ sw) * 16)) =
                  make_uint4(ypk[4 * k], ypk[4 * k + 1], ypk[4 * k + 2], ypk[4 * k + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&om.c, stg_b + 2048, col0, grow);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          }
        }
      } else {
      long long gate_off[8];
      if constexpr (EPI == EPI_GATE_RES) {
#pragma unroll
        for (int i = 0; i < 8; ++i) gate_off[i] = (long long)((row0 + 4 * i) / p.rows_per_sample) * p.ldg;
      }
#pragma unroll 1
      for (int c = part; c < NCH; c += kEpiParts) {
        const int col = wk.n0 + c * 32 + sub_c * 4;
        // issue the chunk's global operand loads first: their latency overlaps the TMEM load and the transpose
        EpiPre pre[8];
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!wk.partial) {
          if (p.bias && col < p.N) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
#pragma unroll
          for (int i = 0; i < 8; ++i)
            epilogue_load<EPI>(p, row0 + 4 * i, col, EPI == EPI_GATE_RES ? gate_off[i] : 0, interior, pre[i]);
        }
        uint32_t v[32];
        tmem_ld32(t_row + (uint32_t)(c * 32), v);
        tmem_ld_wait();
        if (p.dbg & 8) {   // experiment: no shared-memory transpose, no global traffic (results are wrong)
          if (p.dbg & 64) {   // ... plus a pure-ALU load of (dbg >> 8) x 32 dependent-free FFMAs per chunk
            float a0 = __uint_as_float(v[0]), a1 = __uint_as_float(v[1]), a2 = __uint_as_float(v[2]), a3 = __uint_as_float(v[3]);
            float a4 = __uint_as_float(v[4]), a5 = __uint_as_float(v[5]), a6 = __uint_as_float(v[6]), a7 = __uint_as_float(v[7]);
            const int n = p.dbg >> 8;
            for (int it = 0; it < n; ++it) {
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                a0 = fmaf(a0, 1.0001f, 0.5f); a1 = fmaf(a1, 1.0001f, 0.5f); a2 = fmaf(a2, 1.0001f, 0.5f); a3 = fmaf(a3, 1.0001f, 0.5f);
                a4 = fmaf(a4, 1.0001f, 0.5f); a5 = fmaf(a5, 1.0001f, 0.5f); a6 = fmaf(a6, 1.0001f, 0.5f); a7 = fmaf(a7, 1.0001f, 0.5f);
              }
            }
            v[0] = __float_as_uint(a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7);
          }
          if (v[0] == 0x7fc12345u) stg[lane] = make_float4(0.f, 0.f, 0.f, 0.f);
          continue;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          stg[lane * 8 + (j ^ (lane & 7))] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                         __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = 4 * i + sub_r;
          float4 a4 = stg[r * 8 + (sub_c ^ (r & 7))];
          if (!wk.partial) {
            {
              const float2 lo = __fadd2_rn(make_float2(a4.x, a4.y), make_float2(bias4.x, bias4.y));
              const float2 hi = __fadd2_rn(make_float2(a4.z, a4.w), make_float2(bias4.z, bias4.w));
              a4 = make_float4(lo.x, lo.y, hi.x, hi.y);
            }
            epilogue_vec4<EPI>(p, row0 + 4 * i, col, a4, interior, pre[i]);
          } else {
            const int lrow = q * 32 + r, lcol = c * 32 + sub_c * 4;
            *reinterpret_cast<float4*>(p.split_ws + ((long long)wk.slab * TM + (int)rank * 128 + lrow) * BN + lcol) = a4;
          }
        }
        __syncwarp();
      }
      }  // generic (transposing) epilogue
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (!PAIR || rank == 0) mbar_arrive(&tempty[acc]);
        else mbar_arrive_remote(&tempty[acc], 0);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if constexpr (TmaEpi<EPI>::value || EPI == EPI_GATE_RES || EPI == EPI_RES) {
      // shared memory must outlive the store engine's reads
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  }

  __syncwarp();  // reconverge the single-lane role warps before the block / cluster barrier
  tc_fence_before();
  // Idle lanes / warps park in the hardware block barrier (bar.sync blocks without issuing).  The cluster barrier
  // polls, so it is only entered once the whole CTA is done: waiting in it from the start stole issue slots from
  // the single-lane TMA / MMA roles and serialised the pipeline (measured: 1 us per k-block).
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<PAIR>(tmem_base, C::kTmemCols);
  }
}

// Split-K fix-up (EPI_F32 only): one CTA per split tile sums its `splits` slabs in fixed order and applies
constexpr int kFixupSlices = 8;
// bias / accumulate:  out[r, c] = (accumulate ? out : 0) + bias[c] + sum_s slab[s][r, c]
template <int TM, int BN>
__global__ void __launch_bounds__(256)
splitk_fixup_kernel(const float* __restrict__ ws, int first_tile, int splits, int n_tiles, float* __restrict__ out,
                    const float* __restrict__ bias, int M, int N, long long ldo, int accumulate, int n_out,
                    float* __restrict__ out2) {
  const int tile = first_tile + blockIdx.x;
  const int m0 = (tile / n_tiles) * TM, n0 = (tile % n_tiles) * BN;
  const float* base = ws + (long long)blockIdx.x * splits * (TM * BN);
  constexpr int kRows = TM / kFixupSlices;   // a tile is folded by kFixupSlices CTAs (blockIdx.y), kRows rows each
  for (int i = threadIdx.x; i < kRows * BN / 4; i += 256) {
    const int r = blockIdx.y * kRows + i / (BN / 4), c = (i % (BN / 4)) * 4;
    if (m0 + r >= M || n0 + c >= N) continue;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < splits; ++s) {
      const float4 v = *reinterpret_cast<const float4*>(base + (long long)s * (TM * BN) + r * BN + c);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    if (n0 + c >= n_out) {   // the row-sum column of EPI_F32 (see EpiParams::n_out)
      if (n0 + c == n_out) out2[m0 + r] = accumulate ? out2[m0 + r] + a.x : a.x;
      continue;
    }
    if (bias) {
      const float4 b = *reinterpret_cast<const float4*>(bias + n0 + c);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    float4* dst = reinterpret_cast<float4*>(out + (long long)(m0 + r) * ldo + n0 + c);
    if (accumulate) {
      const float4 d = *dst;
      a.x += d.x; a.y += d.y; a.z += d.z; a.w += d.w;
    }
    *dst = a;
  }
}

// SMs the persistent GEMM grids may occupy.  Experiment knob (env VAW_GEMM_SMS): with a collective running next to the
// backward pass, NCCL's CTAs hold some SMs and a statically scheduled 148-CTA grid then runs in two waves.
int gemm_sms() {
  static int n = 0;
  if (n == 0) {
    n = vaw_num_sms();
    const char* e = getenv("VAW_GEMM_SMS");
    if (e && atoi(e) >= 2 && atoi(e) < n) n = atoi(e) & ~1;
  }
  return n;
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

// bf16 output [rows, cols] with leading dimension ld: boxes of 32 columns (64 B, SWIZZLE_64B) x 32 rows
// (fp32 = true: boxes of 16 fp32 columns - the same 64-byte rows)
int make_out_tmap(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, bool fp32 = false) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) {
    vaw_set_error("cuTensorMapEncodeTiled entry point not available");
    return VAW_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * (fp32 ? 4 : 2)};
  cuuint32_t box[2] = {fp32 ? 16u : 32u, 32u};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(map, fp32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    vaw_set_error("cuTensorMapEncodeTiled (output) failed (%d) rows=%lld cols=%lld ld=%lld base=%p", (int)r, rows, cols,
                  ld, base);
    return VAW_ERR_CUDA;
  }
  return VAW_OK;
}

template <int BN, int EPI, bool PAIR>
int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const OutMaps& om, const EpiParams& p,
                cudaStream_t stream) {
  using C = Cfg<BN, PAIR>;
  static bool configured = false;
  auto kern = gemm_bf16_tcgen05_kernel<BN, EPI, PAIR>;
  if (!configured) {
    VAW_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    configured = true;
  }
  const int sms = gemm_sms();
  if constexpr (!PAIR) {
    const int grid = p.num_work < sms ? p.num_work : sms;
    kern<<<grid, kThreads, C::kSmemBytes, stream>>>(tmA, tmB, om, p);
  } else {
    const int pairs = p.num_work < sms / 2 ? p.num_work : sms / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = C::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    VAW_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, om, p));
  }
  VAW_LAUNCH_CHECK();
  return VAW_OK;
}

template <int BN, bool PAIR>
int dispatch_epi(int epi, const CUtensorMap& tmA, const CUtensorMap& tmB, const OutMaps& om, const EpiParams& p,
                 cudaStream_t s) {
  switch (epi) {
    case EPI_BF16: return launch_gemm<BN, EPI_BF16, PAIR>(tmA, tmB, om, p, s);
    case EPI_F32: return launch_gemm<BN, EPI_F32, PAIR>(tmA, tmB, om, p, s);
    case EPI_GELU_TANH: return launch_gemm<BN, EPI_GELU_TANH, PAIR>(tmA, tmB, om, p, s);
    case EPI_GELU_ERF: return launch_gemm<BN, EPI_GELU_ERF, PAIR>(tmA, tmB, om, p, s);
    case EPI_GATE_RES: return launch_gemm<BN, EPI_GATE_RES, PAIR>(tmA, tmB, om, p, s);
    case EPI_RES: return launch_gemm<BN, EPI_RES, PAIR>(tmA, tmB, om, p, s);
    case EPI_DGELU_TANH: return launch_gemm<BN, EPI_DGELU_TANH, PAIR>(tmA, tmB, om, p, s);
    case EPI_DGELU_ERF: return launch_gemm<BN, EPI_DGELU_ERF, PAIR>(tmA, tmB, om, p, s);
    case EPI_SILU: return launch_gemm<BN, EPI_SILU, PAIR>(tmA, tmB, om, p, s);
    case EPI_DSILU: return launch_gemm<BN, EPI_DSILU, PAIR>(tmA, tmB, om, p, s);
    case EPI_ALIGN_MSE: return launch_gemm<BN, EPI_ALIGN_MSE, PAIR>(tmA, tmB, om, p, s);
  }
  vaw_set_error("vaw_gemm_bf16: unknown epilogue %d", epi);
  return VAW_ERR_INVALID;
}

int dispatch_tile(int bn, bool pair, int epi, const CUtensorMap& tmA, const CUtensorMap& tmB, const OutMaps& om,
                  const EpiParams& p, cudaStream_t s) {
  if (pair) {
    switch (bn) {
      case 128: return dispatch_epi<128, true>(epi, tmA, tmB, om, p, s);
      case 192: return dispatch_epi<192, true>(epi, tmA, tmB, om, p, s);
      default: return dispatch_epi<256, true>(epi, tmA, tmB, om, p, s);
    }
  }
  switch (bn) {
    case 128: return dispatch_epi<128, false>(epi, tmA, tmB, om, p, s);
    case 192: return dispatch_epi<192, false>(epi, tmA, tmB, om, p, s);
    default: return dispatch_epi<256, false>(epi, tmA, tmB, om, p, s);
  }
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
  // VAW_EPI_F32 with out2: the last 32 columns of B are not part of the [M, N - 32] output; column N - 32 of the product
  // (a row sum when B = [X | 1 0 ... 0]) goes to out2[M]
  const int n_out = (epi == EPI_F32 && a->out2) ? a->N - 32 : a->N;
  VAW_CHECK_ARG(n_out > 0 && (n_out == a->N || (n_out % 4 == 0 && !a->bias)),
                "vaw_gemm_bf16: the row-sum form of VAW_EPI_F32 needs N - 32 > 0, a multiple of 4, and no bias");
  const long long ldo = a->ldo ? a->ldo : n_out;
  VAW_CHECK_ARG(ldo % 8 == 0, "vaw_gemm_bf16: ldo must be a multiple of 8");
  // the epilogues with a second output may drop the first (the saved pre-activation / branch output only the backward
  // pass reads): forward-only callers pass out = NULL and save its HBM writes
  const bool needs_out = !(epi == EPI_RES || epi == EPI_GATE_RES || epi == EPI_GELU_TANH || epi == EPI_GELU_ERF ||
                           epi == EPI_SILU);
  VAW_CHECK_ARG(!needs_out || a->out, "vaw_gemm_bf16: missing out");
  const bool needs_out2 = (epi == EPI_GELU_TANH || epi == EPI_GELU_ERF || epi == EPI_GATE_RES || epi == EPI_RES ||
                           epi == EPI_SILU);
  VAW_CHECK_ARG(!needs_out2 || a->out2, "vaw_gemm_bf16: missing out2");
  VAW_CHECK_ARG(!(epi == EPI_GATE_RES || epi == EPI_RES) || a->resid, "vaw_gemm_bf16: missing resid");
  VAW_CHECK_ARG(epi != EPI_GATE_RES || (a->gate && a->rows_per_sample > 0), "vaw_gemm_bf16: missing gate");
  VAW_CHECK_ARG(!(epi == EPI_DGELU_TANH || epi == EPI_DGELU_ERF || epi == EPI_DSILU || epi == EPI_ALIGN_MSE) || a->aux,
                "vaw_gemm_bf16: missing aux");
  VAW_CHECK_ARG(a->cta_group >= 0 && a->cta_group <= 2, "vaw_gemm_bf16: cta_group must be 0 (auto), 1 or 2");

  // ---- tile selection -----------------------------------------------------------------------------------
  // Explicit tile_n / cta_group are honoured.  Otherwise pick the (tile_n, cta_group) pair with the smallest
  // modelled time: waves x (k-blocks x cost-per-k-block + per-tile overhead), with the per-k-block costs measured on
  // B200 (profiles/r01_gemm_tile_costs.md): the main loop is bound by shared-memory traffic (TMA writes + UMMA
  // operand reads), which is why an SM pair on a 256x256 tile (half the B traffic per SM) is the cheapest per FLOP.
  bool pair = a->cta_group == 2;
  int bn = a->tile_n;
  if (bn == 0 && a->cta_group == 0) {
    const int sms = gemm_sms();
    const int nkb = (a->K + BK - 1) / BK;
    const bool tail_ok = (a->k_splits == -1);
    double best = 1e30;
    const int cand_bn[4] = {192, 256, 128, 256};
    const bool cand_pair[4] = {false, false, false, true};
    const double cand_cost[4] = {0.43, 0.50, 0.355, 0.455};  // us per k-block per CTA (per CTA pair for the last)
    for (int i = 0; i < 4; ++i) {
      if (cand_pair[i] && a->M < 256) continue;
      const int tm = cand_pair[i] ? 256 : 128;
      const int units = cand_pair[i] ? sms / 2 : sms;
      const long long tiles = (long long)((a->M + tm - 1) / tm) * ((a->N + cand_bn[i] - 1) / cand_bn[i]);
      // measured exceptions to the model (scripts/dev_perf.py): the 10-byte-per-element residual epilogues on a short
      // K loop (proj: 74 vs 86 us) and weight gradients with less than one wave of pair tiles (wgrad proj: 61 vs 72 us)
      // are faster on single CTAs
      if (cand_pair[i] && (epi == EPI_GATE_RES || epi == EPI_RES) && nkb <= 24) continue;
      if (cand_pair[i] && tail_ok && tiles < units) continue;
      const long long fullw = tiles / units, rem = tiles % units;
      double waves = (double)fullw;
      if (rem) {
        const long long sp = units / rem;
        waves += (tail_ok && sp > 1) ? (1.0 / (double)(sp < nkb / 2 ? sp : (nkb / 2 > 0 ? nkb / 2 : 1)) + 0.15) : 1.0;
      }
      const double t = waves * (nkb * cand_cost[i] + 1.0);
      if (t < best) { best = t; bn = cand_bn[i]; pair = cand_pair[i]; }
    }
  } else if (bn == 0) {
    if (a->N % 192 == 0 && !pair) bn = 192;
    else if (a->N % 256 == 0 || pair) bn = 256;
    else if (a->N % 128 == 0) bn = 128;
    else bn = (a->N > 128) ? 192 : 128;
  } else if (a->cta_group == 0) {
    pair = false;
  }
  VAW_CHECK_ARG(bn == 128 || bn == 192 || bn == 256, "vaw_gemm_bf16: tile_n must be 128, 192 or 256");
  const int tile_m = pair ? 256 : 128;
  const int b_rows = pair ? bn / 2 : bn;

  CUtensorMap tmA, tmB;
  int rc;
  if (!a->a_mn) rc = make_tmap(&tmA, a->A, a->M, a->K, a->lda, 128);
  else rc = make_tmap(&tmA, a->A, a->K, a->M, a->lda, BK);
  if (rc) return rc;
  if (!a->b_mn) rc = make_tmap(&tmB, a->B, a->N, a->K, a->ldb, b_rows);
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
  p.n_out = n_out;
  p.a_mn = a->a_mn ? 1 : 0;
  p.b_mn = a->b_mn ? 1 : 0;

  // ---- work decomposition -------------------------------------------------------------------------------
  // k_splits > 1: every tile is split (skinny GEMMs).  k_splits == -1: "tail split" — whole tiles for the full
  // waves, the partial last wave split along K so that all SMs stay busy (wgrad GEMMs, long K).
  const int num_kb = (a->K + BK - 1) / BK;
  const int m_tiles = (a->M + tile_m - 1) / tile_m, n_tiles = (a->N + bn - 1) / bn;
  const int tiles = m_tiles * n_tiles;
  const int G = pair ? gemm_sms() / 2 : gemm_sms();
  int full = tiles, splits = 1, kb_per = num_kb;
  if (a->k_splits != 0 && a->k_splits != 1) {
    VAW_CHECK_ARG(epi == EPI_F32 && a->split_ws, "vaw_gemm_bf16: split-K needs the F32 epilogue and split_ws");
    VAW_CHECK_ARG(a->N % 4 == 0, "vaw_gemm_bf16: split-K needs N %% 4 == 0");
    int want;
    if (a->k_splits > 1) { full = 0; want = a->k_splits; }
    else { full = (tiles / G) * G; want = (tiles - full) > 0 ? G / (tiles - full) : 1; }
    const int rem = tiles - full;
    if (want > num_kb / 2) want = num_kb / 2;  // at least two k-blocks per split
    const long long slab = (long long)tile_m * bn;
    if (a->split_ws_elems > 0 && (long long)rem * want * slab > a->split_ws_elems)
      want = (int)(a->split_ws_elems / (slab * (rem > 0 ? rem : 1)));
    if (rem > 0 && want > 1) {
      kb_per = (num_kb + want - 1) / want;
      splits = (num_kb + kb_per - 1) / kb_per;
    }
    if (splits <= 1) { full = tiles; splits = 1; kb_per = num_kb; }
  }
  p.full_tiles = full;
  p.tail_splits = splits;
  p.kb_per_split = kb_per;
  p.num_work = full + (tiles - full) * splits;
  p.split_ws = a->split_ws;
  {
    const char* e = getenv("VAW_DBG");
    p.dbg = e ? atoi(e) : 0;
  }

  OutMaps om;
  om.c = tmA;   // placeholders for the epilogues that do not store through TMA
  om.c2 = tmA;
  om.c3 = tmA;
  p.tma_res = 0;
  if ((epi == EPI_GATE_RES || epi == EPI_RES) && a->resid_mod == 0 && a->resid && a->out2 &&
      ((reinterpret_cast<uintptr_t>(a->resid) | reinterpret_cast<uintptr_t>(a->out2) |
        reinterpret_cast<uintptr_t>(a->out)) & 15) == 0 &&
      (epi == EPI_RES || a->gate)) {
    rc = make_out_tmap(&om.c2, a->out2, a->M, a->N, ldo, true);
    if (!rc) rc = make_out_tmap(&om.c3, a->resid, a->M, a->N, ldo, true);
    if (!rc && a->out) rc = make_out_tmap(&om.c, a->out, a->M, a->N, ldo);
    if (rc) return rc;
    p.tma_res = 1;
  }
  const bool aux_epi = (epi == EPI_DGELU_TANH || epi == EPI_DGELU_ERF || epi == EPI_DSILU || epi == EPI_ALIGN_MSE);
  VAW_CHECK_ARG(epi != EPI_ALIGN_MSE || (a->out2 && a->N % 4 == 0), "vaw_gemm_bf16: EPI_ALIGN_MSE needs out2 (partials)");
  if (aux_epi) {
    VAW_CHECK_ARG(a->aux && (reinterpret_cast<uintptr_t>(a->aux) & 15) == 0 && a->out &&
                      (reinterpret_cast<uintptr_t>(a->out) & 15) == 0,
                  "vaw_gemm_bf16: d-activation epilogues need 16-byte aligned out and aux");
    rc = make_out_tmap(&om.c, a->out, a->M, a->N, ldo);
    if (!rc) rc = make_out_tmap(&om.c2, a->aux, a->M, a->N, ldo);
    if (rc) return rc;
  }
  if (epi == EPI_BF16 || epi == EPI_GELU_TANH || epi == EPI_GELU_ERF || epi == EPI_SILU) {
    VAW_CHECK_ARG((reinterpret_cast<uintptr_t>(a->out) & 15) == 0, "vaw_gemm_bf16: out must be 16-byte aligned");
    if (a->out) rc = make_out_tmap(&om.c, a->out, a->M, a->N, ldo);
    if (rc) return rc;
    if (epi != EPI_BF16) {
      VAW_CHECK_ARG(a->out2 && (reinterpret_cast<uintptr_t>(a->out2) & 15) == 0,
                    "vaw_gemm_bf16: this epilogue needs a 16-byte aligned out2");
      rc = make_out_tmap(&om.c2, a->out2, a->M, a->N, ldo);
      if (rc) return rc;
    }
  }
  rc = dispatch_tile(bn, pair, epi, tmA, tmB, om, p, stream);
  if (rc || splits == 1) return rc;
  {
    const int rem = tiles - full;
    float* o = reinterpret_cast<float*>(a->out);
#define VAW_FIXUP(TM_, BN_)                                                                                          \
  splitk_fixup_kernel<TM_, BN_><<<dim3(rem, kFixupSlices), 256, 0, stream>>>(a->split_ws, full, splits, n_tiles, o, a->bias, a->M, a->N, \
                                                         ldo, a->accumulate, p.n_out, reinterpret_cast<float*>(a->out2))
    if (pair) {
      if (bn == 128) VAW_FIXUP(256, 128); else if (bn == 192) VAW_FIXUP(256, 192); else VAW_FIXUP(256, 256);
    } else {
      if (bn == 128) VAW_FIXUP(128, 128); else if (bn == 192) VAW_FIXUP(128, 192); else VAW_FIXUP(128, 256);
    }
#undef VAW_FIXUP
    VAW_LAUNCH_CHECK();
  }
  return VAW_OK;
}

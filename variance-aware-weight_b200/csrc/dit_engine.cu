// dit_engine.cu — DiT forward / backward as one C call each: a fixed sequence of the library's kernels on one
// stream, working on a flat fp32 parameter buffer (+ its bf16 shadow), a flat gradient buffer and one caller-owned
// workspace.  Mirrors models/dit.py:258-280 (forward), with autograd replaced by the hand-derived backward.
//
// Data layout in HBM (per GPU):
//   P   fp32 [n_params]      master weights, tensors at 64-element aligned offsets (vaw_dit_param_layout)
//   Pb  bf16 [n_params]      shadow copy read by the GEMMs through TMA
//   G   fp32 [n_params]      gradients (wgrad epilogues write / accumulate here directly)
//   ws  workspace            saved activations (bf16 GEMM operands, fp32 residual stream x_0..x_2L) + backward temps
// The residual stream stays fp32 (as it does under the reference's autocast: x + pos_embed promotes to fp32);
// every GEMM operand is bf16, every reduction fp32.
#include "vaw_common.cuh"
#include "vaw_internal.h"
#include "engine_common.cuh"

namespace {

constexpr int kMaxDepth = 64;
constexpr float kLnEps = 1e-6f;  // dit.py:122,124,147
constexpr int kOnes = 32;        // width of the ones block behind the rows of xn1 / xn2 (training workspace)

// ---- parameter layout -------------------------------------------------------------------------------------
enum ParamId : int {
  P_XEMB_W = 0, P_XEMB_B, P_T0_W, P_T0_B, P_T2_W, P_T2_B, P_YTAB, P_POS, P_FADA_W, P_FADA_B, P_FLIN_W, P_FLIN_B,
  P_PR0_W, P_PR0_B, P_PR2_W, P_PR2_B, P_PR4_W, P_PR4_B, P_ADA_W, P_ADA_B, P_BLOCK0
};
// The four weight matrices of a block are contiguous and its four bias vectors follow: the data-parallel wrapper
// reduce-scatters the former (the optimizer state is sharded over ranks) and all-reduces the latter (a few KB, kept
// replicated because the forward reads biases from the fp32 master buffer) - one collective each per block.
enum BlockParam : int { B_QKV_W = 0, B_PROJ_W, B_FC1_W, B_FC2_W, B_QKV_B, B_PROJ_B, B_FC1_B, B_FC2_B, B_COUNT };

struct Layout {
  long long off[P_BLOCK0 + kMaxDepth * B_COUNT];
  long long numel[P_BLOCK0 + kMaxDepth * B_COUNT];
  int n;
  long long total;
};

long long align64(long long x) { return (x + 63) / 64 * 64; }

void compute_layout(const vaw_dit_cfg& c, Layout& L) {
  const long long D = c.D, Kp = (long long)c.C_in * c.P * c.P, PPC = (long long)c.P * c.P * c.C_out;
  long long cur = 0;
  int n = 0;
  auto add = [&](long long numel) {
    L.off[n] = cur;
    L.numel[n] = numel;
    cur = align64(cur + numel);
    ++n;
  };
  add(D * Kp); add(D);                                   // x_embedder.proj
  add(D * c.freq_dim); add(D); add(D * D); add(D);       // t_embedder.mlp.{0,2}
  add((long long)c.table_rows * D);                      // y_embedder.embedding_table
  add((long long)c.T * D);                               // pos_embed (no grad)
  add(2 * D * D); add(2 * D);                            // final_layer.adaLN_modulation.1
  add(PPC * D); add(PPC);                                // final_layer.linear
  const long long pd = c.learn_align ? c.proj_dim : 0, zd = c.learn_align ? c.z_dim : 0;
  add(pd * D); add(pd); add(pd * pd); add(pd); add(zd * pd); add(zd);  // projectors.{0,2,4}
  add((long long)c.depth * 6 * D * D); add((long long)c.depth * 6 * D);  // all blocks' adaLN_modulation.1, stacked
  for (int i = 0; i < c.depth; ++i) {   // order of BlockParam
    add(3 * D * D);                                      // attn.qkv.weight
    add(D * D);                                          // attn.proj.weight
    add((long long)c.hidden * D);                        // mlp.fc1.weight
    add(D * (long long)c.hidden);                        // mlp.fc2.weight
    add(3 * D); add(D); add(c.hidden); add(D);           // the four biases
  }
  L.n = n;
  L.total = cur;
}

// ---- workspace ----------------------------------------------------------------------------------------------
struct BlockWs {
  float *mean1, *rstd1, *mean2, *rstd2, *lse;
  bf16 *xn1, *qkv, *attn_o, *y_attn, *xn2, *h_pre, *h_act, *y_mlp;
};
struct Ws {
  bf16 *patches, *freq, *t_h_pre, *t_h, *c_silu;
  float *t_emb, *c, *mod_all, *mod_final;
  float* x[2 * kMaxDepth + 1];
  BlockWs blk[kMaxDepth];
  float *meanf, *rstdf;
  bf16 *xnf, *out_tok;
  bf16 *xa, *z1_pre, *z1, *z2_pre, *z2;
  // backward temporaries
  float* dx;
  bf16 *dy, *dh, *dqkv, *d_o, *dxn, *dtok, *dz, *dxa;
  float* attn_delta;   // [B*H*T] rowsum(dO*O) scratch of the attention backward
  float *part, *part_a2, *part_b, *part_c, *part_d, *cpart, *dmod_all, *dmod_final, *dcs, *dc;
  float* align_part;   // [ceil(M/32) * ceil(z_dim/32)] partial sums of the fused alignment loss (VAW_EPI_ALIGN_MSE)
  long long align_parts;
  bf16 *dmod_all_b, *dmod_final_b, *dc_b, *dth;
  float* split_ws;
  long long split_elems;
  long long bytes;
};


int ln_chunks(const vaw_dit_cfg& c) {
  // (chunks x B) CTAs of the LN / gate backward kernels, ~32 rows each: enough CTAs to fill the chip several times
  // over at high occupancy, few enough that the per-CTA partial sums stay small
  int ch = (c.T + 31) / 32;
  while ((c.T + ch - 1) / ch > 64) ++ch;
  // wave quantisation: the staged LayerNorm backward runs 2 CTAs per SM; prefer a chunk count whose grid (B x chunks)
  // nearly fills whole waves (B = 64, T = 256: 9 chunks -> 576 CTAs = 1.95 waves instead of 8 -> 1.73)
  const int slots = 2 * vaw_num_sms();
  int best = ch;
  double best_eff = 0.0;
  for (int k = ch; k <= 2 * ch + 2 && k <= c.T; ++k) {
    const long long ctas = (long long)c.B * k;
    const double eff = (double)ctas / (double)(((ctas + slots - 1) / slots) * slots);
    if (eff > best_eff + 0.02) { best_eff = eff; best = k; }
  }
  return best;
}

void carve(const vaw_dit_cfg& c, void* base, Ws& w) {
  Carver k{reinterpret_cast<uint8_t*>(base)};
  const long long B = c.B, D = c.D, M = (long long)c.B * c.T, Hd = c.hidden;
  const long long Kp = (long long)c.C_in * c.P * c.P, PPC = (long long)c.P * c.P * c.C_out;
  w.patches = k.take<bf16>(M * Kp);
  w.freq = k.take<bf16>(B * c.freq_dim);
  w.t_h_pre = k.take<bf16>(B * D);
  w.t_h = k.take<bf16>(B * D);
  w.c_silu = k.take<bf16>(B * D);
  w.t_emb = k.take<float>(B * D);
  w.c = k.take<float>(B * D);
  w.mod_all = k.take<float>(B * c.depth * 6 * D);
  w.mod_final = k.take<float>(B * 2 * D);
  for (int i = 0; i <= 2 * c.depth; ++i) w.x[i] = k.take<float>(M * D);
  for (int i = 0; i < c.depth; ++i) {
    BlockWs& b = w.blk[i];
    b.mean1 = k.take<float>(M); b.rstd1 = k.take<float>(M); b.mean2 = k.take<float>(M); b.rstd2 = k.take<float>(M);
    b.lse = k.take<float>(B * c.H * c.T);
    // xn1 / xn2 rows carry a [1, 0 x 31] block after their D values (kOnes): the qkv / fc1 weight-gradient GEMMs read
    // them as [M, D + 32] and return the bias gradients as output column D
    b.xn1 = k.take<bf16>(M * (D + kOnes)); b.qkv = k.take<bf16>(M * 3 * D); b.attn_o = k.take<bf16>(M * D);
    b.y_attn = k.take<bf16>(M * D); b.xn2 = k.take<bf16>(M * (D + kOnes)); b.h_pre = k.take<bf16>(M * Hd);
    b.h_act = k.take<bf16>(M * Hd); b.y_mlp = k.take<bf16>(M * D);
  }
  w.meanf = k.take<float>(M); w.rstdf = k.take<float>(M);
  w.xnf = k.take<bf16>(M * D);
  w.out_tok = k.take<bf16>(M * PPC);
  const long long pd = c.learn_align ? c.proj_dim : 0;
  w.xa = k.take<bf16>(c.learn_align ? M * D : 0);
  w.z1_pre = k.take<bf16>(M * pd); w.z1 = k.take<bf16>(M * pd);
  w.z2_pre = k.take<bf16>(M * pd); w.z2 = k.take<bf16>(M * pd);
  // backward temporaries
  w.dx = k.take<float>(M * D);
  w.dy = k.take<bf16>(M * D);
  w.dh = k.take<bf16>(M * Hd);
  w.dqkv = k.take<bf16>(M * 3 * D);
  w.d_o = k.take<bf16>(M * D);
  w.attn_delta = k.take<float>((long long)c.B * c.H * c.T);
  w.dxn = k.take<bf16>(M * D);
  w.dtok = k.take<bf16>(M * PPC);
  w.dz = k.take<bf16>(2 * M * pd);
  w.dxa = k.take<bf16>(c.learn_align ? M * D : 0);
  w.part = k.take<float>(B * ln_chunks(c) * 2 * D);
  w.part_a2 = k.take<float>(B * ln_chunks(c) * 2 * D);
  w.part_b = k.take<float>(B * ln_chunks(c) * 2 * D);
  w.part_c = k.take<float>(B * ln_chunks(c) * 2 * D);
  w.part_d = k.take<float>(B * ln_chunks(c) * 2 * D);
  long long maxN = 3 * D > Hd ? 3 * D : Hd;
  if (pd > maxN) maxN = pd;
  const long long cchunks = (M + colsum_rows((int)M, 64) - 1) / colsum_rows((int)M, 64) + 1;
  w.cpart = k.take<float>(cchunks * maxN + 1024);
  w.dmod_all = k.take<float>(B * c.depth * 6 * D);
  w.dmod_final = k.take<float>(B * 2 * D);
  w.dcs = k.take<float>(B * D);
  w.dc = k.take<float>(B * D);
  w.dmod_all_b = k.take<bf16>(B * c.depth * 6 * D);
  w.dmod_final_b = k.take<bf16>(B * 2 * D);
  w.dc_b = k.take<bf16>(B * D);
  w.dth = k.take<bf16>(B * D);
  w.split_elems = (long long)vaw_num_sms() * 128 * 256;  // one 128x256 fp32 slab per SM
  w.split_ws = k.take<float>(w.split_elems);
  w.align_parts = c.learn_align ? ((M + 31) / 32) * ((c.z_dim + 31) / 32) : 0;
  w.align_part = k.take<float>(w.align_parts);
  w.bytes = k.cur;
}

// Forward-only layout (sampling / evaluation): nothing is kept for a backward pass, so every block re-uses ONE set of
// operand buffers, the residual stream cycles through three buffers, and the tensors only the backward reads (saved
// pre-activations, branch outputs) are not written at all.  DiT-XL/2 at B = 128: 1.3 GB instead of 40 GB, and a third of
// the forward's HBM writes gone.
void carve_infer(const vaw_dit_cfg& c, void* base, Ws& w) {
  Carver k{reinterpret_cast<uint8_t*>(base)};
  const long long B = c.B, D = c.D, M = (long long)c.B * c.T, Hd = c.hidden;
  const long long Kp = (long long)c.C_in * c.P * c.P, PPC = (long long)c.P * c.P * c.C_out;
  memset(&w, 0, sizeof(w));
  w.patches = k.take<bf16>(M * Kp);
  w.freq = k.take<bf16>(B * c.freq_dim);
  w.t_h = k.take<bf16>(B * D);
  w.c_silu = k.take<bf16>(B * D);
  w.t_emb = k.take<float>(B * D);
  w.c = k.take<float>(B * D);
  w.mod_all = k.take<float>(B * c.depth * 6 * D);
  w.mod_final = k.take<float>(B * 2 * D);
  float* ring[3] = {k.take<float>(M * D), k.take<float>(M * D), k.take<float>(M * D)};
  for (int i = 0; i <= 2 * c.depth; ++i) w.x[i] = ring[i % 3];
  BlockWs b{};
  b.mean1 = k.take<float>(M); b.rstd1 = k.take<float>(M); b.mean2 = b.mean1; b.rstd2 = b.rstd1;
  b.lse = k.take<float>(B * c.H * c.T);
  b.xn1 = k.take<bf16>(M * D); b.qkv = k.take<bf16>(M * 3 * D); b.attn_o = k.take<bf16>(M * D);
  b.xn2 = b.xn1;
  b.h_act = k.take<bf16>(M * Hd);
  b.y_attn = k.take<bf16>(M * D);   // branch output on its way to the next LayerNorm pass (vaw_ln_fwd_res); one buffer:
  b.y_mlp = b.y_attn;               // each is consumed before the other branch's GEMM writes
  for (int i = 0; i < c.depth; ++i) w.blk[i] = b;   // h_pre stays null: not stored
  w.meanf = b.mean1; w.rstdf = b.rstd1;
  w.xnf = b.xn1;
  w.out_tok = k.take<bf16>(M * PPC);
  const long long pd = c.learn_align ? c.proj_dim : 0;
  w.xa = c.learn_align ? b.xn1 : nullptr;
  w.z1 = k.take<bf16>(M * pd);
  w.z2 = k.take<bf16>(M * pd);
  w.align_parts = c.learn_align ? ((M + 31) / 32) * ((c.z_dim + 31) / 32) : 0;
  w.align_part = k.take<float>(w.align_parts);
  w.bytes = k.cur;
}

// dW[M, N] (+)= A^T B over K = batch rows: the short-K kernel when the layout allows, the general GEMM otherwise
int wgrad_over_batch(const bf16* A, long long lda, const bf16* Bm, long long ldb, float* out, int M, int N, int K, int acc,
                     cudaStream_t s) {
  const int rc = vaw_wgrad_smallk(A, lda, Bm, ldb, out, N, M, N, K, acc, s);
  if (rc != VAW_ERR_UNSUPPORTED) return rc;
  return G(A, lda, 1, Bm, ldb, 1, M, N, K, VAW_EPI_F32).out(out).acc(acc).run(s);
}

int check_cfg(const vaw_dit_cfg* c) {
  VAW_CHECK_ARG(c, "dit: null config");
  VAW_CHECK_ARG(c->B > 0 && c->T > 0 && c->D > 0 && c->H > 0 && c->depth > 0 && c->depth <= kMaxDepth,
                "dit: bad geometry B=%d T=%d D=%d H=%d depth=%d", c->B, c->T, c->D, c->H, c->depth);
  VAW_CHECK_ARG(c->D % c->H == 0 && (c->D / c->H == 64 || c->D / c->H == 72), "dit: head_dim %d unsupported (64, 72)",
                c->D / c->H);
  VAW_CHECK_ARG(c->D % 8 == 0 && c->hidden % 8 == 0, "dit: D and hidden must be multiples of 8");
  VAW_CHECK_ARG((c->img_h / c->P) * (c->img_w / c->P) == c->T, "dit: T does not match the patch grid");
  VAW_CHECK_ARG((c->C_in * c->P * c->P) % 8 == 0 && (c->C_out * c->P * c->P) % 8 == 0,
                "dit: patch feature counts must be multiples of 8");
  VAW_CHECK_ARG(c->freq_dim % 8 == 0, "dit: freq_dim must be a multiple of 8");
  VAW_CHECK_ARG(!c->learn_align || (c->encoder_depth > 0 && c->encoder_depth <= c->depth && c->proj_dim % 8 == 0 &&
                                    c->z_dim % 8 == 0),
                "dit: bad REPA projector config");
  return VAW_OK;
}

}  // namespace

extern "C" int vaw_dit_param_layout(const vaw_dit_cfg* cfg, long long* offsets, long long* numels, int cap, int* n_out,
                                    long long* total) {
  TRY(check_cfg(cfg));
  Layout L;
  compute_layout(*cfg, L);
  VAW_CHECK_ARG(cap >= L.n, "vaw_dit_param_layout: need room for %d entries", L.n);
  for (int i = 0; i < L.n; ++i) {
    if (offsets) offsets[i] = L.off[i];
    if (numels) numels[i] = L.numel[i];
  }
  if (n_out) *n_out = L.n;
  if (total) *total = L.total;
  return VAW_OK;
}

extern "C" int vaw_dit_workspace_bytes(const vaw_dit_cfg* cfg, long long* bytes) {
  TRY(check_cfg(cfg));
  VAW_CHECK_ARG(bytes, "vaw_dit_workspace_bytes: null output");
  Ws w;
  carve(*cfg, nullptr, w);
  *bytes = w.bytes;
  return VAW_OK;
}

// x_t [B,C_in,H,W] fp32, t [B] fp32 (already scaled, gaussian_diffusion.py:417-420), y [B] int64 (may be null when
// table_rows == 0) -> out [B,C_out,H,W] bf16 and, with learn_align, zs [B*T, z_dim] bf16.
static int dit_forward_impl(const vaw_dit_cfg* cfg, const float* P, const void* Pb_, void* ws_, const float* x_t,
                            const float* t, const long long* y, void* out, void* zs, const void* feat, float* align_loss,
                            cudaStream_t s, bool infer = false, void** wait_events = nullptr) {
  TRY(check_cfg(cfg));
  VAW_CHECK_ARG(P && Pb_ && ws_ && x_t && t && out, "vaw_dit_forward: null pointer");
  const vaw_dit_cfg& c = *cfg;
  VAW_CHECK_ARG(c.table_rows == 0 || y, "vaw_dit_forward: labels required");
  VAW_CHECK_ARG(!c.learn_align || zs, "vaw_dit_forward: zs output required with learn_align");
  Layout L;
  compute_layout(c, L);
  Ws w;
  if (infer) carve_infer(c, ws_, w);
  else carve(c, ws_, w);
  const bf16* Pb = reinterpret_cast<const bf16*>(Pb_);
  const int B = c.B, T = c.T, D = c.D, M = B * T, Hd = c.hidden, hd = D / c.H;
  const int Kp = c.C_in * c.P * c.P, PPC = c.P * c.P * c.C_out;
  const long long ldm = (long long)c.depth * 6 * D;

  // wait_events (sharded data-parallel optimizer): [0] = the stacked adaLN weights' bf16 shadow is complete, [1 + i] =
  // block i's.  The all-gathers of the later blocks run under the earlier blocks' compute.
  auto wait_for = [&](int k) -> int {
    if (wait_events && wait_events[k]) VAW_CUDA_TRY(cudaStreamWaitEvent(s, reinterpret_cast<cudaEvent_t>(wait_events[k]), 0));
    return VAW_OK;
  };
  // patch embedding + pos_embed -> x[0]
  TRY(vaw_patchify_in(x_t, w.patches, B, c.C_in, c.img_h, c.img_w, c.P, s));
  TRY(G(w.patches, Kp, 0, Pb + L.off[P_XEMB_W], Kp, 0, M, D, Kp, VAW_EPI_RES)
          .out(nullptr, w.x[0]).bias(P + L.off[P_XEMB_B]).resid(P + L.off[P_POS], T).run(s));
  // conditioning: c = t_mlp(freq(t)) + table[y]
  TRY(vaw_timestep_embedding(t, w.freq, nullptr, B, c.freq_dim, s));
  TRY(G(w.freq, c.freq_dim, 0, Pb + L.off[P_T0_W], c.freq_dim, 0, B, D, c.freq_dim, VAW_EPI_SILU)
          .out(w.t_h_pre, w.t_h).bias(P + L.off[P_T0_B]).run(s));
  TRY(G(w.t_h, D, 0, Pb + L.off[P_T2_W], D, 0, B, D, D, VAW_EPI_F32).out(w.t_emb).bias(P + L.off[P_T2_B]).run(s));
  TRY(vaw_cond_combine(w.t_emb, c.table_rows ? P + L.off[P_YTAB] : nullptr, y, w.c, w.c_silu, B, D, s));
  // adaLN modulation of every block (one GEMM) and of the final layer
  TRY(wait_for(0));
  TRY(G(w.c_silu, D, 0, Pb + L.off[P_ADA_W], D, 0, B, c.depth * 6 * D, D, VAW_EPI_F32)
          .out(w.mod_all).bias(P + L.off[P_ADA_B]).run(s));
  TRY(G(w.c_silu, D, 0, Pb + L.off[P_FADA_W], D, 0, B, 2 * D, D, VAW_EPI_F32)
          .out(w.mod_final).bias(P + L.off[P_FADA_B]).run(s));

  // VAW_DIT_GATE_EPI=1 (A/B measurement): the residual updates run in the proj / fc2 GEMM epilogues (VAW_EPI_GATE_RES)
  static const bool gate_epi = getenv("VAW_DIT_GATE_EPI") && atoi(getenv("VAW_DIT_GATE_EPI")) != 0;
  bool pending = false;
  const long long ldx = infer ? D : D + kOnes;   // row stride of xn1 / xn2
  const int ones = infer ? 0 : 1;
  for (int i = 0; i < c.depth; ++i) {
    BlockWs& b = w.blk[i];
    const int pb = P_BLOCK0 + i * B_COUNT;
    const float* mod = w.mod_all + (long long)i * 6 * D;  // [shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp]
    float* x_in = w.x[2 * i];
    float* x_mid = w.x[2 * i + 1];
    float* x_out = w.x[2 * i + 2];
    TRY(wait_for(1 + i));
    // The residual update  x += gate * branch  of each branch is formed by the LayerNorm pass that FOLLOWS it
    // (vaw_ln_fwd_res): proj / fc2 are plain bf16-output GEMMs, and the fp32 residual stream is read once instead of by
    // a GEMM epilogue and again by the LayerNorm.  `pending` = the previous block's MLP branch is not yet folded into
    // x_in (its x_out buffer is written here).
    if (pending) {
      const float* pmod = w.mod_all + (long long)(i - 1) * 6 * D;
      TRY(vaw_ln_fwd_ex(w.x[2 * i - 1], w.blk[i - 1].y_mlp, pmod + 5 * D, ldm, x_in, mod, mod + D, ldm, T, nullptr, nullptr,
                        b.xn1, ldx, ones, b.mean1, b.rstd1, M, D, kLnEps, s));
    } else {
      TRY(vaw_ln_fwd_ex(x_in, nullptr, nullptr, 0, nullptr, mod, mod + D, ldm, T, nullptr, nullptr, b.xn1, ldx, ones,
                        b.mean1, b.rstd1, M, D, kLnEps, s));
    }
    TRY(G(b.xn1, ldx, 0, Pb + L.off[pb + B_QKV_W], D, 0, M, 3 * D, D, VAW_EPI_BF16)
            .out(b.qkv).bias(P + L.off[pb + B_QKV_B]).run(s));
    TRY(vaw_attn_fwd(b.qkv, b.attn_o, b.lse, B, T, c.H, hd, s));
    if (gate_epi) {
      TRY(G(b.attn_o, D, 0, Pb + L.off[pb + B_PROJ_W], D, 0, M, D, D, VAW_EPI_GATE_RES)
              .out(b.y_attn, x_mid).bias(P + L.off[pb + B_PROJ_B]).resid(x_in).gate(mod + 2 * D, ldm, T).run(s));
      TRY(vaw_ln_fwd_ex(x_mid, nullptr, nullptr, 0, nullptr, mod + 3 * D, mod + 4 * D, ldm, T, nullptr, nullptr, b.xn2, ldx,
                        ones, b.mean2, b.rstd2, M, D, kLnEps, s));
    } else {
      TRY(G(b.attn_o, D, 0, Pb + L.off[pb + B_PROJ_W], D, 0, M, D, D, VAW_EPI_BF16)
              .out(b.y_attn).bias(P + L.off[pb + B_PROJ_B]).run(s));
      TRY(vaw_ln_fwd_ex(x_in, b.y_attn, mod + 2 * D, ldm, x_mid, mod + 3 * D, mod + 4 * D, ldm, T, nullptr, nullptr, b.xn2,
                        ldx, ones, b.mean2, b.rstd2, M, D, kLnEps, s));
    }
    TRY(G(b.xn2, ldx, 0, Pb + L.off[pb + B_FC1_W], D, 0, M, Hd, D, VAW_EPI_GELU_TANH)
            .out(b.h_pre, b.h_act).bias(P + L.off[pb + B_FC1_B]).run(s));
    // the block whose output feeds the REPA projectors needs x_out right away: it keeps the gated-residual epilogue
    const bool tap = gate_epi || (c.learn_align && i + 1 == c.encoder_depth);
    if (tap) {
      TRY(G(b.h_act, Hd, 0, Pb + L.off[pb + B_FC2_W], Hd, 0, M, D, Hd, VAW_EPI_GATE_RES)
              .out(b.y_mlp, x_out).bias(P + L.off[pb + B_FC2_B]).resid(x_mid).gate(mod + 5 * D, ldm, T).run(s));
    } else {
      TRY(G(b.h_act, Hd, 0, Pb + L.off[pb + B_FC2_W], Hd, 0, M, D, Hd, VAW_EPI_BF16)
              .out(b.y_mlp).bias(P + L.off[pb + B_FC2_B]).run(s));
    }
    pending = !tap;
    if (c.learn_align && i + 1 == c.encoder_depth) {
      const int pd = c.proj_dim, zd = c.z_dim;
      TRY(vaw_cast_f32_bf16(x_out, w.xa, (long long)M * D, s));
      TRY(G(w.xa, D, 0, Pb + L.off[P_PR0_W], D, 0, M, pd, D, VAW_EPI_SILU)
              .out(w.z1_pre, w.z1).bias(P + L.off[P_PR0_B]).run(s));
      TRY(G(w.z1, pd, 0, Pb + L.off[P_PR2_W], pd, 0, M, pd, pd, VAW_EPI_SILU)
              .out(w.z2_pre, w.z2).bias(P + L.off[P_PR2_B]).run(s));
      if (feat) {
        // REPA: Sigma (zs - feat)^2 is accumulated in the epilogue that produces zs (north-star piece 5); the partials
        // are folded in fixed order by one small launch
        TRY(G(w.z2, pd, 0, Pb + L.off[P_PR4_W], pd, 0, M, zd, pd, VAW_EPI_ALIGN_MSE)
                .out(zs, w.align_part).bias(P + L.off[P_PR4_B]).aux(feat).run(s));
        TRY(vaw_align_mse_finish(w.align_part, w.align_parts, (long long)M * zd, align_loss, s));
      } else {
        TRY(G(w.z2, pd, 0, Pb + L.off[P_PR4_W], pd, 0, M, zd, pd, VAW_EPI_BF16).out(zs).bias(P + L.off[P_PR4_B]).run(s));
      }
    }
  }
  // final layer: LN -> modulate -> linear -> unpatchify
  float* x_last = w.x[2 * c.depth];
  if (pending) {
    const float* pmod = w.mod_all + (long long)(c.depth - 1) * 6 * D;
    TRY(vaw_ln_fwd_res(w.x[2 * c.depth - 1], w.blk[c.depth - 1].y_mlp, pmod + 5 * D, ldm, x_last, w.mod_final,
                       w.mod_final + D, 2LL * D, T, w.xnf, w.meanf, w.rstdf, M, D, kLnEps, s));
  } else {
    TRY(vaw_ln_fwd(x_last, w.mod_final, w.mod_final + D, 2LL * D, T, nullptr, nullptr, w.xnf, w.meanf, w.rstdf, M, D,
                   kLnEps, s));
  }
  TRY(G(w.xnf, D, 0, Pb + L.off[P_FLIN_W], D, 0, M, PPC, D, VAW_EPI_BF16).out(w.out_tok).bias(P + L.off[P_FLIN_B]).run(s));
  TRY(vaw_unpatchify(w.out_tok, out, 1, B, c.C_out, c.img_h, c.img_w, c.P, 1, s));
  return VAW_OK;
}

extern "C" int vaw_dit_forward(const vaw_dit_cfg* cfg, const float* P, const void* Pb_, void* ws_, const float* x_t,
                               const float* t, const long long* y, void* out, void* zs, cudaStream_t s) {
  return dit_forward_impl(cfg, P, Pb_, ws_, x_t, t, y, out, zs, nullptr, nullptr, s);
}

// vaw_dit_forward with per-block wait events: depth + 1 cudaEvent_t (or NULL entries) that gate the first read of the
// stacked adaLN weights ([0]) and of each block's weights ([1 + i]) - see parallel.ShardedGradSync.
extern "C" int vaw_dit_forward_ev(const vaw_dit_cfg* cfg, const float* P, const void* Pb_, void* ws_, const float* x_t,
                                  const float* t, const long long* y, void* out, void* zs, void** wait_events,
                                  cudaStream_t s) {
  return dit_forward_impl(cfg, P, Pb_, ws_, x_t, t, y, out, zs, nullptr, nullptr, s, false, wait_events);
}

// Forward-only entry (sampling / evaluation, torch.no_grad()): same arithmetic and results as vaw_dit_forward, but on
// the compact workspace of carve_infer - no activation stash.  The training workspace is not touched.
extern "C" int vaw_dit_infer_workspace_bytes(const vaw_dit_cfg* cfg, long long* bytes) {
  TRY(check_cfg(cfg));
  VAW_CHECK_ARG(bytes, "vaw_dit_infer_workspace_bytes: null output");
  Ws w;
  carve_infer(*cfg, nullptr, w);
  *bytes = w.bytes;
  return VAW_OK;
}

extern "C" int vaw_dit_forward_infer(const vaw_dit_cfg* cfg, const float* P, const void* Pb_, void* ws_,
                                     const float* x_t, const float* t, const long long* y, void* out, void* zs,
                                     cudaStream_t s) {
  return dit_forward_impl(cfg, P, Pb_, ws_, x_t, t, y, out, zs, nullptr, nullptr, s, true);
}

extern "C" int vaw_dit_forward_align(const vaw_dit_cfg* cfg, const float* P, const void* Pb_, void* ws_,
                                     const float* x_t, const float* t, const long long* y, void* out, void* zs,
                                     const void* feat, float* align_loss, cudaStream_t s) {
  VAW_CHECK_ARG(cfg && cfg->learn_align && feat && align_loss,
                "vaw_dit_forward_align: needs a learn_align config, teacher features and a loss output");
  VAW_CHECK_ARG((reinterpret_cast<uintptr_t>(feat) & 15) == 0 && cfg->z_dim % 8 == 0,
                "vaw_dit_forward_align: features must be 16-byte aligned bf16 [B*T, z_dim]");
  return dit_forward_impl(cfg, P, Pb_, ws_, x_t, t, y, out, zs, feat, align_loss, s);
}

// dout [B,C_out,H,W] bf16 (gradient of the model output), dzs [B*T, z_dim] bf16 or null.
// accumulate = 0: G is overwritten for every trainable tensor; 1: gradients are added (grad accumulation).
// events: optional array of depth + 1 cudaEvent_t; events[i] is recorded once block i's gradients are complete
// (i = depth-1 .. 0), events[depth] after the remaining (embedder / adaLN / final-layer) gradients.
extern "C" int vaw_dit_backward(const vaw_dit_cfg* cfg, const float* P, const void* Pb_, float* Gd, void* ws_,
                                const void* dout, const void* dzs, const long long* y, int accumulate, void** events,
                                cudaStream_t s) {
  TRY(check_cfg(cfg));
  VAW_CHECK_ARG(P && Pb_ && Gd && ws_ && dout, "vaw_dit_backward: null pointer");
  const vaw_dit_cfg& c = *cfg;
  Layout L;
  compute_layout(c, L);
  Ws w;
  carve(c, ws_, w);
  const bf16* Pb = reinterpret_cast<const bf16*>(Pb_);
  const int B = c.B, T = c.T, D = c.D, M = B * T, Hd = c.hidden, hd = D / c.H;
  const int Kp = c.C_in * c.P * c.P, PPC = c.P * c.P * c.C_out;
  const long long ldm = (long long)c.depth * 6 * D;
  const int ch = ln_chunks(c);
  const int acc = accumulate ? 1 : 0;

  // ---- final layer ----
  TRY(vaw_unpatchify(w.dtok, const_cast<void*>(dout), 1, B, c.C_out, c.img_h, c.img_w, c.P, 0, s));
  TRY(vaw_colsum_bf16(w.dtok, PPC, M, PPC, w.cpart, colsum_rows(M, PPC), Gd + L.off[P_FLIN_B], acc, s));
  TRY(G(w.dtok, PPC, 1, w.xnf, D, 1, PPC, D, M, VAW_EPI_F32).out(Gd + L.off[P_FLIN_W]).acc(acc)
          .autosplit(w.split_ws, w.split_elems).run(s));
  TRY(G(w.dtok, PPC, 0, Pb + L.off[P_FLIN_W], D, 1, M, D, PPC, VAW_EPI_BF16).out(w.dxn).run(s));
  // Every LayerNorm backward also runs the gate * branch backward of the branch that FOLLOWS it in the backward pass
  // (vaw_ln_bwd_gate): the updated residual gradient is used while it is still in registers instead of being re-read.
  // pa(i) = partial buffer of block i's MLP-branch gate backward; two buffers alternate because block i's buffer is
  // still unread (vaw_dit_block_finish at the end of block i) when block i-1's is produced.
  auto pa = [&](int i) { return (i & 1) ? w.part_a2 : w.part; };
  // the REPA projector injects an extra gradient into dx at the top of one block: that block's MLP gate backward must
  // see it, so it cannot ride on the previous LayerNorm backward
  auto injects = [&](int i) { return c.learn_align && i + 1 == c.encoder_depth && dzs != nullptr; };
  {
    const int i = c.depth - 1;
    const float* mod = w.mod_all + (long long)i * 6 * D;
    if (!injects(i)) {
      TRY(vaw_ln_bwd_gate(w.dxn, w.x[2 * c.depth], w.meanf, w.rstdf, w.mod_final + D, 2LL * D, nullptr, w.dx, 0, w.part_d,
                          w.blk[i].y_mlp, mod + 5 * D, ldm, w.dy, pa(i), T, B, ch, M, D, s));
    } else {
      TRY(vaw_ln_bwd(w.dxn, w.x[2 * c.depth], w.meanf, w.rstdf, w.mod_final + D, 2LL * D, nullptr, w.dx, 0, w.part_d, T,
                     B, ch, M, D, s));
    }
  }
  TRY(vaw_finish_group(w.part_d, 0, B, ch, D, w.dmod_final, 2LL * D, 0, s));      // d shift
  TRY(vaw_finish_group(w.part_d, 1, B, ch, D, w.dmod_final + D, 2LL * D, 0, s));  // d scale

  // VAW_DIT_COLSUM_BIAS=1 (A/B measurement): qkv / fc1 bias gradients from separate column-sum passes
  static const bool colsum_bias = getenv("VAW_DIT_COLSUM_BIAS") && atoi(getenv("VAW_DIT_COLSUM_BIAS")) != 0;
  for (int i = c.depth - 1; i >= 0; --i) {
    BlockWs& b = w.blk[i];
    const int pb = P_BLOCK0 + i * B_COUNT;
    const float* mod = w.mod_all + (long long)i * 6 * D;
    float* dmod = w.dmod_all + (long long)i * 6 * D;
    if (c.learn_align && i + 1 == c.encoder_depth && dzs) {
      // REPA projector backward: dzs -> grads of projectors.{4,2,0} and an extra gradient on x[2i+2]
      const int pd = c.proj_dim, zd = c.z_dim;
      bf16* dz2 = w.dz;
      bf16* dz1 = w.dz + (long long)M * pd;
      TRY(vaw_colsum_bf16(dzs, zd, M, zd, w.cpart, colsum_rows(M, zd), Gd + L.off[P_PR4_B], acc, s));
      TRY(G(dzs, zd, 1, w.z2, pd, 1, zd, pd, M, VAW_EPI_F32).out(Gd + L.off[P_PR4_W]).acc(acc)
              .autosplit(w.split_ws, w.split_elems).run(s));
      TRY(G(dzs, zd, 0, Pb + L.off[P_PR4_W], pd, 1, M, pd, zd, VAW_EPI_DSILU).out(dz2).aux(w.z2_pre).run(s));
      TRY(vaw_colsum_bf16(dz2, pd, M, pd, w.cpart, colsum_rows(M, pd), Gd + L.off[P_PR2_B], acc, s));
      TRY(G(dz2, pd, 1, w.z1, pd, 1, pd, pd, M, VAW_EPI_F32).out(Gd + L.off[P_PR2_W]).acc(acc)
              .autosplit(w.split_ws, w.split_elems).run(s));
      TRY(G(dz2, pd, 0, Pb + L.off[P_PR2_W], pd, 1, M, pd, pd, VAW_EPI_DSILU).out(dz1).aux(w.z1_pre).run(s));
      TRY(vaw_colsum_bf16(dz1, pd, M, pd, w.cpart, colsum_rows(M, pd), Gd + L.off[P_PR0_B], acc, s));
      TRY(G(dz1, pd, 1, w.xa, D, 1, pd, D, M, VAW_EPI_F32).out(Gd + L.off[P_PR0_W]).acc(acc)
              .autosplit(w.split_ws, w.split_elems).run(s));
      TRY(G(dz1, pd, 0, Pb + L.off[P_PR0_W], D, 1, M, D, pd, VAW_EPI_BF16).out(w.dxa).run(s));
      TRY(vaw_add_bf16_into_f32(w.dxa, w.dx, (long long)M * D, s));
    }
    // ---- MLP branch: x_out = x_mid + gate_mlp * fc2(gelu(fc1(modulate(LN(x_mid))))) ----
    // (the partial sums of the four row kernels below are folded by one vaw_dit_block_finish at the end of the block)
    if (injects(i)) TRY(vaw_gate_bwd(w.dx, b.y_mlp, mod + 5 * D, ldm, w.dy, pa(i), T, B, ch, M, D, s));
    // else: w.dy and pa(i) were produced by the LayerNorm backward that precedes this block in the backward pass
    TRY(G(w.dy, D, 1, b.h_act, Hd, 1, D, Hd, M, VAW_EPI_F32).out(Gd + L.off[pb + B_FC2_W]).acc(acc)
            .autosplit(w.split_ws, w.split_elems).run(s));
    TRY(G(w.dy, D, 0, Pb + L.off[pb + B_FC2_W], Hd, 1, M, Hd, D, VAW_EPI_DGELU_TANH).out(w.dh).aux(b.h_pre).run(s));
    if (colsum_bias) {   // A/B knob: separate column-sum pass for the bias gradient
      TRY(vaw_colsum_bf16(w.dh, Hd, M, Hd, w.cpart, colsum_rows(M, Hd), Gd + L.off[pb + B_FC1_B], acc, s));
      TRY(G(w.dh, Hd, 1, b.xn2, D + kOnes, 1, Hd, D, M, VAW_EPI_F32).out(Gd + L.off[pb + B_FC1_W]).acc(acc)
              .autosplit(w.split_ws, w.split_elems).run(s));
    } else {
      // weight and bias gradient of fc1 in one GEMM: xn2 carries the ones block, output column D is sum_rows(dh)
      TRY(G(w.dh, Hd, 1, b.xn2, D + kOnes, 1, Hd, D + kOnes, M, VAW_EPI_F32)
              .out(Gd + L.off[pb + B_FC1_W], Gd + L.off[pb + B_FC1_B]).acc(acc).autosplit(w.split_ws, w.split_elems).run(s));
    }
    TRY(G(w.dh, Hd, 0, Pb + L.off[pb + B_FC1_W], D, 1, M, D, Hd, VAW_EPI_BF16).out(w.dxn).run(s));
    // ---- attention branch: x_mid = x_in + gate_msa * proj(attn(qkv(modulate(LN(x_in))))) ----
    // (its gate backward is fused into the MLP LayerNorm backward)
    TRY(vaw_ln_bwd_gate(w.dxn, w.x[2 * i + 1], b.mean2, b.rstd2, mod + 4 * D, ldm, nullptr, w.dx, 1, w.part_b, b.y_attn,
                        mod + 2 * D, ldm, w.dy, w.part_c, T, B, ch, M, D, s));
    TRY(G(w.dy, D, 1, b.attn_o, D, 1, D, D, M, VAW_EPI_F32).out(Gd + L.off[pb + B_PROJ_W]).acc(acc)
            .autosplit(w.split_ws, w.split_elems).run(s));
    TRY(G(w.dy, D, 0, Pb + L.off[pb + B_PROJ_W], D, 1, M, D, D, VAW_EPI_BF16).out(w.d_o).run(s));
    TRY(vaw_attn_bwd_ws(b.qkv, b.attn_o, w.d_o, b.lse, w.dqkv, w.attn_delta, B, T, c.H, hd, s));
    if (colsum_bias) {
      TRY(vaw_colsum_bf16(w.dqkv, 3LL * D, M, 3 * D, w.cpart, colsum_rows(M, 3 * D), Gd + L.off[pb + B_QKV_B], acc, s));
      TRY(G(w.dqkv, 3LL * D, 1, b.xn1, D + kOnes, 1, 3 * D, D, M, VAW_EPI_F32).out(Gd + L.off[pb + B_QKV_W]).acc(acc)
              .autosplit(w.split_ws, w.split_elems).run(s));
    } else {
      TRY(G(w.dqkv, 3LL * D, 1, b.xn1, D + kOnes, 1, 3 * D, D + kOnes, M, VAW_EPI_F32)   // qkv weight + bias gradient
              .out(Gd + L.off[pb + B_QKV_W], Gd + L.off[pb + B_QKV_B]).acc(acc).autosplit(w.split_ws, w.split_elems).run(s));
    }
    TRY(G(w.dqkv, 3LL * D, 0, Pb + L.off[pb + B_QKV_W], D, 1, M, D, 3 * D, VAW_EPI_BF16).out(w.dxn).run(s));
    if (i > 0 && !injects(i - 1)) {   // + the MLP-branch gate backward of block i-1
      TRY(vaw_ln_bwd_gate(w.dxn, w.x[2 * i], b.mean1, b.rstd1, mod + D, ldm, nullptr, w.dx, 1, w.part_d,
                          w.blk[i - 1].y_mlp, mod - 6 * D + 5 * D, ldm, w.dy, pa(i - 1), T, B, ch, M, D, s));
    } else {
      TRY(vaw_ln_bwd(w.dxn, w.x[2 * i], b.mean1, b.rstd1, mod + D, ldm, nullptr, w.dx, 1, w.part_d, T, B, ch, M, D, s));
    }
    // adaLN_modulation.1 of this block: mod_i = silu(c) W_i^T + b_i.  Its gradient is final here, so the block's
    // whole parameter set can be all-reduced while the remaining blocks are still in backward.
    {
      bf16* dmod_b = w.dmod_all_b + (long long)i * 6 * D;
      // d mod (fp32 + bf16), d fc2.bias, d proj.bias and the adaLN bias gradient from the four partial buffers
      TRY(vaw_dit_block_finish(pa(i), w.part_b, w.part_c, w.part_d, B, ch, D, mod, ldm, dmod, dmod_b,
                               Gd + L.off[pb + B_FC2_B], Gd + L.off[pb + B_PROJ_B],
                               Gd + L.off[P_ADA_B] + (long long)i * 6 * D, acc, s));
      TRY(wgrad_over_batch(dmod_b, ldm, w.c_silu, D, Gd + L.off[P_ADA_W] + (long long)i * 6 * D * D, 6 * D, D, B, acc, s));
    }
    if (events && events[i]) VAW_CUDA_TRY(cudaEventRecord(reinterpret_cast<cudaEvent_t>(events[i]), s));
  }

  // ---- adaLN linears: mod = silu(c) W^T + b ----
  const int NA = c.depth * 6 * D;
  TRY(vaw_cast_f32_bf16(w.dmod_final, w.dmod_final_b, (long long)B * 2 * D, s));
  TRY(vaw_colsum_f32_small(w.dmod_final, 2LL * D, B, 2 * D, Gd + L.off[P_FADA_B], acc, s));
  TRY(wgrad_over_batch(w.dmod_final_b, 2LL * D, w.c_silu, D, Gd + L.off[P_FADA_W], 2 * D, D, B, acc, s));
  TRY(G(w.dmod_all_b, NA, 0, Pb + L.off[P_ADA_W], D, 1, B, D, NA, VAW_EPI_F32).out(w.dcs)
          .autosplit(w.split_ws, w.split_elems).run(s));
  TRY(G(w.dmod_final_b, 2LL * D, 0, Pb + L.off[P_FADA_W], D, 1, B, D, 2 * D, VAW_EPI_F32).out(w.dcs).acc(1).run(s));
  // ---- conditioning: c = t_emb + table[y] ----
  TRY(vaw_cond_bwd(w.dcs, w.c, w.dc, w.dc_b, B * D, s));
  if (c.table_rows) TRY(vaw_embedding_grad(w.dc, y, Gd + L.off[P_YTAB], c.table_rows, B, D, acc, s));
  TRY(vaw_colsum_f32_small(w.dc, D, B, D, Gd + L.off[P_T2_B], acc, s));
  TRY(wgrad_over_batch(w.dc_b, D, w.t_h, D, Gd + L.off[P_T2_W], D, D, B, acc, s));
  TRY(G(w.dc_b, D, 0, Pb + L.off[P_T2_W], D, 1, B, D, D, VAW_EPI_DSILU).out(w.dth).aux(w.t_h_pre).run(s));
  TRY(vaw_colsum_bf16(w.dth, D, B, D, w.cpart, 8, Gd + L.off[P_T0_B], acc, s));
  TRY(wgrad_over_batch(w.dth, D, w.freq, c.freq_dim, Gd + L.off[P_T0_W], D, c.freq_dim, B, acc, s));
  // ---- patch embedding: x[0] = patches W^T + b + pos ----
  TRY(vaw_gate_bwd(w.dx, nullptr, nullptr, 0, w.dy, w.part, T, B, ch, M, D, s));
  TRY(vaw_finish_all(w.part, 0, B, ch, D, nullptr, 0, Gd + L.off[P_XEMB_B], acc, s));
  TRY(G(w.dy, D, 1, w.patches, Kp, 1, D, Kp, M, VAW_EPI_F32).out(Gd + L.off[P_XEMB_W]).acc(acc)
          .autosplit(w.split_ws, w.split_elems).run(s));
  if (events && events[c.depth]) VAW_CUDA_TRY(cudaEventRecord(reinterpret_cast<cudaEvent_t>(events[c.depth]), s));
  return VAW_OK;
}

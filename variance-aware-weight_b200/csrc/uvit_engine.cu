// uvit_engine.cu — U-ViT forward / backward as one C call each (models/uvit.py:139-250), on the same kernels as
// the DiT engine: tcgen05 GEMMs (bias / exact-GELU / plain-residual epilogues), flash attention (T = 2 + L tokens,
// head_dim 64), affine LayerNorm kernels, plus the U-ViT glue in uvit_kernels.cu (token assembly, skip concat, conv).
//
// Block (uvit.py:96-121):  [x = skip_linear(cat(x, skip))] ; x = x + proj(attn(qkv(LN1(x)))) ; x = x + fc2(gelu(fc1(LN2(x))))
// Model (uvit.py:220-250): tokens = [label, time, patches] + pos ; depth/2 in-blocks (outputs pushed as skips) ; mid ;
//                          depth/2 out-blocks (pop skip) ; LN ; decoder_pred ; drop extras ; unpatchify ; 3x3 conv.
// Layout conventions are those of dit_engine.cu: flat fp32 parameters P (+ bf16 shadow Pb), flat fp32 gradients G,
// one caller-owned workspace, fp32 residual stream, bf16 GEMM operands.
#include "engine_common.cuh"

namespace {

constexpr int kMaxBlocks = 64;
constexpr float kLnEps = 1e-5f;  // nn.LayerNorm default (uvit.py:141 norm_layer=nn.LayerNorm)

enum HeadParam : int { U_PE_W = 0, U_PE_B, U_LABEL, U_POS, U_NORM_W, U_NORM_B, U_DEC_W, U_DEC_B, U_CONV_W, U_CONV_B, U_BLOCK0 };
enum UBlockParam : int { UB_N1_W = 0, UB_N1_B, UB_QKV_W, UB_PROJ_W, UB_PROJ_B, UB_N2_W, UB_N2_B, UB_FC1_W, UB_FC1_B,
                         UB_FC2_W, UB_FC2_B, UB_SKIP_W, UB_SKIP_B, UB_COUNT };

struct ULayout {
  long long off[U_BLOCK0 + kMaxBlocks * UB_COUNT];
  long long numel[U_BLOCK0 + kMaxBlocks * UB_COUNT];
  int n;
  long long total;
};

long long ualign64(long long x) { return (x + 63) / 64 * 64; }
inline bool has_skip(const vaw_uvit_cfg& c, int blk) { return blk > c.depth / 2; }  // out-blocks follow in + mid

void u_layout(const vaw_uvit_cfg& c, ULayout& L) {
  const long long D = c.D, Kp = (long long)c.C * c.P * c.P;
  long long cur = 0;
  int n = 0;
  auto add = [&](long long numel) {
    L.off[n] = cur;
    L.numel[n] = numel;
    cur = ualign64(cur + numel);
    ++n;
  };
  add(D * Kp); add(D);                       // patch_embed.proj
  add((long long)c.table_rows * D);          // label_emb
  add((long long)c.T * D);                   // pos_embed (trainable)
  add(D); add(D);                            // norm
  add(Kp * D); add(Kp);                      // decoder_pred
  add(c.conv ? (long long)c.C * c.C * 9 : 0); add(c.conv ? c.C : 0);  // final_layer (3x3 conv)
  for (int i = 0; i < c.depth; ++i) {
    add(D); add(D);                          // norm1
    add(3 * D * D);                          // attn.qkv (no bias, uvit.py:62)
    add(D * D); add(D);                      // attn.proj
    add(D); add(D);                          // norm2
    add((long long)c.hidden * D); add(c.hidden);
    add(D * (long long)c.hidden); add(D);
    const bool sk = has_skip(c, i);
    add(sk ? 2 * D * D : 0); add(sk ? D : 0);  // skip_linear
  }
  L.n = n;
  L.total = cur;
}

struct UBlockWs {
  float *x_in, *x_mid, *mean1, *rstd1, *mean2, *rstd2, *lse;
  bf16 *xn1, *qkv, *attn_o, *xn2, *h_pre, *h_act, *cat, *dcat;
};
struct UWs {
  bf16* patches;
  float *patch_tok, *x0, *x_last;
  UBlockWs blk[kMaxBlocks];
  float *meanf, *rstdf, *tokp, *img;
  bf16* xnf;
  // backward temporaries
  float *dx, *dimg;
  bf16 *dy, *dh, *dqkv, *d_o, *dxn, *dtokp, *dtok;
  float* attn_delta;   // [B*H*T] rowsum(dO*O) scratch of the attention backward
  float *part, *cpart, *split_ws;
  long long split_elems, bytes;
};

int u_chunks(const vaw_uvit_cfg& c) {
  int ch = (c.T + 31) / 32;
  while ((c.T + ch - 1) / ch > 64) ++ch;
  return ch;
}

void u_carve(const vaw_uvit_cfg& c, void* base, UWs& w) {
  Carver k{reinterpret_cast<uint8_t*>(base)};
  const long long B = c.B, D = c.D, M = (long long)c.B * c.T, Hd = c.hidden;
  const long long Lp = c.T - c.extras, Kp = (long long)c.C * c.P * c.P;
  w.patches = k.take<bf16>(B * Lp * Kp);
  w.patch_tok = k.take<float>(B * Lp * D);
  w.x0 = k.take<float>(M * D);
  for (int i = 0; i < c.depth; ++i) {
    UBlockWs& b = w.blk[i];
    const bool sk = has_skip(c, i);
    b.x_in = sk ? k.take<float>(M * D) : nullptr;  // out-blocks: the skip_linear output; others alias the previous x_out
    b.x_mid = k.take<float>(M * D);
    b.mean1 = k.take<float>(M); b.rstd1 = k.take<float>(M); b.mean2 = k.take<float>(M); b.rstd2 = k.take<float>(M);
    b.lse = k.take<float>(B * c.H * c.T);
    b.xn1 = k.take<bf16>(M * D); b.qkv = k.take<bf16>(M * 3 * D); b.attn_o = k.take<bf16>(M * D);
    // xn2 rows carry a [1, 0 x 31] block after their D values: the fc1 weight-gradient GEMM reads them as [M, D + 32] and
    // returns the bias gradient as output column D (row-sum form of VAW_EPI_F32)
    b.xn2 = k.take<bf16>(M * (D + 32)); b.h_pre = k.take<bf16>(M * Hd); b.h_act = k.take<bf16>(M * Hd);
    b.cat = sk ? k.take<bf16>(M * 2 * D) : nullptr;
    b.dcat = sk ? k.take<bf16>(M * 2 * D) : nullptr;
  }
  // x_out of block i lives in xs[i]; x_in of a non-skip block i is xs[i-1] (x0 for block 0)
  w.x_last = nullptr;
  w.meanf = k.take<float>(M); w.rstdf = k.take<float>(M);
  w.xnf = k.take<bf16>(M * D);
  w.tokp = k.take<float>(M * Kp);
  w.img = k.take<float>(B * c.C * c.img_h * c.img_w);
  w.dx = k.take<float>(M * D);
  w.dimg = k.take<float>(B * c.C * c.img_h * c.img_w);
  w.dy = k.take<bf16>(M * D);
  w.dh = k.take<bf16>(M * Hd);
  w.dqkv = k.take<bf16>(M * 3 * D);
  w.d_o = k.take<bf16>(M * D);
  w.attn_delta = k.take<float>((long long)c.B * c.H * c.T);
  w.dxn = k.take<bf16>(M * D);
  w.dtokp = k.take<bf16>(M * Kp);
  w.dtok = k.take<bf16>(B * Lp * D);
  w.part = k.take<float>(B * u_chunks(c) * 2 * D);
  const long long maxN = 3 * D > Hd ? 3 * D : Hd;
  const long long cchunks = (M + colsum_rows((int)M, 64) - 1) / colsum_rows((int)M, 64) + 1;
  w.cpart = k.take<float>(cchunks * maxN + 1024);
  w.split_elems = (long long)vaw_num_sms() * 128 * 256;
  w.split_ws = k.take<float>(w.split_elems);
  w.bytes = k.cur;
}

// x_out buffers: one fp32 [M, D] per block, carved after everything else so the struct above stays simple
struct UX {
  float* xs[kMaxBlocks];
};
void u_carve_x(const vaw_uvit_cfg& c, void* base, long long offset, UX& x, long long& bytes) {
  Carver k{reinterpret_cast<uint8_t*>(base)};
  k.cur = offset;
  for (int i = 0; i < c.depth; ++i) x.xs[i] = k.take<float>((long long)c.B * c.T * c.D);
  bytes = k.cur;
}

int u_check(const vaw_uvit_cfg* c) {
  VAW_CHECK_ARG(c, "uvit: null config");
  VAW_CHECK_ARG(c->B > 0 && c->T > 0 && c->D > 0 && c->H > 0 && c->depth > 0 && c->depth <= kMaxBlocks && (c->depth & 1),
                "uvit: bad geometry B=%d T=%d D=%d H=%d depth=%d (depth must be odd)", c->B, c->T, c->D, c->H, c->depth);
  VAW_CHECK_ARG(c->D % c->H == 0 && (c->D / c->H == 64 || c->D / c->H == 72), "uvit: head_dim %d unsupported (64, 72)",
                c->D / c->H);
  VAW_CHECK_ARG(c->D % 8 == 0 && c->hidden % 8 == 0, "uvit: D and hidden must be multiples of 8");
  VAW_CHECK_ARG(c->extras == 1 || c->extras == 2, "uvit: extras must be 1 or 2");
  VAW_CHECK_ARG((c->img_h / c->P) * (c->img_w / c->P) + c->extras == c->T, "uvit: T does not match the patch grid");
  VAW_CHECK_ARG((c->C * c->P * c->P) % 8 == 0, "uvit: patch feature count must be a multiple of 8");
  VAW_CHECK_ARG(c->extras == 1 || c->table_rows > 0, "uvit: class-conditional model needs an embedding table");
  VAW_CHECK_ARG(c->C <= 8, "uvit: at most 8 image channels");
  return VAW_OK;
}

}  // namespace

extern "C" int vaw_uvit_param_layout(const vaw_uvit_cfg* cfg, long long* offsets, long long* numels, int cap, int* n_out,
                                     long long* total) {
  TRY(u_check(cfg));
  ULayout L;
  u_layout(*cfg, L);
  VAW_CHECK_ARG(cap >= L.n, "vaw_uvit_param_layout: need room for %d entries", L.n);
  for (int i = 0; i < L.n; ++i) {
    if (offsets) offsets[i] = L.off[i];
    if (numels) numels[i] = L.numel[i];
  }
  if (n_out) *n_out = L.n;
  if (total) *total = L.total;
  return VAW_OK;
}

extern "C" int vaw_uvit_workspace_bytes(const vaw_uvit_cfg* cfg, long long* bytes) {
  TRY(u_check(cfg));
  VAW_CHECK_ARG(bytes, "vaw_uvit_workspace_bytes: null output");
  UWs w;
  u_carve(*cfg, nullptr, w);
  UX x;
  long long total = 0;
  u_carve_x(*cfg, nullptr, w.bytes, x, total);
  *bytes = total;
  return VAW_OK;
}

// x_t [B,C,H,W] fp32, t [B] fp32, y [B] int64 (extras == 2) -> out [B,C,H,W] fp32
extern "C" int vaw_uvit_forward(const vaw_uvit_cfg* cfg, const float* P, const void* Pb_, void* ws_, const float* x_t,
                                const float* t, const long long* y, float* out, cudaStream_t s) {
  TRY(u_check(cfg));
  VAW_CHECK_ARG(P && Pb_ && ws_ && x_t && t && out, "vaw_uvit_forward: null pointer");
  const vaw_uvit_cfg& c = *cfg;
  VAW_CHECK_ARG(c.extras == 1 || y, "vaw_uvit_forward: labels required");
  ULayout L;
  u_layout(c, L);
  UWs w;
  u_carve(c, ws_, w);
  UX X;
  long long tot;
  u_carve_x(c, ws_, w.bytes, X, tot);
  const bf16* Pb = reinterpret_cast<const bf16*>(Pb_);
  const int B = c.B, T = c.T, D = c.D, M = B * T, Hd = c.hidden, hd = D / c.H;
  const int Lp = T - c.extras, Kp = c.C * c.P * c.P;

  TRY(vaw_patchify_in(x_t, w.patches, B, c.C, c.img_h, c.img_w, c.P, s));
  TRY(G(w.patches, Kp, 0, Pb + L.off[U_PE_W], Kp, 0, B * Lp, D, Kp, VAW_EPI_F32).out(w.patch_tok).bias(P + L.off[U_PE_B]).run(s));
  TRY(vaw_uvit_assemble(w.patch_tok, t, c.extras == 2 ? P + L.off[U_LABEL] : nullptr, y, P + L.off[U_POS], w.x0, B, T,
                        c.extras, D, s));
  const int n_in = c.depth / 2;
  const float* x = w.x0;
  for (int i = 0; i < c.depth; ++i) {
    UBlockWs& b = w.blk[i];
    const int pb = U_BLOCK0 + i * UB_COUNT;
    if (has_skip(c, i)) {
      const float* skip = X.xs[n_in - 1 - (i - n_in - 1)];  // out-block j pops the output of in-block n_in-1-j
      TRY(vaw_cat_cast(x, skip, b.cat, M, D, s));
      TRY(G(b.cat, 2LL * D, 0, Pb + L.off[pb + UB_SKIP_W], 2LL * D, 0, M, D, 2 * D, VAW_EPI_F32)
              .out(b.x_in).bias(P + L.off[pb + UB_SKIP_B]).run(s));
      x = b.x_in;
    }
    TRY(vaw_ln_fwd(x, nullptr, nullptr, 0, 1, P + L.off[pb + UB_N1_W], P + L.off[pb + UB_N1_B], b.xn1, b.mean1, b.rstd1,
                   M, D, kLnEps, s));
    TRY(G(b.xn1, D, 0, Pb + L.off[pb + UB_QKV_W], D, 0, M, 3 * D, D, VAW_EPI_BF16).out(b.qkv).run(s));
    TRY(vaw_attn_fwd(b.qkv, b.attn_o, b.lse, B, T, c.H, hd, s));
    TRY(G(b.attn_o, D, 0, Pb + L.off[pb + UB_PROJ_W], D, 0, M, D, D, VAW_EPI_RES)
            .out(nullptr, b.x_mid).bias(P + L.off[pb + UB_PROJ_B]).resid(x).run(s));
    TRY(vaw_ln_fwd_ex(b.x_mid, nullptr, nullptr, 0, nullptr, nullptr, nullptr, 0, 1, P + L.off[pb + UB_N2_W],
                      P + L.off[pb + UB_N2_B], b.xn2, D + 32, 1, b.mean2, b.rstd2, M, D, kLnEps, s));
    TRY(G(b.xn2, D + 32, 0, Pb + L.off[pb + UB_FC1_W], D, 0, M, Hd, D, VAW_EPI_GELU_ERF)
            .out(b.h_pre, b.h_act).bias(P + L.off[pb + UB_FC1_B]).run(s));
    TRY(G(b.h_act, Hd, 0, Pb + L.off[pb + UB_FC2_W], Hd, 0, M, D, Hd, VAW_EPI_RES)
            .out(nullptr, X.xs[i]).bias(P + L.off[pb + UB_FC2_B]).resid(b.x_mid).run(s));
    x = X.xs[i];
  }
  TRY(vaw_ln_fwd(x, nullptr, nullptr, 0, 1, P + L.off[U_NORM_W], P + L.off[U_NORM_B], w.xnf, w.meanf, w.rstdf, M, D,
                 kLnEps, s));
  TRY(G(w.xnf, D, 0, Pb + L.off[U_DEC_W], D, 0, M, Kp, D, VAW_EPI_F32).out(w.tokp).bias(P + L.off[U_DEC_B]).run(s));
  float* img = c.conv ? w.img : out;
  TRY(vaw_unpatchify_strided(w.tokp, 0, img, 0, B, c.C, c.img_h, c.img_w, c.P, 1, c.extras, T, 0, s));
  if (c.conv) TRY(vaw_conv3x3(w.img, P + L.off[U_CONV_W], P + L.off[U_CONV_B], out, B, c.C, c.img_h, c.img_w, 0, s));
  return VAW_OK;
}

// dout [B,C,H,W] fp32.  accumulate = 0: G is overwritten for every tensor; 1: gradients are added.
// events: optional array of depth + 1 cudaEvent_t; events[i] is recorded once block i's thirteen tensors are final
// (the blocks finish in reverse order - the long skip connections only route activation gradients, so an out-block's
// parameters do not wait for its partner in-block), events[depth] after the embedder / head tensors.
extern "C" int vaw_uvit_backward_ev(const vaw_uvit_cfg* cfg, const float* P, const void* Pb_, float* Gd, void* ws_,
                                    const float* dout, const long long* y, int accumulate, void** events,
                                    cudaStream_t s) {
  TRY(u_check(cfg));
  VAW_CHECK_ARG(P && Pb_ && Gd && ws_ && dout, "vaw_uvit_backward: null pointer");
  const vaw_uvit_cfg& c = *cfg;
  ULayout L;
  u_layout(c, L);
  UWs w;
  u_carve(c, ws_, w);
  UX X;
  long long tot;
  u_carve_x(c, ws_, w.bytes, X, tot);
  const bf16* Pb = reinterpret_cast<const bf16*>(Pb_);
  const int B = c.B, T = c.T, D = c.D, M = B * T, Hd = c.hidden, hd = D / c.H;
  const int Lp = T - c.extras, Kp = c.C * c.P * c.P;
  const int ch = u_chunks(c), acc = accumulate ? 1 : 0;
  const int n_in = c.depth / 2;

  // ---- final conv, decoder_pred, final norm ----
  const float* dimg = dout;
  if (c.conv) {
    TRY(vaw_conv3x3_wgrad(w.img, dout, Gd + L.off[U_CONV_W], Gd + L.off[U_CONV_B], B, c.C, c.img_h, c.img_w, acc, s));
    TRY(vaw_conv3x3(dout, P + L.off[U_CONV_W], nullptr, w.dimg, B, c.C, c.img_h, c.img_w, 1, s));
    dimg = w.dimg;
  }
  TRY(vaw_unpatchify_strided(w.dtokp, 1, const_cast<float*>(dimg), 0, B, c.C, c.img_h, c.img_w, c.P, 0, c.extras, T, 1, s));
  TRY(vaw_colsum_bf16(w.dtokp, Kp, M, Kp, w.cpart, colsum_rows(M, Kp), Gd + L.off[U_DEC_B], acc, s));
  TRY(G(w.dtokp, Kp, 1, w.xnf, D, 1, Kp, D, M, VAW_EPI_F32).out(Gd + L.off[U_DEC_W]).acc(acc)
          .autosplit(w.split_ws, w.split_elems).run(s));
  TRY(G(w.dtokp, Kp, 0, Pb + L.off[U_DEC_W], D, 1, M, D, Kp, VAW_EPI_BF16).out(w.dxn).run(s));
  TRY(vaw_ln_bwd(w.dxn, X.xs[c.depth - 1], w.meanf, w.rstdf, nullptr, 0, P + L.off[U_NORM_W], w.dx, 0, w.part, T, B, ch,
                 M, D, s));
  TRY(vaw_finish_all(w.part, 0, B, ch, D, nullptr, 0, Gd + L.off[U_NORM_B], acc, s));
  TRY(vaw_finish_all(w.part, 1, B, ch, D, nullptr, 0, Gd + L.off[U_NORM_W], acc, s));

  for (int i = c.depth - 1; i >= 0; --i) {
    UBlockWs& b = w.blk[i];
    const int pb = U_BLOCK0 + i * UB_COUNT;
    const bool sk = has_skip(c, i);
    const float* x_in = sk ? b.x_in : (i == 0 ? w.x0 : X.xs[i - 1]);
    if (i < n_in) {
      // this in-block's output was also consumed as the skip of out-block (depth - 1 - i): add that gradient
      const int partner = c.depth - 1 - i;
      TRY(vaw_unpack_cols(w.blk[partner].dcat, 2LL * D, D, w.dx, M, D, 1, s));
    }
    // ---- MLP branch ----
    TRY(vaw_gate_bwd(w.dx, nullptr, nullptr, 0, w.dy, w.part, T, B, ch, M, D, s));
    TRY(vaw_finish_all(w.part, 0, B, ch, D, nullptr, 0, Gd + L.off[pb + UB_FC2_B], acc, s));
    TRY(G(w.dy, D, 1, b.h_act, Hd, 1, D, Hd, M, VAW_EPI_F32).out(Gd + L.off[pb + UB_FC2_W]).acc(acc)
            .autosplit(w.split_ws, w.split_elems).run(s));
    TRY(G(w.dy, D, 0, Pb + L.off[pb + UB_FC2_W], Hd, 1, M, Hd, D, VAW_EPI_DGELU_ERF).out(w.dh).aux(b.h_pre).run(s));
    TRY(G(w.dh, Hd, 1, b.xn2, D + 32, 1, Hd, D + 32, M, VAW_EPI_F32)   // fc1 weight + bias gradient in one GEMM
            .out(Gd + L.off[pb + UB_FC1_W], Gd + L.off[pb + UB_FC1_B]).acc(acc).autosplit(w.split_ws, w.split_elems).run(s));
    TRY(G(w.dh, Hd, 0, Pb + L.off[pb + UB_FC1_W], D, 1, M, D, Hd, VAW_EPI_BF16).out(w.dxn).run(s));
    TRY(vaw_ln_bwd(w.dxn, b.x_mid, b.mean2, b.rstd2, nullptr, 0, P + L.off[pb + UB_N2_W], w.dx, 1, w.part, T, B, ch, M, D, s));
    TRY(vaw_finish_all(w.part, 0, B, ch, D, nullptr, 0, Gd + L.off[pb + UB_N2_B], acc, s));
    TRY(vaw_finish_all(w.part, 1, B, ch, D, nullptr, 0, Gd + L.off[pb + UB_N2_W], acc, s));
    // ---- attention branch ----
    TRY(vaw_gate_bwd(w.dx, nullptr, nullptr, 0, w.dy, w.part, T, B, ch, M, D, s));
    TRY(vaw_finish_all(w.part, 0, B, ch, D, nullptr, 0, Gd + L.off[pb + UB_PROJ_B], acc, s));
    TRY(G(w.dy, D, 1, b.attn_o, D, 1, D, D, M, VAW_EPI_F32).out(Gd + L.off[pb + UB_PROJ_W]).acc(acc)
            .autosplit(w.split_ws, w.split_elems).run(s));
    TRY(G(w.dy, D, 0, Pb + L.off[pb + UB_PROJ_W], D, 1, M, D, D, VAW_EPI_BF16).out(w.d_o).run(s));
    TRY(vaw_attn_bwd_ws(b.qkv, b.attn_o, w.d_o, b.lse, w.dqkv, w.attn_delta, B, T, c.H, hd, s));
    TRY(G(w.dqkv, 3LL * D, 1, b.xn1, D, 1, 3 * D, D, M, VAW_EPI_F32).out(Gd + L.off[pb + UB_QKV_W]).acc(acc)
            .autosplit(w.split_ws, w.split_elems).run(s));
    TRY(G(w.dqkv, 3LL * D, 0, Pb + L.off[pb + UB_QKV_W], D, 1, M, D, 3 * D, VAW_EPI_BF16).out(w.dxn).run(s));
    TRY(vaw_ln_bwd(w.dxn, x_in, b.mean1, b.rstd1, nullptr, 0, P + L.off[pb + UB_N1_W], w.dx, 1, w.part, T, B, ch, M, D, s));
    TRY(vaw_finish_all(w.part, 0, B, ch, D, nullptr, 0, Gd + L.off[pb + UB_N1_B], acc, s));
    TRY(vaw_finish_all(w.part, 1, B, ch, D, nullptr, 0, Gd + L.off[pb + UB_N1_W], acc, s));
    // ---- skip_linear: x_in = [x_prev, skip] W^T + b ----
    if (sk) {
      TRY(vaw_gate_bwd(w.dx, nullptr, nullptr, 0, w.dy, w.part, T, B, ch, M, D, s));
      TRY(vaw_finish_all(w.part, 0, B, ch, D, nullptr, 0, Gd + L.off[pb + UB_SKIP_B], acc, s));
      TRY(G(w.dy, D, 1, b.cat, 2LL * D, 1, D, 2 * D, M, VAW_EPI_F32).out(Gd + L.off[pb + UB_SKIP_W]).acc(acc)
              .autosplit(w.split_ws, w.split_elems).run(s));
      TRY(G(w.dy, D, 0, Pb + L.off[pb + UB_SKIP_W], 2LL * D, 1, M, 2 * D, D, VAW_EPI_BF16).out(b.dcat).run(s));
      TRY(vaw_unpack_cols(b.dcat, 2LL * D, 0, w.dx, M, D, 0, s));  // gradient of the block's x input (first D columns)
    }
    if (events && events[i]) VAW_CUDA_TRY(cudaEventRecord(reinterpret_cast<cudaEvent_t>(events[i]), s));
  }
  // ---- token assembly: pos_embed, label table, patch embedding ----
  TRY(vaw_uvit_pos_grad(w.dx, Gd + L.off[U_POS], B, T, D, acc, s));
  if (c.extras == 2)
    TRY(vaw_embedding_grad_strided(w.dx, (long long)T * D, y, Gd + L.off[U_LABEL], c.table_rows, B, D, acc, s));
  TRY(vaw_uvit_gather_patch_grad(w.dx, w.dtok, B, T, c.extras, D, s));
  TRY(vaw_colsum_bf16(w.dtok, D, B * Lp, D, w.cpart, colsum_rows(B * Lp, D), Gd + L.off[U_PE_B], acc, s));
  TRY(G(w.dtok, D, 1, w.patches, Kp, 1, D, Kp, B * Lp, VAW_EPI_F32).out(Gd + L.off[U_PE_W]).acc(acc)
          .autosplit(w.split_ws, w.split_elems).run(s));
  if (events && events[c.depth]) VAW_CUDA_TRY(cudaEventRecord(reinterpret_cast<cudaEvent_t>(events[c.depth]), s));
  return VAW_OK;
}

extern "C" int vaw_uvit_backward(const vaw_uvit_cfg* cfg, const float* P, const void* Pb_, float* Gd, void* ws_,
                                 const float* dout, const long long* y, int accumulate, cudaStream_t s) {
  return vaw_uvit_backward_ev(cfg, P, Pb_, Gd, ws_, dout, y, accumulate, nullptr, s);
}

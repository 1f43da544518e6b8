// vaw_internal.h — declarations shared between the translation units of libvaw_b200.so.
// The structs here are layout-identical to the public ones in include/vaw_b200.h (checked by tests/test_abi.py).
#pragma once
#include <cuda_runtime.h>

struct vaw_gemm_args {
  const void* A;       // bf16; K-major: [M,K] row-major (lda) ; MN-major: [K,M] row-major (lda)
  const void* B;       // bf16; K-major: [N,K] row-major (ldb) ; MN-major: [K,N] row-major (ldb)
  long long lda, ldb;
  int a_mn, b_mn;
  int M, N, K;
  int epilogue;
  void* out;           // primary output  [M,N] (ldo)
  void* out2;          // secondary output [M,N] (ldo)
  const float* bias;   // [N] or null
  const float* resid;  // [M,N] fp32 (ldo)
  const float* gate;   // [M / rows_per_sample, >=N] fp32 (ldg)
  const void* aux;     // bf16 [M,N] (ldo): saved pre-activation for the d-activation epilogues
  long long ldo, ldg;
  int rows_per_sample;
  int accumulate;      // EPI_F32: out += result
  int tile_n;          // 0 = auto, else 128 / 192 / 256
  int resid_mod;       // > 0: resid is [resid_mod, N], indexed by row % resid_mod
  int k_splits;        // EPI_F32 only.  > 1: split every tile's K loop; -1: split only the partial last wave
  float* split_ws;     // fp32 scratch for the split partials (slabs of 128 x tile_n)
  long long split_ws_elems;  // capacity of split_ws in floats (0 = unchecked)
  int cta_group;       // 0 = auto, 1 = one CTA per 128-row tile, 2 = SM pair per 256-row tile (tcgen05 cta_group::2)
};

enum : int {
  VAW_EPI_BF16 = 0, VAW_EPI_F32 = 1, VAW_EPI_GELU_TANH = 2, VAW_EPI_GELU_ERF = 3, VAW_EPI_GATE_RES = 4,
  VAW_EPI_RES = 5, VAW_EPI_DGELU_TANH = 6, VAW_EPI_DGELU_ERF = 7, VAW_EPI_SILU = 8, VAW_EPI_DSILU = 9,
  VAW_EPI_ALIGN_MSE = 10
};

// U-ViT geometry (models/uvit.py:139-205)
struct vaw_uvit_cfg {
  int B, T, D, H, depth, hidden;   // batch, tokens/sample incl. extras, width, heads, blocks (odd), mlp hidden
  int C, P, img_h, img_w;          // channels, patch size, image size
  int extras, table_rows;          // 1 (time token) or 2 (label + time); label-embedding rows
  int conv;                        // 1: final 3x3 convolution (uvit.py:192)
};

// DiT geometry (models/dit.py:157-204)
struct vaw_dit_cfg {
  int B, T, D, H, depth, hidden;        // batch, tokens/sample, width, heads, blocks, mlp hidden
  int C_in, C_out, P, img_h, img_w;     // channels, patch size, image size
  int table_rows, freq_dim;             // label-embedding rows (0 = unconditional), sinusoid width (256)
  int learn_align, encoder_depth, proj_dim, z_dim;  // REPA projector (dit.py:27-34,201,274-275)
};

// tcgen05 attention (attention_sm100.cu); VAW_ERR_UNSUPPORTED when the shape is outside its range (T > 256)
int vaw_attn_fwd_sm100(const void* qkv, void* o, float* lse2, int B, int T, int H, int head_dim, cudaStream_t stream);
int vaw_attn_bwd_sm100(const void* qkv, const void* o, const void* d_o, const float* lse2, void* dqkv, float* delta_ws,
                       int B, int T, int H, int head_dim, cudaStream_t stream);

// sequences slightly longer than the tensor-core tile (attention_border.cu): T in (256, 264]
int vaw_attn_border_supported(int T);
int vaw_attn_border_fwd(const void* qkv, void* o, float* lse2, int B, int T, int H, int head_dim, cudaStream_t stream);
int vaw_attn_border_bwd(const void* qkv, const void* d_o, const float* lse2, const float* delta, void* dqkv, int B, int T,
                        int H, int head_dim, cudaStream_t stream);

extern "C" {
int vaw_wgrad_smallk(const void* A, long long lda, const void* B, long long ldb, float* out, long long ldo, int M, int N,
                     int K, int accumulate, cudaStream_t stream);
int vaw_align_mse_finish(const float* part, long long nparts, long long n, float* loss, cudaStream_t stream);
int vaw_gemm_bf16(const vaw_gemm_args* a, cudaStream_t stream);
int vaw_attn_fwd(const void* qkv, void* o, float* lse2, int B, int T, int H, int head_dim, cudaStream_t stream);
int vaw_attn_bwd(const void* qkv, const void* o, const void* d_o, const float* lse2, void* dqkv, int B, int T, int H,
                 int head_dim, cudaStream_t stream);
int vaw_attn_bwd_ws(const void* qkv, const void* o, const void* d_o, const float* lse2, void* dqkv, float* delta_ws,
                    int B, int T, int H, int head_dim, cudaStream_t stream);
int vaw_ln_fwd(const float* x, const float* shift, const float* scale, long long ld_mod, int rows_per_sample,
               const float* weight, const float* bias, void* y, float* mean, float* rstd, int M, int D, float eps,
               cudaStream_t stream);
int vaw_ln_fwd_res(const float* x, const void* branch, const float* gate, long long ld_gate, float* x_out,
                   const float* shift, const float* scale, long long ld_mod, int rows_per_sample, void* y, float* mean,
                   float* rstd, int M, int D, float eps, cudaStream_t stream);
int vaw_ln_fwd_ex(const float* x, const void* branch, const float* gate, long long ld_gate, float* x_out,
                  const float* shift, const float* scale, long long ld_mod, int rows_per_sample, const float* weight,
                  const float* bias, void* y, long long ldy, int ones_block, float* mean, float* rstd, int M, int D,
                  float eps, cudaStream_t stream);
int vaw_ln_bwd(const void* dy, const float* x, const float* mean, const float* rstd, const float* scale,
               long long ld_mod, const float* weight, float* dx_io, int add_into, float* part, int rows_per_group,
               int groups, int chunks, int M, int D, cudaStream_t stream);
int vaw_ln_bwd_gate(const void* dy, const float* x, const float* mean, const float* rstd, const float* scale,
                    long long ld_mod, const float* weight, float* dx_io, int add_into, float* part, const void* y_next,
                    const float* gate_next, long long ld_gate, void* dy_next, float* part_gate, int rows_per_group,
                    int groups, int chunks, int M, int D, cudaStream_t stream);
int vaw_gate_bwd(const float* dx, const void* y, const float* gate, long long ld_gate, void* dy, float* part,
                 int rows_per_group, int groups, int chunks, int M, int D, cudaStream_t stream);
int vaw_finish_group(const float* part, int which, int groups, int chunks, int D, float* out, long long ld_out,
                     int accumulate, cudaStream_t stream);
int vaw_finish_all(const float* part, int which, int groups, int chunks, int D, const float* w, long long ld_w,
                   float* out, int accumulate, cudaStream_t stream);
int vaw_dit_block_finish(const float* pA, const float* pB, const float* pC, const float* pD, int B, int chunks, int D,
                         const float* mod, long long ldm, float* dmod, void* dmod_b, float* g_fc2_b, float* g_proj_b,
                         float* g_ada_b, int accumulate, cudaStream_t stream);
int vaw_colsum_bf16(const void* a, long long lda, int M, int N, float* part, int rows_per_chunk, float* out,
                    int accumulate, cudaStream_t stream);
int vaw_patchify_in(const float* x, void* patches, int B, int C, int H, int W, int P, cudaStream_t stream);
int vaw_unpatchify(void* tokens, void* image, int dtype, int B, int C, int H, int W, int P, int to_image,
                   cudaStream_t stream);
int vaw_timestep_embedding(const float* t, void* out_bf16, float* out_f32, int B, int dim, cudaStream_t stream);
int vaw_cond_combine(const float* t_emb, const float* table, const long long* labels, float* c, void* c_silu, int B,
                     int D, cudaStream_t stream);
int vaw_cond_bwd(const float* dc_silu, const float* c, float* dc, void* dc_bf16, int n, cudaStream_t stream);
int vaw_embedding_grad(const float* dc, const long long* labels, float* dtable, int rows, int B, int D, int accumulate,
                       cudaStream_t stream);
int vaw_embedding_grad_strided(const float* dc, long long ld, const long long* labels, float* dtable, int rows, int B,
                               int D, int accumulate, cudaStream_t stream);
int vaw_cast_f32_bf16(const float* src, void* dst, long long n, cudaStream_t stream);
int vaw_uvit_assemble(const float* patch_tok, const float* t, const float* table, const long long* labels,
                      const float* pos, float* x0, int B, int T, int extras, int D, cudaStream_t stream);
int vaw_uvit_pos_grad(const float* dx0, float* dpos, int B, int T, int D, int accumulate, cudaStream_t stream);
int vaw_uvit_gather_patch_grad(const float* dx0, void* dtok, int B, int T, int extras, int D, cudaStream_t stream);
int vaw_cat_cast(const float* x, const float* skip, void* cat, long long M, int D, cudaStream_t stream);
int vaw_unpack_cols(const void* src, long long ld, int col_off, float* dst, long long M, int D, int accumulate,
                    cudaStream_t stream);
int vaw_unpatchify_strided(void* tokens, int tok_dtype, void* image, int img_dtype, int B, int C, int H, int W, int P,
                           int to_image, int row0, int rows_per_sample, int zero_extras, cudaStream_t stream);
int vaw_conv3x3(const float* in, const float* w, const float* bias, float* out, int B, int C, int H, int W,
                int transpose, cudaStream_t stream);
int vaw_conv3x3_wgrad(const float* in, const float* dout, float* dw, float* dbias, int B, int C, int H, int W,
                      int accumulate, cudaStream_t stream);
int vaw_cast_f32_bf16_2d(const float* src, long long lds, void* dst, long long ldd, int rows, int cols,
                         cudaStream_t stream);
int vaw_add_bf16_into_f32(const void* src, float* dst, long long n, cudaStream_t stream);
int vaw_colsum_f32_small(const float* a, long long lda, int rows, int N, float* out, int accumulate,
                         cudaStream_t stream);
}

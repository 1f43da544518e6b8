// engine_common.cuh — helpers shared by the model engines (dit_engine.cu, uvit_engine.cu): workspace carving, the GEMM
// call builder and the status-propagation macro.
#pragma once
#include "vaw_common.cuh"
#include "vaw_internal.h"

namespace {

struct Carver {
  uint8_t* base;
  long long cur = 0;
  template <typename T>
  T* take(long long n) {
    T* p = base ? reinterpret_cast<T*>(base + cur) : nullptr;
    cur += (n * (long long)sizeof(T) + 255) / 256 * 256;
    return p;
  }
};

struct G {  // small GEMM call builder
  vaw_gemm_args a;
  G(const void* A, long long lda, int a_mn, const void* B, long long ldb, int b_mn, int M, int N, int K, int epi) {
    memset(&a, 0, sizeof(a));
    a.A = A; a.lda = lda; a.a_mn = a_mn; a.B = B; a.ldb = ldb; a.b_mn = b_mn;
    a.M = M; a.N = N; a.K = K; a.epilogue = epi;
  }
  G& out(void* o, void* o2 = nullptr) { a.out = o; a.out2 = o2; return *this; }
  G& bias(const float* b) { a.bias = b; return *this; }
  G& resid(const float* r, int mod = 0) { a.resid = r; a.resid_mod = mod; return *this; }
  G& gate(const float* g, long long ldg, int rps) { a.gate = g; a.ldg = ldg; a.rows_per_sample = rps; return *this; }
  G& aux(const void* x) { a.aux = x; return *this; }
  G& acc(int f) { a.accumulate = f; return *this; }
  // split-K policy ("tail split"): whole tiles for the full waves, the partial last wave (or, for skinny GEMMs, every
  // tile) split along K so that all SMs stay busy; partials are folded in fixed order
  G& autosplit(float* ws, long long ws_elems) {
    a.split_ws = ws;
    a.split_ws_elems = ws_elems;
    a.k_splits = -1;
    return *this;
  }
  int run(cudaStream_t s) { return vaw_gemm_bf16(&a, s); }
};

#define TRY(expr)            \
  do {                       \
    int _rc = (expr);        \
    if (_rc != VAW_OK) return _rc; \
  } while (0)


inline int colsum_rows(int M, int N) {
  const int strips = (N + 63) / 64;
  int chunks = (4 * vaw_num_sms() + strips - 1) / strips;
  if (chunks < 1) chunks = 1;
  int rows = (M + chunks - 1) / chunks;
  rows = (rows + 7) / 8 * 8;
  if (rows < 8) rows = 8;
  return rows;
}

}  // namespace

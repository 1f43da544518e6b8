"""Helpers shared by the -m gpu tests."""
import ctypes as C

import torch

from vaw_b200 import _lib as L


def relerr(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def run_gemm(A, B, a_mn, b_mn, M, N, K, epi, out=None, out2=None, bias=None, resid=None, gate=None, aux=None,
             rows_per_sample=1, accumulate=0, tile_n=0, resid_mod=0, k_splits=0, split_ws=None, cta_group=0):
    g = L.GemmArgs()
    g.A, g.B = A.data_ptr(), B.data_ptr()
    g.lda, g.ldb = A.stride(0), B.stride(0)
    g.a_mn, g.b_mn = a_mn, b_mn
    g.M, g.N, g.K = M, N, K
    g.epilogue = epi
    g.out, g.out2, g.bias, g.resid = L.ptr(out), L.ptr(out2), L.ptr(bias), L.ptr(resid)
    g.gate, g.aux = L.ptr(gate), L.ptr(aux)
    g.rows_per_sample, g.accumulate, g.tile_n, g.resid_mod = rows_per_sample, accumulate, tile_n, resid_mod
    g.k_splits, g.split_ws = k_splits, L.ptr(split_ws)
    g.split_ws_elems = split_ws.numel() if split_ws is not None else 0
    g.cta_group = cta_group
    L.call("vaw_gemm_bf16", C.byref(g), L.stream_ptr())


def dezero(model):
    with torch.no_grad():
        for p in model.parameters():
            if p.requires_grad and p.abs().sum() == 0:
                p.normal_(0, 0.02)

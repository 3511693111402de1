"""Shared builders for the implicit-GEMM engine tests (GPU).  Each case returns
(desc-builder kwargs, torch fp32 reference computed from the fp16-rounded operands)."""
from __future__ import annotations

import ctypes as C
import math
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "retrieval-based-voice-conversion-webui_b200")
if PKG not in sys.path:
    sys.path.insert(0, PKG)

from rvc_b200 import _lib  # noqa: E402

ACT = dict(none=0, relu=1, gelu=2, lrelu=3, tanh=4, sigmoid=5)


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def run_gemm(impl, A, B, M, N, segs, block_k=64, a_rows=None, a_cols=None, lda=None, conv2d_W=0, b_rows=None, b_cols=None,
             ldb=None, batch=1, a_row_z=0, a_col_z=0, b_row_z=0, b_col_z=0, c_z=0, bias_z=0, b_col0=0, bias=None,
             bias_per_row=0, res1=None, res2=None, alpha=1.0, act1="none", act1_p=0.0, act2="none", act2_p=0.0, gate=0,
             out32=None, ld32=0, out16=None, ld16=0, up2_C=0):
    d = _lib.GemmDesc()
    d.A, d.lda = A.data_ptr(), lda if lda is not None else A.stride(-2)
    d.a_rows = a_rows if a_rows is not None else A.shape[0]
    d.a_cols = a_cols if a_cols is not None else A.shape[-1]
    d.conv2d_W = conv2d_W
    d.B, d.ldb = B.data_ptr(), ldb if ldb is not None else B.stride(0)
    d.b_rows = b_rows if b_rows is not None else B.shape[0]
    d.b_cols = b_cols if b_cols is not None else B.shape[1]
    d.M, d.N, d.block_k, d.nseg = M, N, block_k, len(segs)
    for i, (r, c, dw, nk) in enumerate(segs):
        d.seg[i].row_off, d.seg[i].col_off, d.seg[i].dw, d.seg[i].nk = r, c, dw, nk
    d.batch, d.a_row_z, d.a_col_z, d.b_row_z, d.b_col_z, d.c_z, d.bias_z, d.b_col0 = batch, a_row_z, a_col_z, b_row_z, b_col_z, c_z, bias_z, b_col0
    d.bias = bias.data_ptr() if bias is not None else None
    d.bias_per_row = bias_per_row
    d.res1 = res1.data_ptr() if res1 is not None else None
    d.ldres1 = res1.stride(0) if res1 is not None else 0
    d.res2 = res2.data_ptr() if res2 is not None else None
    d.ldres2 = res2.stride(0) if res2 is not None else 0
    d.alpha, d.act1, d.act1_p, d.act2, d.act2_p, d.gate = alpha, ACT[act1], act1_p, ACT[act2], act2_p, gate
    d.out32 = out32.data_ptr() if out32 is not None else None
    d.ld32 = ld32
    d.out16 = out16.data_ptr() if out16 is not None else None
    d.ld16 = ld16
    d.up2_C = up2_C
    _lib.check(_lib.lib().rvcb_op_gemm(C.byref(d), impl, None))
    torch.cuda.synchronize()


def pack_conv1d(w, bk):
    """w [Cout, Cin, k] -> B [Cout, k*Cin_pad] (fp16), tap-major."""
    co, ci, k = w.shape
    cip = (ci + bk - 1) // bk * bk
    B = torch.zeros(co, k, cip, dtype=torch.float16, device=w.device)
    B[:, :, :ci] = w.permute(0, 2, 1).half()
    return B.reshape(co, k * cip).contiguous()


def pack_conv2d(w, bk):
    """w [Cout, Cin, 3, 3] -> B [Cout, 9*Cin_pad], (dh, dw, ci) order."""
    co, ci, kh, kw = w.shape
    cip = (ci + bk - 1) // bk * bk
    B = torch.zeros(co, kh * kw, cip, dtype=torch.float16, device=w.device)
    B[:, :, :ci] = w.permute(0, 2, 3, 1).reshape(co, kh * kw, ci).half()
    return B.reshape(co, kh * kw * cip).contiguous()


def pack_convT1d(w, stride, pad, bk):
    """w [Cin, Cout, k] (ConvTranspose1d) -> polyphase B [stride*Cout, 3*Cin_pad]; row = r*Cout+co,
    segment delta in (-1,0,+1): tap j = r + pad - delta*stride."""
    ci, co, k = w.shape
    cip = (ci + bk - 1) // bk * bk
    B = torch.zeros(stride, co, 3, cip, dtype=torch.float16, device=w.device)
    for r in range(stride):
        for di, delta in enumerate((-1, 0, 1)):
            j = r + pad - delta * stride
            if 0 <= j < k:
                B[r, :, di, :ci] = w[:, :, j].t().half()
    return B.reshape(stride * co, 3 * cip).contiguous()


def pack_convT2d_up2(w, bk):
    """w [Cin, Cout, 3, 3] ConvTranspose2d(stride 2, pad 1, output_padding 1) -> B [4*Cout, 4*Cin_pad];
    row = (a*2+b)*Cout + co; segments (dh, dw) in ((0,0),(0,1),(1,0),(1,1)); kh = a + 1 - 2*dh, kw = b + 1 - 2*dw."""
    ci, co, _, _ = w.shape
    cip = (ci + bk - 1) // bk * bk
    B = torch.zeros(2, 2, co, 4, cip, dtype=torch.float16, device=w.device)
    for a in range(2):
        for b in range(2):
            for si, (dh, dw) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
                kh, kw = a + 1 - 2 * dh, b + 1 - 2 * dw
                if 0 <= kh < 3 and 0 <= kw < 3:
                    B[a, b, :, si, :ci] = w[:, :, kh, kw].t().half()
    return B.reshape(4 * co, 4 * cip).contiguous()

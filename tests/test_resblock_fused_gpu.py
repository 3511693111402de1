"""GPU op-level parity: the fused residual-block kernel (csrc/resblock_fused.cu) against a plain PyTorch fp32 ResBlock1
(rvc/layers/residuals.py:68-85) and against the same recurrence with the MMA operands rounded to fp16 (what the kernel computes:
fp16 operands, fp32 accumulation, fp32 residual stream)."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ref(x, w1, b1, w2, b2, k, dil, emulate_fp16):
    """x [T, C] -> [T, C]"""
    r16 = (lambda t: t.half().float()) if emulate_fp16 else (lambda t: t)
    y = x.t()[None]
    for i, d in enumerate(dil):
        t = F.conv1d(r16(F.leaky_relu(y, 0.1)), r16(w1[i]), b1[i], dilation=d, padding=(k - 1) // 2 * d)
        t = F.conv1d(r16(F.leaky_relu(t, 0.1)), r16(w2[i]), b2[i], padding=(k - 1) // 2)
        y = y + t
    return y[0].t().contiguous()


def _run(Cc, k, dil, T, seed):
    from rvc_b200 import _lib
    _lib.init(0)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(T, Cc, generator=g)
    w1 = [torch.randn(Cc, Cc, k, generator=g) / np.sqrt(Cc * k) for _ in range(3)]
    w2 = [torch.randn(Cc, Cc, k, generator=g) / np.sqrt(Cc * k) for _ in range(3)]
    b1 = [torch.randn(Cc, generator=g) * 0.1 for _ in range(3)]
    b2 = [torch.randn(Cc, generator=g) * 0.1 for _ in range(3)]
    dil_a = (C.c_int * 3)(*dil)
    rows = _lib.lib().rvcb_op_resblock1_out_rows(Cc, k, dil_a, T)
    assert rows == T
    xd = x.cuda()
    yd = torch.full((rows, Cc), float("nan"), device="cuda")
    arr = lambda ts: (C.c_void_p * 3)(*[t.data_ptr() for t in ts])
    keep = [t.contiguous() for t in w1 + b1 + w2 + b2]
    _lib.check(_lib.lib().rvcb_op_resblock1(Cc, k, dil_a, arr(keep[0:3]), arr(keep[3:6]), arr(keep[6:9]), arr(keep[9:12]),
                                            C.c_void_p(xd.data_ptr()), T, C.c_void_p(yd.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    y = yd[:T].cpu()
    assert torch.isfinite(y).all()
    with torch.no_grad():
        r32 = _ref(x, w1, b1, w2, b2, k, dil, False)
        r16 = _ref(x, w1, b1, w2, b2, k, dil, True)
    scale = r32.abs().max().item()
    e16 = (y - r16).abs().max().item() / scale
    e32 = (y - r32).abs().max().item() / scale
    print(f"[resblock] C={Cc} k={k} T={T}: vs fp16-operand reference {e16:.2e}, vs fp32 reference {e32:.2e} (of max |y| = {scale:.2f})")
    # same arithmetic up to fp32 summation order (and the rare fp16 rounding tie it can flip)
    assert e16 < 5e-4, e16
    assert e32 < 4e-3, e32
    return y, r16


@pytest.mark.parametrize("Cc,k,T", [(32, 3, 5000), (32, 11, 2000), (32, 7, 150000), (64, 7, 3000), (64, 11, 700), (64, 3, 90000),
                                    (64, 11, 40000), (32, 11, 100000)])
def test_fused_resblock_matches_torch(Cc, k, T):
    _run(Cc, k, (1, 3, 5), T, seed=Cc * 100 + k)


def test_fused_resblock_sequence_edges_and_tile_seams():
    """Rows next to t = 0 / t = T-1 (each convolution zero-pads its own input) and rows on both sides of every tile seam."""
    y, r = _run(64, 11, (1, 3, 5), 2 * 392 + 37, seed=5)        # R = 512 - 120 = 392 output rows per tile: 3 tiles, last one partial
    for lo, hi in ((0, 64), (392 - 64, 392 + 64), (784 - 64, 784 + 37)):
        assert (y[lo:hi] - r[lo:hi]).abs().max().item() < 5e-4 * r.abs().max().item()

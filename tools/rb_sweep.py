"""Event-timed launches of the fused residual-block kernel at the vocoder's own shapes (stage 2: C=64, T=383520; stage 3: C=32,
T=767040), k in {3, 7, 11}; variant chosen by RVCB_RB_CFG.  Prints ms and algorithmic TFLOP/s per launch."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "retrieval-based-voice-conversion-webui_b200"))
from rvc_b200 import _lib  # noqa: E402

_lib.init(0)
L = _lib.lib()
g = torch.Generator().manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
tot = 0.0
for Cc, T in ((64, 383520), (32, 767040)):
    x = torch.randn(T, Cc, generator=g).cuda()
    y = torch.empty(T, Cc, device="cuda")
    for k in (3, 7, 11):
        ws = [(torch.randn(Cc, Cc, k, generator=g) / np.sqrt(Cc * k)).contiguous() for _ in range(6)]
        bs = [(torch.randn(Cc, generator=g) * 0.1).contiguous() for _ in range(6)]
        dil = (C.c_int * 3)(1, 3, 5)
        arr = lambda ts: (C.c_void_p * 3)(*[t.data_ptr() for t in ts])
        best = 1e9
        for it in range(4):
            flush.fill_(1)
            _lib.check(L.rvcb_prof_begin())
            _lib.check(L.rvcb_op_resblock1(Cc, k, dil, arr(ws[0:3]), arr(bs[0:3]), arr(ws[3:6]), arr(bs[3:6]), C.c_void_p(x.data_ptr()), T,
                                           C.c_void_p(y.data_ptr()), st))
            ms, n = C.c_double(0), C.c_ulonglong(0)
            _lib.check(L.rvcb_prof_end(C.byref(ms), C.byref(n)))
            if it:
                best = min(best, ms.value)
        fl = 12.0 * T * Cc * Cc * k
        tot += best
        print(f"cfg={os.environ.get('RVCB_RB_CFG', '0')} C={Cc} k={k:2d}: {best * 1e3:7.1f} us  {fl / best / 1e9:7.1f} TFLOP/s")
print(f"cfg={os.environ.get('RVCB_RB_CFG', '0')} total {tot:.3f} ms")

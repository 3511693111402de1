"""One launch of the fused residual-block kernel at a vocoder stage's own shape (for ncu / event timing).
    python tools/rb_prof.py C k T [reps]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "retrieval-based-voice-conversion-webui_b200"))
from rvc_b200 import _lib  # noqa: E402

Cc, k, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
_lib.init(0)
g = torch.Generator().manual_seed(0)
x = torch.randn(T, Cc, generator=g).cuda()
ws = [(torch.randn(Cc, Cc, k, generator=g) / np.sqrt(Cc * k)).contiguous() for _ in range(6)]
bs = [(torch.randn(Cc, generator=g) * 0.1).contiguous() for _ in range(6)]
dil = (C.c_int * 3)(1, 3, 5)
arr = lambda ts: (C.c_void_p * 3)(*[t.data_ptr() for t in ts])
y = torch.empty(T, Cc, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for i in range(reps):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    _lib.check(_lib.lib().rvcb_op_resblock1(Cc, k, dil, arr(ws[0:3]), arr(bs[0:3]), arr(ws[3:6]), arr(bs[3:6]), C.c_void_p(x.data_ptr()), T,
                                            C.c_void_p(y.data_ptr()), st))
    e.record()
    torch.cuda.synchronize()
    fl = 12.0 * T * Cc * Cc * k
    print(f"C={Cc} k={k} T={T}: {s.elapsed_time(e):.3f} ms incl. host packing; algorithmic {fl / 1e9:.1f} GF")

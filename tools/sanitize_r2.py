"""Small invocations of the round-2 kernels for compute-sanitizer (memcheck / racecheck / synccheck): the fused residual block
(streamed and resident weights), the fused attention (through HuBERT), the tensor-core kNN short list, the realtime tail."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "retrieval-based-voice-conversion-webui_b200"))
from rvc_b200 import _lib, engine, synthetic as SY  # noqa: E402

_lib.init(0)
L = _lib.lib()
g = torch.Generator().manual_seed(0)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for Cc, k, T in ((32, 3, 1500), (64, 7, 900)):
    x = torch.randn(T, Cc, generator=g).cuda()
    ws = [(torch.randn(Cc, Cc, k, generator=g) / np.sqrt(Cc * k)).contiguous() for _ in range(6)]
    bs = [(torch.randn(Cc, generator=g) * 0.1).contiguous() for _ in range(6)]
    dil = (C.c_int * 3)(1, 3, 5)
    arr = lambda ts: (C.c_void_p * 3)(*[t.data_ptr() for t in ts])
    y = torch.empty(T, Cc, device="cuda")
    _lib.check(L.rvcb_op_resblock1(Cc, k, dil, arr(ws[0:3]), arr(bs[0:3]), arr(ws[3:6]), arr(bs[3:6]), C.c_void_p(x.data_ptr()), T, C.c_void_p(y.data_ptr()), st))
    assert torch.isfinite(y).all()
hub = engine.Hubert(SY.hubert_weights(777, n_layers=2))
feats = hub.extract(SY.synth_voice(0.9, seed=1).cuda(), 2)            # fused attention + streaming / split-K GEMMs
db = torch.randn(3000, 768, generator=g).cuda()
q = db[:130] + 0.01
D0, I0 = engine.knn_bruteforce_top1(db, q)
D1, I1 = engine.FlatIndex(db).search(q)
assert torch.equal(I0, I1) and torch.equal(D0, D1)
buf = torch.zeros(1920, device="cuda")
out = engine.rt_tail(torch.randn(10080, generator=g).cuda(), torch.randn(10600, generator=g).cuda(), 480, 0.0, buf, 7680, 480)
torch.cuda.synchronize()
print("ok", feats.shape, float(out.abs().max()))

"""Small invocations of the round-2 kernels for compute-sanitizer (memcheck / racecheck / synccheck): the fused residual block
(streamed and resident weights), the fused attention (through HuBERT), the tensor-core kNN short list, the realtime tail, and
(SANITIZE_MORE=1) the halo convolutions + split-precision DFT through a short RMVPE run, the spectral gate, the resampler and the
keep-mode synthesizer."""
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
if os.environ.get("SANITIZE_MORE"):
    rm = engine.Rmvpe(SY.rmvpe_weights(4321))
    f0 = rm.infer(SY.synth_voice(0.35, seed=3).cuda(), 0.03)[0]              # conv2d_row / conv2d_tile / split DFT / GRU
    from infer.modules.gui import Resample, TorchGate
    xn = torch.randn(1, 9600, generator=g).cuda() * 0.1
    yg = TorchGate(sr=48000, n_fft=1920, prop_decrease=0.9).to("cuda:0")(xn[:, -3840:].contiguous(), xn)
    yr = Resample(48000, 16000).to("cuda:0")(xn[0])
    cfg = SY.V2_48K_CONFIG if hasattr(SY, "V2_48K_CONFIG") else None
    if cfg is not None:
        syn = engine.Synth(SY.synth_weights(1234), cfg, 768)
        T = 150
        kept = syn.infer_keep((torch.randn(T, 768, generator=g) * 0.5).cuda(), 0, torch.randint(1, 255, (T,), generator=g).cuda(),
                              (torch.rand(T, generator=g) * 300 + 80).cuda(), torch.randn(192, T, generator=g).cuda(),
                              torch.randn(T * 480, generator=g).cuda(), 40, 60)
        assert torch.isfinite(kept).all()
    assert torch.isfinite(f0).all() and torch.isfinite(yg).all() and torch.isfinite(yr).all()
torch.cuda.synchronize()
print("ok", feats.shape, float(out.abs().max()))

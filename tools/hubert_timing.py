"""CUDA-event timing of HuBERT feature extraction and RMVPE, each alone on the GPU (one 16 s padded utterance), warm."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "retrieval-based-voice-conversion-webui_b200"))
from rvc_b200 import engine, synthetic as SY  # noqa: E402

hub = engine.Hubert(SY.hubert_weights(777))
rmv = engine.Rmvpe(SY.rmvpe_weights(4321))
x = torch.randn(256000, device="cuda") * 0.1
for name, fn in (("hubert.extract", lambda: hub.extract(x, 12)), ("rmvpe.infer", lambda: rmv.infer(x, 0.03))):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        fn()
    e.record()
    torch.cuda.synchronize()
    print(f"{name} {s.elapsed_time(e) / 20:.3f} ms")

"""RVCB_GRU_TRACE=1: one RMVPE run on 16 s of audio; the BiGRU kernel prints where a recurrence step's cycles go."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "retrieval-based-voice-conversion-webui_b200"))
from rvc_b200 import engine, synthetic as SY  # noqa: E402
rm = engine.Rmvpe(SY.rmvpe_weights(4321))
wav = SY.synth_voice(16.0, seed=0).cuda()
for _ in range(2):
    rm.infer(wav, 0.03)
torch.cuda.synchronize()
if not os.environ.get("RVCB_GRU_TRACE"):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        rm.infer(wav, 0.03)
    b.record()
    torch.cuda.synchronize()
    print(f"RMVPE eager on 16 s: {a.elapsed_time(b) / 10:.3f} ms per call (RVCB_GRU_FAST={os.environ.get('RVCB_GRU_FAST', '')})")

"""One device-resident utterance (same dev_step as bench.py) bracketed by cudaProfilerStart/Stop, for
`ncu --profile-from-start off ...`; with RVCB_PROF_CSV set it also dumps the per-launch GEMM table."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "retrieval-based-voice-conversion-webui_b200"))
from rvc_b200 import _lib, engine, synthetic as SY  # noqa: E402
from rvc_b200.index_build import build_ivf_layout  # noqa: E402

dev = torch.device("cuda", 0)
_lib.init(0)
hub = engine.Hubert(SY.hubert_weights(777))
rmv = engine.Rmvpe(SY.rmvpe_weights(4321))
net = engine.Synth(SY.synth_weights(1234), SY.V2_48K_CONFIG, 768)
n_index = int(os.environ.get("N_INDEX", "100000"))
index = engine.Index.from_oracle_layout(build_ivf_layout(SY.index_vectors(n_index, 768, 0).numpy(), None, seed=0, device="cuda"))
audio = SY.synth_voice(10.0, seed=0)
audio_pad = torch.from_numpy(np.pad(audio.numpy(), (48000, 48000), mode="reflect")).to(dev)
T2 = 2 * hub.num_frames(audio_pad.shape[0])
pitchf = (200 + 50 * torch.sin(torch.arange(T2) / 40.0)).to(dev)
pitch = torch.full((T2,), 80, dtype=torch.long, device=dev)
n1 = torch.randn(192, T2, device=dev)
n2 = torch.randn(T2 * 480, device=dev)


side = torch.cuda.Stream(device=dev)
USE_SIDE = os.environ.get("SIDE_STREAM", "1") == "1"


def dev_step():
    cur = torch.cuda.current_stream()
    if USE_SIDE:
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            f0, _, _ = rmv.infer(audio_pad, 0.03)
    else:
        f0, _, _ = rmv.infer(audio_pad, 0.03)
    feats = hub.extract(audio_pad, 12)
    D, I = index.search_device(feats, 8)
    fb = index.blend_device(feats, D, I, 0.75)
    phone = engine.upsample_protect(fb, feats, pitchf, T2, 0.33)
    out = net.infer(phone, 0, pitch, pitchf, n1, n2)
    if USE_SIDE:
        cur.wait_stream(side)
    return out


for _ in range(3):
    dev_step()
torch.cuda.synchronize()
if os.environ.get("RVCB_PROF_CSV"):
    _lib.check(_lib.lib().rvcb_prof_begin())
    dev_step()
    ms, n = C.c_double(0), C.c_ulonglong(0)
    _lib.check(_lib.lib().rvcb_prof_end(C.byref(ms), C.byref(n)))
    print("gemm launches", n.value, "gemm ms", ms.value)
torch.cuda.profiler.start()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
dev_step()
e.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("step ms", s.elapsed_time(e))

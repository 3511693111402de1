"""CUDA-event timing of the device front-end / back-end kernels of one 10 s utterance (the pieces the reference runs on the host)."""
import os
import sys

import numpy as np
import torch
from scipy import signal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "retrieval-based-voice-conversion-webui_b200"))
from rvc_b200 import engine  # noqa: E402

bh, ah = signal.butter(N=5, Wn=48, btype="high", fs=16000)
sos, zi = engine.highpass_sos_from_ba(bh, ah)
rng = np.random.default_rng(0)
x = torch.from_numpy((rng.standard_normal(160000) * 0.3).astype(np.float32)).cuda()
f0 = torch.from_numpy((200 + 50 * np.sin(np.arange(1601) / 40.0)).astype(np.float32)).cuda()
wav = torch.randn(479040, device="cuda")


def timeit(name, fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    print(f"{name:28s} {s.elapsed_time(e) / n * 1e3:9.1f} us")


timeit("sosfiltfilt 160000", lambda: engine.sosfiltfilt(sos, zi, 18, x))
timeit("reflect_pad 160000+96000", lambda: engine.reflect_pad(x, 48000))
timeit("f0_post 1601->1600", lambda: engine.f0_post(f0, 1600, 0))
timeit("post_mix 479040", lambda: engine.post_mix(wav.clone(), 48000, x, 0.25))
timeit("f32_to_i16 479040", lambda: engine.f32_to_i16(wav))

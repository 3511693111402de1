"""Tiny end-to-end run of every handle for compute-sanitizer (memcheck / racecheck): short audio, small index."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "retrieval-based-voice-conversion-webui_b200"))
from rvc_b200 import engine, synthetic as SY  # noqa: E402
from rvc_b200.index_build import build_ivf_layout  # noqa: E402

hub = engine.Hubert(SY.hubert_weights(777))
rmv = engine.Rmvpe(SY.rmvpe_weights(4321))
net = engine.Synth(SY.synth_weights(1234), SY.V2_48K_CONFIG, 768)
ix = engine.Index.from_oracle_layout(build_ivf_layout(SY.index_vectors(2000, 768, 0).numpy(), 32, seed=0, device="cuda"))
wav = SY.synth_voice(0.9, seed=1).cuda()
f0, mel, hid = rmv.infer(wav, 0.03, True, True)
feats = hub.extract(wav, 12)
D, I = ix.search_device(feats, 8)
fb = ix.blend_device(feats, D, I, 0.75)
T = 2 * feats.shape[0]
pitchf = torch.full((T,), 200.0, device="cuda")
pitch = torch.full((T,), 70, dtype=torch.long, device="cuda")
phone = engine.upsample_protect(fb, feats, pitchf, T, 0.33)
out = net.infer(phone, 0, pitch, pitchf, torch.randn(192, T, device="cuda"), torch.randn(T * 480, device="cuda"))
out2 = net.infer(phone, 0, pitch, pitchf, torch.randn(192, T - 40, device="cuda"), torch.randn(12 * 480, device="cuda"), 64, 12, 13)
y = engine.post_mix(out.clone(), 48000, wav, 0.25)
Db, Ib = engine.knn_bruteforce_top1(torch.randn(3000, 768, device="cuda"), torch.randn(5, 768, device="cuda"))
torch.cuda.synchronize()
print("ok", out.shape, out2.shape, float(y.abs().max()), f0.shape)

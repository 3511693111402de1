"""Debug helper: where does keep-mode synthesis differ from the slice of the full decode?"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "retrieval-based-voice-conversion-webui_b200")); sys.path.insert(0, ROOT)
from oracle import weights as OW
from rvc_b200.engine import Synth
cfg = OW.V2_48K_CONFIG
for f0 in (1, 0):
    w = OW.synth_weights(1234) if f0 else {k: v for k, v in OW.synth_weights(1234).items() if "emb_pitch" not in k and "noise_convs" not in k and "m_source" not in k}
    syn = Synth(w, cfg, 768)
    g = torch.Generator().manual_seed(5)
    for T, kh, kl in ((420, 100, 220), (300, 40, 200), (260, 0, 260), (500, 300, 150), (1598, 284, 1030)):
        phone = (torch.randn(T, 768, generator=g) * 0.5).cuda()
        pitch = torch.randint(1, 255, (T,), generator=g).cuda() if f0 else None
        pitchf = (torch.rand(T, generator=g) * 300 + 80).cuda() if f0 else None
        if f0:
            pitchf[T // 3: T // 3 + 20] = 0
        n1 = torch.randn(192, T, generator=g).cuda()
        n2 = torch.randn(T * 480, generator=g).cuda() if f0 else None
        full = syn.infer(phone, 0, pitch, pitchf, n1, n2)
        kept = syn.infer_keep(phone, 0, pitch, pitchf, n1, n2, kh, kl)
        ref = full[kh * 480: (kh + kl) * 480]
        d = (kept - ref).abs()
        nz = torch.nonzero(d > 0).flatten()
        print(f"f0={f0} T={T} kh={kh} kl={kl} max={d.max().item():.3e} ndiff={nz.numel()}",
              (f"first={nz[0].item()} ({nz[0].item()/480:.2f} fr) last={nz[-1].item()} ({nz[-1].item()/480:.2f} fr)" if nz.numel() else ""), flush=True)

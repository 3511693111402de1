"""BASELINE config #3: realtime gui.py block (160 ms, 48 kHz, extra 2.5 s, crossfade 0.05 s -> 43520-sample 16 kHz window,
skip_head 250, return_length 21): p50 / p99 latency of rtrvc.RVC.infer per block over N blocks."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "retrieval-based-voice-conversion-webui_b200"))
from rvc_b200 import engine, synthetic as SY  # noqa: E402
from rvc_b200.index_build import build_ivf_layout  # noqa: E402
from infer.lib.rtrvc import RVC  # noqa: E402
from infer.modules.vc.utils import HubertB200  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
hub = HubertB200(SY.hubert_weights(777), "cuda:0")
ix = engine.Index.from_oracle_layout(build_ivf_layout(SY.index_vectors(100000, 768, 0).numpy(), None, seed=0, device="cuda"))
rt = RVC(0, 0, SY.synth_cpt(1234, "v2"), ix, 0.75, device="cuda:0", hubert_model=hub, rmvpe_state_dict=SY.rmvpe_weights(4321))
stream = SY.synth_voice(2.72 + 0.16 * (N + 20), seed=5).cuda()
lat = []
for b in range(N + 20):
    win = stream[b * 2560: b * 2560 + 43520]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    y = rt.infer(win, 2560, 250, 21, "rmvpe")
    y_host = y.cpu()                      # the GUI copies the block to the output ring (gui.py:1091-1126)
    lat.append((time.perf_counter() - t0) * 1e3)
lat = np.array(lat[20:])
print(f"realtime block (160 ms audio): n={len(lat)} p50={np.percentile(lat, 50):.2f} ms p90={np.percentile(lat, 90):.2f} ms "
      f"p99={np.percentile(lat, 99):.2f} ms max={lat.max():.2f} ms  -> {160.0 / np.percentile(lat, 50):.1f}x faster than the block period")

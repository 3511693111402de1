"""cProfile of the e2e front door (VC.vc_single, host numpy in -> host int16 out) on one 10 s utterance."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "retrieval-based-voice-conversion-webui_b200"))
from rvc_b200 import synthetic as SY  # noqa: E402
from rvc_b200.engine import Index  # noqa: E402
from rvc_b200.index_build import build_ivf_layout  # noqa: E402
from infer.modules.vc.modules import VC  # noqa: E402
from infer.modules.vc.utils import HubertB200  # noqa: E402


class Cfg:
    x_pad, x_query, x_center, x_max, is_half = 3, 10, 60, 65, True
    device = "cuda:0"
    rmvpe_state_dict = None


cfg = Cfg()
cfg.rmvpe_state_dict = SY.rmvpe_weights(4321)
vc = VC(cfg)
vc.hubert_model = HubertB200(SY.hubert_weights(777), "cuda:0")
vc.get_vc(SY.synth_cpt(1234, "v2"))
index = Index.from_oracle_layout(build_ivf_layout(SY.index_vectors(100000, 768, 0).numpy(), None, seed=0, device="cuda"))
audio = SY.synth_voice(10.0, seed=0).numpy()


def step():
    info, out = vc.vc_single(0, audio, 0, None, "rmvpe", index, "", 0.75, 3, 0, 0.25, 0.33)
    assert out is not None, info
    return info


for _ in range(5):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    info = step()
torch.cuda.synchronize()
print("e2e ms/step", (time.perf_counter() - t0) / 20 * 1e3, info.replace("\n", " | "))
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)

# GPU-side view of the same call: graph replay alone (events on the current stream), then the host-visible pieces
pipe = vc.pipeline
ent = next((e for e in pipe._graphs.values() if "graph" in e), None)
if ent is not None:
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(20):
        ent["graph"].replay()
    e.record()
    torch.cuda.synchronize()
    print("graph replay alone ms", s.elapsed_time(e) / 20)
    t0 = time.perf_counter()
    for _ in range(20):
        x = pipe._stage_h2d(audio.astype(np.float32))
    torch.cuda.synchronize()
    print("pinned stage + H2D ms", (time.perf_counter() - t0) / 20 * 1e3)
    t0 = time.perf_counter()
    for _ in range(20):
        y = ent["out"].cpu().numpy()
    print("D2H + numpy ms", (time.perf_counter() - t0) / 20 * 1e3)

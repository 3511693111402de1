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

from infer.modules.vc.modules import VC  # noqa: E402
from infer.modules.vc.utils import HubertB200  # noqa: E402

dev = torch.device("cuda", 0)
_lib.init(0)


class Cfg:
    x_pad, x_query, x_center, x_max, is_half = 3, 10, 60, 65, True
    device = "cuda:0"
    rmvpe_state_dict = None


cfg = Cfg()
cfg.rmvpe_state_dict = SY.rmvpe_weights(4321)
vc = VC(cfg)
vc.hubert_model = HubertB200(SY.hubert_weights(777), dev)
vc.get_vc(SY.synth_cpt(1234, "v2"))
n_index = int(os.environ.get("N_INDEX", "100000"))
index = engine.Index.from_oracle_layout(build_ivf_layout(SY.index_vectors(n_index, 768, 0).numpy(), None, seed=0, device="cuda"))
audio = SY.synth_voice(10.0, seed=0).numpy()
x_dev = torch.from_numpy(np.divide(audio, max(1.0, np.abs(audio).max() / 0.95)).astype(np.float32)).to(dev)
pipe = vc.pipeline
body_args = (vc.hubert_model, vc.net_g, torch.tensor(0).unsqueeze(0).long(), [0, 0, 0], 0, index, index.vectors, 0.75, 1, 48000, 0.25,
             "v2", 0.33, True)
USE_SIDE = os.environ.get("SIDE_STREAM", "1") == "1"
if not USE_SIDE:
    pipe._side = torch.cuda.current_stream()      # every launch on one stream: per-launch event timing sees each kernel alone


def dev_step():
    """The product path's own device-resident body (Pipeline._dev_body), utterance already in HBM."""
    return pipe._dev_body(x_dev, *body_args)


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

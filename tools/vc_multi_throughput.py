"""End-to-end batch conversion through the reference-facing call (VC.vc_multi: wav files in -> wav files out) for RVCB_LANES = 1, 2, 4:
16 x 10 s utterances, config #2 models (v2/48k, RMVPE, 100 k-vector IVF index), wall clock per file after one warm-up pass."""
import os
import sys
import tempfile
import time

import numpy as np
import torch
from scipy.io import wavfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "retrieval-based-voice-conversion-webui_b200"))
from rvc_b200 import _lib, engine, synthetic as SY  # noqa: E402
from rvc_b200.index_build import build_ivf_layout  # noqa: E402
from infer.modules.vc.modules import VC  # noqa: E402
from infer.modules.vc.utils import HubertB200  # noqa: E402

_lib.init(0)


class Cfg:
    x_pad, x_query, x_center, x_max, is_half = 3, 10, 60, 65, True
    device = "cuda:0"
    rmvpe_state_dict = None


cfg = Cfg()
cfg.rmvpe_state_dict = SY.rmvpe_weights(4321)
vc = VC(cfg)
vc.hubert_model = HubertB200(SY.hubert_weights(777), "cuda:0")
vc.get_vc(SY.synth_cpt(1234, "v2"))
index = engine.Index.from_oracle_layout(build_ivf_layout(SY.index_vectors(100000, 768, 0).numpy(), None, seed=0, device="cuda"))
tmp = tempfile.mkdtemp()
indir = os.path.join(tmp, "in")
os.makedirs(indir)
N = 16
for i in range(N):
    wavfile.write(os.path.join(indir, f"u{i:02d}.wav"), 16000, (SY.synth_voice(10.0, seed=200 + i).numpy() * 32767).astype(np.int16))
for lanes in (1, 2, 4):
    os.environ["RVCB_LANES"] = str(lanes)
    for rep in range(2):                       # first pass: lane creation, eager run, graph capture
        out = os.path.join(tmp, f"out{lanes}_{rep}")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        msgs = list(vc.vc_multi(0, indir, out, [], 0, "rmvpe", index, "", 0.75, 3, 0, 0.25, 0.33, "wav"))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    ok = msgs[-1].count("Success")
    print(f"RVCB_LANES={lanes}: {N} files of 10 s in {dt * 1e3:7.1f} ms = {dt / N * 1e3:6.2f} ms per file (wav decode + conversion + float32 wav "
          f"write), {N * 479040 / dt / 1e6:6.1f} M samples/s, {ok} succeeded", flush=True)

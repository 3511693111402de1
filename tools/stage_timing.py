"""Per-stage device time of one config-#2 utterance, each stage captured into its own CUDA graph and replayed (so the numbers are
in-graph times like the product's, without eager launch overhead): prologue DSP, RMVPE + f0 post, HuBERT + retrieval, the two front
branches together (the product's fork / join with the grid cap), synthesizer (keep mode), epilogue."""
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
from infer.modules.vc import pipeline as P  # noqa: E402

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
index = engine.Index.from_oracle_layout(build_ivf_layout(SY.index_vectors(100000, 768, 0).numpy(), None, seed=0, device="cuda"))
audio = SY.synth_voice(10.0, seed=0).numpy()
x_dev = torch.from_numpy(np.divide(audio, max(1.0, np.abs(audio).max() / 0.95)).astype(np.float32)).to(dev)
pipe = vc.pipeline
sid = torch.tensor(0).unsqueeze(0).long()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def graph_time(name, fn, reps=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        n0 = _lib.lib().rvcb_launch_count()
        with torch.cuda.graph(g, stream=s):
            out = fn()
        n1 = _lib.lib().rvcb_launch_count()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    print(f"{name:44s} {ts[len(ts) // 2] * 1e3:9.1f} us   ({n1 - n0} launches)", flush=True)
    return out


if os.environ.get("ONLY_RMVPE"):
    a16 = engine.sosfiltfilt(P.sos_h, P.sos_zi_h, 3 * max(len(P.ah), len(P.bh)), x_dev)
    audio_pad = engine.reflect_pad(a16, pipe.t_pad)
    graph_time("RMVPE + f0 post (alone, all SMs)", lambda: pipe.f0_gen.calculate_device(audio_pad, audio_pad.numel() // pipe.window, 0))
    graph_time("HuBERT only", lambda: vc.hubert_model.extract_features(source=audio_pad.view(1, -1), padding_mask=None, output_layer=12))
    sys.exit(0)
if not os.environ.get("ONLY_STEP"):
    a16 = graph_time("prologue: sosfiltfilt", lambda: engine.sosfiltfilt(P.sos_h, P.sos_zi_h, 3 * max(len(P.ah), len(P.bh)), x_dev))
    audio_pad = graph_time("prologue: reflect_pad", lambda: engine.reflect_pad(a16, pipe.t_pad))
    p_len = audio_pad.numel() // pipe.window
    pp = graph_time("RMVPE + f0 post (alone, all SMs)", lambda: pipe.f0_gen.calculate_device(audio_pad, p_len, 0))
    ff = graph_time("HuBERT + retrieval + blend (alone, all SMs)", lambda: pipe._features(vc.hubert_model, audio_pad, index, index.vectors, 0.75, "v2"))
    graph_time("HuBERT only", lambda: vc.hubert_model.extract_features(source=audio_pad.view(1, -1), padding_mask=None, output_layer=12))


    def both():
        cur = torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(cur)
        prev = engine.set_grid_cap(engine.front_branch_cap())
        try:
            r = pipe.f0_gen.calculate_device(audio_pad, p_len, 0)
            pipe._side.wait_event(fork)
            with torch.cuda.stream(pipe._side):
                f = pipe._features(vc.hubert_model, audio_pad, index, index.vectors, 0.75, "v2")
                ev = torch.cuda.Event()
                ev.record(pipe._side)
        finally:
            engine.set_grid_cap(prev)
        cur.wait_event(ev)
        return r, f


    graph_time("both front branches (fork / join, grid cap)", both)
    pitch, pitchf = pp[0].unsqueeze(0), pp[1].unsqueeze(0)


    def synth(trim):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        pipe._prefetched = (audio_pad, ff[0], ff[1], ev)
        return pipe._vc_dev(vc.hubert_model, vc.net_g, sid, audio_pad, pitch, pitchf, [0, 0, 0], index, index.vectors, 0.75, "v2", 0.33, trim=trim)


    try:
        out = graph_time("upsample/protect + synthesizer (keep mode)", lambda: synth(True))
        graph_time("upsample/protect + synthesizer (full decode)", lambda: synth(False))
        graph_time("epilogue: rms mix + scale + int16", lambda: engine.f32_to_i16(engine.post_mix(out.contiguous(), 48000, a16, 0.25)))
    except Exception as e:                                   # _prefetched contract differs: report and go on
        print("synth stage skipped:", repr(e)[:300])
graph_time("whole step (product _dev_body)", lambda: pipe._dev_body(x_dev, vc.hubert_model, vc.net_g, sid, [0, 0, 0], 0, index, index.vectors,
                                                                    0.75, 1, 48000, 0.25, "v2", 0.33, True))

"""GPU end-to-end: the drop-in Pipeline / VC / rtrvc.RVC front doors against the oracle pipeline on one utterance."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class Cfg:
    x_pad, x_query, x_center, x_max, is_half = 1, 6, 38, 41, False
    device = "cuda:0"
    rmvpe_state_dict = None


def _setup(seconds=2.0, n_index=3000):
    from oracle import ivf as OI, pipeline as OP, weights as OW
    hw, rw, sw = OW.hubert_weights(777), OW.rmvpe_weights(4321), OW.synth_weights(1234)
    audio = OW.synth_voice(seconds, seed=4).numpy()
    idx = OI.build_ivf(OW.index_vectors(n_index, 768, 1).numpy(), None, seed=0, exact_assign=True)
    return OI, OP, OW, hw, rw, sw, audio, idx


def test_pipeline_matches_oracle_with_shared_pitch_and_noise():
    from infer.modules.vc.pipeline import Pipeline
    from infer.modules.vc.utils import HubertB200
    from rvc.synthesizer import get_synthesizer
    from rvc_b200.engine import Index
    OI, OP, OW, hw, rw, sw, audio, idx = _setup()
    cfg = Cfg()
    cfg.rmvpe_state_dict = rw
    op = OP.OraclePipeline(48000, 1, 6, 38, 41, hw, rw, sw, OW.V2_48K_CONFIG, noise_seed=3)
    with torch.no_grad():
        ref = op.pipeline(0, audio.copy(), 0, "rmvpe", idx, 0.75, 1, 48000, 0, 0.25, "v2", 0.33)
    pitch, pitchf = op.pitch[0].numpy(), op.pitchf[0].numpy()
    pipe = Pipeline(48000, cfg)
    hub = HubertB200(hw, "cuda:0")
    net_g, cpt = get_synthesizer(OW.synth_cpt(1234, "v2"), "cuda:0")
    gidx = Index.from_oracle_layout(idx)
    # 1) same f0 track + same noise -> isolates HuBERT + retrieval + synthesizer
    net_g.set_noise(*op.taps[0]["noise"])
    times = [0, 0, 0]
    out = pipe.pipeline(hub, net_g, 0, audio.copy(), times, 0, (pitch, pitchf.astype(np.float64)), gidx, 0.75, 2, 3, 48000, 0, 0.25, "v2", 0.33)
    assert out.shape == ref.shape
    # retrieval indices: the nearest neighbour (the argmax index) must be the oracle's on every frame; fp16-operand feature
    # noise (3e-4 mean) may flip a few near-ties among ranks 2..8
    feats = hub.extract_features(source=torch.from_numpy(np.pad(__import__("scipy.signal").signal.filtfilt(
        __import__("infer.modules.vc.pipeline", fromlist=["bh"]).bh, __import__("infer.modules.vc.pipeline", fromlist=["ah"]).ah, audio),
        (16000, 16000), mode="reflect").astype(np.float32))[None].cuda(), output_layer=12)[0][0]
    _, I = gidx.search_device(feats, 8)
    I = I.cpu().numpy()
    assert np.array_equal(I[:, 0], op.taps[0]["ix"][:, 0])
    same = (I == op.taps[0]["ix"]).mean()
    assert same >= 0.98, same
    err = np.abs(out - ref).max() / 32768.0
    print(f"[parity] 2 s utterance: all-8 neighbours same {same:.4f}, e2e max abs err {err:.3e}")
    assert err < 1.5e-3, f"end-to-end max abs err (full scale) {err}"      # measured 7.4e-4
    # 2) the full path with its own RMVPE f0: coarse pitch agrees on >= 95 % of frames, f0 within 1 % where both voiced
    c2, f2 = pipe.f0_gen.calculate(np.pad(__import__("scipy.signal").signal.filtfilt(
        __import__("infer.modules.vc.pipeline", fromlist=["bh"]).bh, __import__("infer.modules.vc.pipeline", fromlist=["ah"]).ah, audio),
        (16000, 16000), mode="reflect").astype(np.float32), len(pitch), 0, "rmvpe", 3)
    assert (c2[: len(pitch)] == pitch).mean() >= 0.95
    both = (f2[: len(pitchf)] > 0) & (pitchf > 0)
    assert np.median(np.abs(f2[: len(pitchf)][both] / pitchf[both] - 1)) < 1e-2


def test_resample_sr_branch_on_the_device():
    """pipeline.py:351-354 (resample_sr != tgt_sr): change_rms, resampler and peak scaling stay on the device (rvcb_rms_mix ->
    rvcb_resample_sinc -> rvcb_post_mix(rate 1)) in both the single-chunk graph path and the host-chunking path; compared with the
    oracle pipeline, whose stand-in for librosa's soxr resampler is torchaudio's sinc resampler (parity unpinned for this branch)."""
    from infer.modules.vc.pipeline import Pipeline
    from infer.modules.vc.utils import HubertB200
    from rvc.synthesizer import get_synthesizer
    from rvc_b200.engine import Index
    OI, OP, OW, hw, rw, sw, audio, idx = _setup()
    cfg = Cfg()
    cfg.rmvpe_state_dict = rw
    op = OP.OraclePipeline(48000, 1, 6, 38, 41, hw, rw, sw, OW.V2_48K_CONFIG, noise_seed=3)
    with torch.no_grad():
        ref = op.pipeline(0, audio.copy(), 0, "rmvpe", idx, 0.75, 1, 48000, 44100, 0.25, "v2", 0.33)
    pitch, pitchf = op.pitch[0].numpy(), op.pitchf[0].numpy()
    pipe = Pipeline(48000, cfg)
    hub = HubertB200(hw, "cuda:0")
    net_g, cpt = get_synthesizer(OW.synth_cpt(1234, "v2"), "cuda:0")
    gidx = Index.from_oracle_layout(idx)
    assert ref.shape[0] == -(-95040 * 147 // 160)               # ceil(n_48k * 44100 / 48000), n_48k = (2 * 199 - 200) * 480
    # 1) the reference's chunking control flow (host DSP prologue, shared f0 + noise) with the device epilogue, vs the oracle
    pipe._force_host = True
    net_g.set_noise(*op.taps[0]["noise"])
    out = pipe.pipeline(hub, net_g, 0, audio.copy(), [0, 0, 0], 0, (pitch, pitchf.astype(np.float64)), gidx, 0.75, 2, 3, 48000, 44100, 0.25,
                        "v2", 0.33)
    assert out.shape == ref.shape, (out.shape, ref.shape)
    err = np.abs(out - ref).max() / 32768.0
    print(f"[parity] resample_sr 44100 (shared f0 + noise): e2e max abs err {err:.3e}")
    assert err < 1.5e-3, err
    # 2) the device-resident single-chunk path (own RMVPE, eager then captured graph): same utterance without resampling, resampled
    #    on the host by torchaudio and peak-scaled like pipeline.py:356-360, must equal the device branch
    import torchaudio
    pipe._force_host = False
    outs = {}
    for rs in (0, 44100, 44100, 44100):                       # the 3rd / 4th call with the same key capture and replay the graph
        net_g._noise.clear()
        net_g.set_noise(*op.taps[0]["noise"])
        outs[rs] = pipe.pipeline(hub, net_g, 0, audio.copy(), [0, 0, 0], 0, "rmvpe", gidx, 0.75, 1, 3, 48000, rs, 1.0, "v2", 0.33)
    net_g._noise.clear()
    y = outs[0] / 32768.0
    assert np.abs(outs[0]).max() < 0.98 * 32768                # unscaled regime: outs[0] = y * 32768 exactly
    exp = torchaudio.functional.resample(torch.from_numpy(y.astype(np.float32)), 48000, 44100).numpy()
    exp = exp * (32768 / max(1.0, np.abs(exp).max() / 0.99))
    assert outs[44100].shape == exp.shape == ref.shape
    err = np.abs(outs[44100] - exp).max() / 32768.0
    print(f"[parity] resample_sr 44100 (device-resident, own RMVPE, graph replay) vs host torchaudio on the 48 kHz result: {err:.3e}")
    assert err < 2e-5, err


def test_vc_facade_and_realtime_engine_run():
    from infer.lib.rtrvc import RVC
    from infer.modules.vc.modules import VC
    from infer.modules.vc.utils import HubertB200
    from rvc_b200.engine import Index
    OI, OP, OW, hw, rw, sw, audio, idx = _setup(1.5, 1200)
    cfg = Cfg()
    cfg.rmvpe_state_dict = rw
    vc = VC(cfg)
    vc.hubert_model = HubertB200(hw, "cuda:0")
    upd = vc.get_vc(OW.synth_cpt(1234, "v2"))
    assert upd["maximum"] == 109
    gidx = Index.from_oracle_layout(idx)
    info, (sr, wav) = vc.vc_single(0, audio, 0, None, "rmvpe", gidx, "", 0.75, 3, 0, 0.25, 0.33)
    assert info.startswith("Success") and sr == 48000 and wav.dtype == np.int16 and wav.shape[0] == 71040   # (2*174 frames)*480 - 2*x_pad*48000
    assert np.isfinite(wav.astype(np.float32)).all() and np.abs(wav).max() > 100
    # error convention: exceptions become the info string (modules.py:196-199)
    info, out = vc.vc_single(0, audio, 0, None, "harvest", gidx, "", 0.75, 3, 0, 0.25, 0.33)
    assert out is None and "harvest" in info
    # realtime block (gui.py sizes): 43520-sample window, skip_head 250, return_length 21
    rt = RVC(0, 0, OW.synth_cpt(1234, "v2"), gidx, 0.5, device="cuda:0", hubert_model=vc.hubert_model, rmvpe_state_dict=rw)
    assert rt.tgt_sr == 48000 and rt.if_f0 == 1 and rt.version == "v2"
    win = torch.from_numpy(OW.synth_voice(2.72, seed=9).numpy()[:43520]).cuda()
    y = rt.infer(win, 2560, 250, 21, "rmvpe")
    assert y.shape == (21 * 480,) and torch.isfinite(y).all()
    y2 = rt.infer(win, 2560, 250, 21, "rmvpe", protect=0.33)
    assert y2.shape == (10080,)
    # same shapes and settings again: capture into a CUDA graph, then replay; the pitch cache keeps rolling inside the graph
    cp0 = rt.cache_pitchf.clone()
    ys = [rt.infer(torch.roll(win, -2560 * i), 2560, 250, 21, "rmvpe") for i in range(1, 4)]
    assert any("graph" in e for e in rt._graphs.values())
    assert all(y.shape == (10080,) and torch.isfinite(y).all() and y.abs().max() > 1e-3 for y in ys)
    assert not torch.equal(ys[1], ys[2]) and not torch.equal(cp0, rt.cache_pitchf)


@pytest.mark.parametrize("rate", [0.25, 1.0])
def test_post_mix_matches_reference_host_dsp(rate):
    """Device RMS-envelope mix + peak normalisation vs the oracle's restatement of change_rms / scaling (pipeline.py:26-45, 356-360)."""
    from oracle import pipeline as OP, weights as OW
    from rvc_b200 import engine
    a16 = OW.synth_voice(3.0, seed=8).numpy()
    rng = np.random.RandomState(0)
    env = np.repeat(rng.rand(9) * 1.5 + 0.05, 16000)[: 3 * 48000]
    wav = (rng.randn(3 * 48000) * env).astype(np.float32)
    ref = wav.copy()
    if rate != 1:
        ref = OP.change_rms(a16, 16000, ref, 48000, rate)
    amax = np.abs(ref).max() / 0.99
    ref = ref * (32768 / amax if amax > 1 else 32768)
    out = engine.post_mix(torch.from_numpy(wav).cuda(), 48000, torch.from_numpy(a16).cuda(), rate).cpu().numpy()
    assert np.abs(out - ref).max() <= 2e-3 * np.abs(ref).max()


def test_pipeline_multichunk_long_audio_matches_oracle():
    """Audio longer than x_max is cut at the quietest samples (pipeline.py:222-236) and converted chunk by chunk."""
    from infer.modules.vc.pipeline import Pipeline
    from infer.modules.vc.utils import HubertB200
    from rvc.synthesizer import get_synthesizer
    OI, OP, OW, hw, rw, sw, audio, idx = _setup(7.0, 1500)

    class C2(Cfg):
        x_pad, x_query, x_center, x_max = 1, 1, 2, 3
    cfg = C2()
    cfg.rmvpe_state_dict = rw
    op = OP.OraclePipeline(48000, 1, 1, 2, 3, hw, rw, sw, OW.V2_48K_CONFIG, noise_seed=5)
    p_len = (len(audio) + 32000) // 160
    pitchf = (150 + 60 * np.sin(np.arange(p_len) / 30.0)).astype(np.float64)
    pitchf[200:260] = 0
    from oracle import rmvpe as ORM
    pitch, _ = ORM.post_process(pitchf.copy(), 0)
    with torch.no_grad():
        ref = op.pipeline(0, audio.copy(), 0, (pitch, pitchf), None, 0.0, 2, 48000, 0, 1.0, "v2", 0.5)
    assert len(op.taps) == 4                      # 3 cut points -> 4 chunks
    pipe = Pipeline(48000, cfg)
    net_g, _ = get_synthesizer(OW.synth_cpt(1234, "v2"), "cuda:0")
    for tp in op.taps:
        net_g.set_noise(*tp["noise"])
    out = pipe.pipeline(HubertB200(hw, "cuda:0"), net_g, 0, audio.copy(), [0, 0, 0], 0, (pitch, pitchf), "", 0.0, 2, 3, 48000, 0, 1.0, "v2", 0.5)
    assert out.shape == ref.shape
    err = np.abs(out - ref).max() / 32768.0
    print(f"[parity] multi-chunk 7 s utterance: e2e max abs err {err:.3e}")
    assert err < 1.5e-3, err      # measured 4.9e-4


def test_vc_single_with_index_file_and_vc_multi(tmp_path):
    """file_index as an on-disk faiss ``IwFl`` file (read without faiss), wav files in, wav files out (modules.py:201-266)."""
    from scipy.io import wavfile
    from infer.modules.vc.modules import VC
    from infer.modules.vc.utils import HubertB200
    from rvc_b200 import faiss_io
    OI, OP, OW, hw, rw, sw, audio, idx = _setup(1.2, 800)
    cfg = Cfg()
    cfg.rmvpe_state_dict = rw
    vc = VC(cfg)
    vc.hubert_model = HubertB200(hw, "cuda:0")
    vc.get_vc(OW.synth_cpt(1234, "v2"))
    ipath = str(tmp_path / "added_IVF20_Flat_nprobe_1_x_v2.index")
    faiss_io.write_index(ipath, idx)
    indir, outdir = tmp_path / "in", tmp_path / "out"
    indir.mkdir()
    for i in range(2):
        wavfile.write(str(indir / f"u{i}.wav"), 16000, (OW.synth_voice(1.2, seed=20 + i).numpy() * 32767).astype(np.int16))
    info, out = vc.vc_single(0, str(indir / "u0.wav"), 0, None, "rmvpe", ipath, "", 0.75, 3, 0, 0.25, 0.33)
    assert info.startswith("Success") and "Index: " + ipath in info and out[1].dtype == np.int16
    msgs = list(vc.vc_multi(0, str(indir), str(outdir), [], 0, "rmvpe", ipath, "", 0.75, 3, 0, 0.25, 0.33, "wav"))
    assert msgs and "Success" in msgs[-1]
    outs = sorted(os.listdir(outdir))
    assert outs == ["u0.wav.wav", "u1.wav.wav"]
    sr, y = wavfile.read(str(outdir / outs[0]))
    assert sr == 48000 and len(y) > 40000


def test_vc_multi_lanes_equal_the_serial_loop(tmp_path, monkeypatch):
    """SURVEY 8f-3 (batched front door) as built: ``vc_multi`` keeps RVCB_LANES utterances in flight on one GPU, each on its own
    lane (thread, streams, handles, captured graph).  With the noise draws pinned, the files written by two lanes equal the files
    of the serial loop sample for sample, and the log lines keep the input order."""
    from scipy.io import wavfile
    from infer.modules.vc.modules import VC
    from infer.modules.vc.utils import HubertB200
    from rvc_b200.engine import Index
    OI, OP, OW, hw, rw, sw, audio, idx = _setup(1.0, 800)
    cfg = Cfg()
    cfg.rmvpe_state_dict = rw
    vc = VC(cfg)
    vc.hubert_model = HubertB200(hw, "cuda:0")
    vc.get_vc(OW.synth_cpt(1234, "v2"))
    gidx = Index.from_oracle_layout(idx)
    indir = tmp_path / "in"
    indir.mkdir()
    names = [f"u{i}.wav" for i in range(6)]
    for i, nm in enumerate(names):
        wavfile.write(str(indir / nm), 16000, (OW.synth_voice(1.0, seed=40 + i).numpy() * 32767).astype(np.int16))
    monkeypatch.setattr(VC, "_inputs", staticmethod(lambda d, u: [os.path.join(d, n) for n in names]))     # listdir order is arbitrary
    n_pad = 16000 + 2 * 16000 * cfg.x_pad
    T = min(2 * ((n_pad - 400) // 320 + 1), n_pad // 160)          # synthesizer frames: min(2 x HuBERT frames, p_len), pipeline.py:142-146
    g = torch.Generator().manual_seed(3)
    n1 = torch.randn(1, 192, T, generator=g).cuda()
    n2 = torch.randn(1, T * 480, 1, generator=g).cuda()
    results = {}
    for lanes in (1, 2):
        monkeypatch.setenv("RVCB_LANES", str(lanes))
        outdir = tmp_path / f"out{lanes}"
        containers = [vc.net_g] + ([vc._lane(1).net_g] if lanes > 1 else [])
        for c in containers:
            c._noise.clear()
            for _ in range(len(names)):
                c.set_noise(n1, n2)
        msgs = list(vc.vc_multi(0, str(indir), str(outdir), [], 0, "rmvpe", gidx, "", 0.75, 3, 0, 0.25, 0.33, "wav"))
        lines = [ln for ln in msgs[-1].split("\n") if "->" in ln]          # one "name->Success." line per file (the info text has more lines)
        assert [ln.split("->")[0] for ln in lines] == names and all("Success" in ln for ln in lines), msgs[-1]
        results[lanes] = [wavfile.read(str(outdir / (nm + ".wav")))[1] for nm in names]
    for c in [vc.net_g, vc._lane(1).net_g]:
        c._noise.clear()
    for a, b in zip(results[1], results[2]):
        assert a.shape == b.shape and np.array_equal(a, b)
    assert not np.array_equal(results[1][0], results[1][1])


def test_long_single_chunk_utterance_runs_twice():
    """A 45 s utterance at the GPU config's x_pad = 3 stays one chunk (t_max = 65 s): 51 s of model compute, 2549 HuBERT frames,
    T = 5098 synthesizer frames -- larger than any shape the parity tests use (arena growth, M-keyed dispatch, attention over 5 k
    frames).  No oracle at this size (the CPU path needs minutes): the output must be finite, non-silent, of the reference's length,
    and the second call (captured graph) must reproduce the first with the noise pinned."""
    from infer.modules.vc.modules import VC
    from infer.modules.vc.utils import HubertB200
    from rvc_b200.engine import Index
    OI, OP, OW, hw, rw, sw, _, idx = _setup(1.0, 2000)

    class CfgH:
        x_pad, x_query, x_center, x_max, is_half = 3, 10, 60, 65, True
        device = "cuda:0"
        rmvpe_state_dict = rw

    vc = VC(CfgH())
    vc.hubert_model = HubertB200(hw, "cuda:0")
    vc.get_vc(OW.synth_cpt(1234, "v2"))
    gidx = Index.from_oracle_layout(idx)
    audio = OW.synth_voice(45.0, seed=8).numpy()
    n_pad = 45 * 16000 + 2 * 48000
    T = min(2 * ((n_pad - 400) // 320 + 1), n_pad // 160)
    g = torch.Generator().manual_seed(4)
    n1, n2 = torch.randn(1, 192, T, generator=g).cuda(), torch.randn(1, T * 480, 1, generator=g).cuda()
    outs = []
    for _ in range(3):
        vc.net_g._noise.clear()
        vc.net_g.set_noise(n1, n2)
        info, (sr, wav) = vc.vc_single(0, audio.copy(), 0, None, "rmvpe", gidx, "", 0.75, 3, 0, 0.25, 0.33)
        assert info.startswith("Success"), info
        outs.append(wav)
    vc.net_g._noise.clear()
    assert sr == 48000 and outs[0].dtype == np.int16 and outs[0].shape[0] == (T - 600) * 480
    assert np.abs(outs[0]).max() > 1000 and np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])


def test_no_f0_model_through_the_facade_and_realtime_engine():
    """cpt["f0"] == 0 end to end: get_vc picks the no-f0 container (modules.py:87-99 class table), the pipeline skips
    RMVPE (pipeline.py:203) and passes pitch=None, rtrvc skips its pitch cache (rtrvc.py:if_f0)."""
    from infer.lib.rtrvc import RVC
    from infer.modules.vc.modules import VC
    from infer.modules.vc.utils import HubertB200
    from rvc_b200.engine import Index
    OI, OP, OW, hw, rw, sw, audio, idx = _setup(1.5, 1200)
    cfg = Cfg()
    cfg.rmvpe_state_dict = rw
    vc = VC(cfg)
    vc.hubert_model = HubertB200(hw, "cuda:0")
    cpt = OW.synth_cpt(11, "v2", f0=0)
    vc.get_vc(cpt)
    assert vc.if_f0 == 0
    gidx = Index.from_oracle_layout(idx)
    info, (sr, wav) = vc.vc_single(0, audio, 0, None, "rmvpe", gidx, "", 0.75, 3, 0, 0.25, 0.33)
    assert info.startswith("Success"), info
    assert sr == 48000 and wav.dtype == np.int16 and wav.shape[0] == 71040 and np.abs(wav).max() > 100
    rt = RVC(0, 0, cpt, gidx, 0.5, device="cuda:0", hubert_model=vc.hubert_model, rmvpe_state_dict=rw)
    assert rt.if_f0 == 0
    win = torch.from_numpy(OW.synth_voice(2.72, seed=9).numpy()[:43520]).cuda()
    y = rt.infer(win, 2560, 250, 21, "rmvpe")
    assert y.shape == (21 * 480,) and torch.isfinite(y).all()


def test_device_front_end_kernels_match_host_dsp():
    """filtfilt (float64, block-parallel IIR), reflect padding, f0 post-processing and the int16 cast on the device vs
    scipy / numpy / the oracle's f0 post-processing."""
    from scipy import signal
    from rvc_b200 import engine
    from oracle import rmvpe as ORM
    bh, ah = signal.butter(N=5, Wn=48, btype="high", fs=16000)
    sos, zi = engine.highpass_sos_from_ba(bh, ah)
    rng = np.random.default_rng(5)
    for n in (19, 4000, 160000, 700001):
        x = (rng.standard_normal(n) * 0.3 + 0.05).astype(np.float32)
        if n == 4000:
            x += (0.4 * np.sin(2 * np.pi * 30 * np.arange(n) / 16000)).astype(np.float32)       # energy below the 48 Hz corner
        ref = signal.filtfilt(bh, ah, x)                      # the reference's direct-form filter (pipeline.py:221)
        got = engine.sosfiltfilt(sos, zi, 18, torch.from_numpy(x).cuda()).cpu().numpy()
        assert np.abs(got - ref).max() <= 2e-7, (n, np.abs(got - ref).max())     # float32 output: half an ulp at |y| ~ 1 is 6e-8
    x = torch.from_numpy(rng.standard_normal(5000).astype(np.float32))
    for pad in (0, 48, 4999, 5000, 12345):        # pad >= n: a 2 s utterance under the default 3 s pad (periodic reflection)
        assert np.array_equal(engine.reflect_pad(x.cuda(), pad).cpu().numpy(), np.pad(x.numpy(), (pad, pad), mode="reflect")), pad
    w = (rng.standard_normal(100000) * 20000).astype(np.float32)
    assert np.array_equal(engine.f32_to_i16(torch.from_numpy(w).cuda()).cpu().numpy(), w.astype(np.int16))
    # f0 contours with leading / interior / trailing unvoiced runs, isolated frames, all-unvoiced, resize up and down
    cases = []
    f0 = (200 + 50 * np.sin(np.arange(1601) / 40.0)).astype(np.float32)
    f0[:30] = 0; f0[100:160] = 0; f0[700:701] = 0; f0[1500:] = 0; f0[rng.random(1601) < 0.05] = 0
    cases += [(f0, 1600, 0), (f0, 1601, 5), (f0, 2000, -7), (f0[:300], 157, 12), (np.zeros(50, np.float32), 49, 3)]
    g = (80 + 900 * rng.random(400)).astype(np.float32); g[rng.random(400) < 0.5] = 0
    cases += [(g, 399, 0), (g, 400, 24), (g[-1:], 1, 0)]
    for _ in range(150):       # short random contours: every run / edge pattern of the gap fill (incl. the overwritten last frame)
        n = int(rng.integers(1, 40))
        h = (50 + 500 * rng.random(n)).astype(np.float32); h[rng.random(n) < rng.random()] = 0
        cases.append((h, n if rng.random() < 0.7 else int(rng.integers(1, 60)), int(rng.integers(-12, 13))))
    for f0, p_len, key in cases:
        want_c, want_f = ORM.post_process(ORM.interpolate_f0(ORM.resize_f0(f0.astype(np.float64), p_len)), key)
        pitch, pitchf = engine.f0_post(torch.from_numpy(f0).cuda(), p_len, key)
        assert np.array_equal(pitch.cpu().numpy(), want_c.astype(np.int64)), (p_len, key)
        assert np.array_equal(pitchf.cpu().numpy(), want_f.astype(np.float32)), (p_len, key)


def test_device_resident_utterance_path_equals_host_path():
    """vc_single through the device-resident single-chunk path (no host round trip) vs the host-DSP path of the same
    drop-in, same torch seed -> same noise draws: identical int16 audio up to 2 LSB."""
    from infer.modules.vc.modules import VC
    from infer.modules.vc.utils import HubertB200
    from rvc_b200.engine import Index
    OI, OP, OW, hw, rw, sw, audio, idx = _setup(1.5, 1200)
    cfg = Cfg()
    cfg.rmvpe_state_dict = rw
    vc = VC(cfg)
    vc.hubert_model = HubertB200(hw, "cuda:0")
    vc.get_vc(OW.synth_cpt(1234, "v2"))
    gidx = Index.from_oracle_layout(idx)
    outs = []
    for force_host in (False, True):
        vc.pipeline._force_host = force_host
        torch.manual_seed(11); torch.cuda.manual_seed(11)
        info, (sr, wav) = vc.vc_single(0, audio, 2, None, "rmvpe", gidx, "", 0.75, 3, 0, 0.25, 0.33)
        assert info.startswith("Success"), info
        outs.append(wav.astype(np.int32))
    assert outs[0].shape == outs[1].shape
    # not bit for bit: the device filter is a second-order-section cascade (~4e-8 from scipy's direct form), and one float32 ulp
    # at the input moves fp16 operand roundings downstream and the global peak-normalisation factor
    a, b = outs[0].astype(np.float64), outs[1].astype(np.float64)
    assert np.sqrt(np.mean((a - b) ** 2)) <= 0.01 * np.sqrt(np.mean(b ** 2)), np.sqrt(np.mean((a - b) ** 2)) / np.sqrt(np.mean(b ** 2))
    # third and fourth call with the same settings: CUDA-graph capture, then replay; fresh noise every call, same envelope
    vc.pipeline._force_host = False
    reps = []
    for _ in range(3):
        info, (sr, wav) = vc.vc_single(0, audio, 2, None, "rmvpe", gidx, "", 0.75, 3, 0, 0.25, 0.33)
        assert info.startswith("Success"), info
        reps.append(wav.astype(np.float64))
    assert any("graph" in e for e in vc.pipeline._graphs.values())
    assert not np.array_equal(reps[1], reps[2])                      # the captured randn draws advance on every replay
    for r in reps:
        assert r.shape == b.shape and np.abs(r).max() > 1000
        assert abs(np.sqrt(np.mean(r ** 2)) / np.sqrt(np.mean(b ** 2)) - 1) < 0.05


def test_batched_front_door_equals_per_utterance_calls():
    """padding_mask through the HuBERT object and phone_lengths (sequence_mask) through the synthesizer container for B in {2, 3}:
    frame for frame / sample for sample what the B = 1 calls return (the kernels stay B = 1)."""
    from infer.modules.vc.utils import HubertB200
    from rvc.synthesizer import get_synthesizer
    OI, OP, OW, hw, rw, sw, audio, idx = _setup(1.0, 600)
    hub = HubertB200(hw, "cuda:0")
    lens = [16000, 11000, 13500]
    wavs = [OW.synth_voice(1.0, seed=40 + i)[:n] for i, n in enumerate(lens)]
    src = torch.zeros(3, 16000)
    pm = torch.ones(3, 16000, dtype=torch.bool)
    for b, w in enumerate(wavs):
        src[b, : len(w)] = w
        pm[b, : len(w)] = False
    feats, fmask = hub.extract_features(source=src.cuda(), padding_mask=pm.cuda(), output_layer=12)
    for b, w in enumerate(wavs):
        one = hub.extract_features(source=w[None].cuda(), padding_mask=None, output_layer=12)[0][0]
        assert torch.equal(feats[b, : one.shape[0]], one) and bool((feats[b, one.shape[0]:] == 0).all())
        assert fmask is not None and int((~fmask[b]).sum()) == one.shape[0]
    net_g, _ = get_synthesizer(OW.synth_cpt(1234, "v2"), "cuda:0")
    Ts = [40, 28]
    g = torch.Generator().manual_seed(1)
    phone = torch.randn(2, 40, 768, generator=g).cuda() * 0.5
    pitch = torch.randint(1, 200, (2, 40), generator=g).cuda()
    pitchf = (torch.rand(2, 40, generator=g) * 300 + 80).cuda()
    noises = [(torch.randn(1, 192, T, generator=g), torch.randn(1, T * 480, 1, generator=g)) for T in Ts]
    for n in noises:
        net_g.set_noise(*n)
    yb = net_g.infer(phone, torch.tensor(Ts), torch.tensor([0, 0]), pitch, pitchf)
    assert yb.shape == (2, 1, 40 * 480)
    for b, T in enumerate(Ts):
        net_g.set_noise(*noises[b])
        y1 = net_g.infer(phone[b: b + 1, :T], torch.tensor([T]), torch.tensor([0]), pitch[b: b + 1, :T], pitchf[b: b + 1, :T])[0, 0]
        assert torch.equal(yb[b, 0, : T * 480], y1) and bool((yb[b, 0, T * 480:] == 0).all())

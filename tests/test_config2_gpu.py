"""GPU parity at BASELINE config #2's own shapes: 10 s utterance, x_pad = 3 -> 256 000 padded samples, 799 HuBERT frames,
1601 RMVPE frames (padded to 1632), synthesizer T = 1598 -> 767 040 decoder samples, 100 k-vector IVF2564 index, k = 8,
index_rate 0.75.  These are the shapes bench.py times: the weight-stationary / fused vocoder kernels, the split-K kernel and the
M-keyed dispatch rules are reached here by MODEL-level comparisons against the fp32 CPU oracle, not only by op-level tests.

Measured on B200 (profiles/r2a_parity_config2.json): HuBERT features 1.9e-3 max / 3.0e-4 mean, top-1 neighbours identical on
all 799 frames, synthesizer waveform 4.5e-4, end to end with shared pitch + noise 9.3e-4 of full scale.  The reference's own fp16
GPU path (the oracle modules eager on the same GPU, oracle/gpu_ref.py) is 5.2e-3 / 1.0e-3 / 2.0e-3 on the same inputs, its
fp32 (TF32 convolutions) path 1.0e-3 end to end.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

XP = 3


class Cfg:
    x_pad, x_query, x_center, x_max, is_half = XP, 10, 60, 65, True
    device = "cuda:0"
    rmvpe_state_dict = None


@pytest.fixture(scope="module")
def world():
    from scipy import signal
    from oracle import ivf as OI, pipeline as OP, synth as OS, weights as OW
    torch.set_num_threads(min(32, torch.get_num_threads() if torch.get_num_threads() > 8 else 32))
    hw, rw, sw = OW.hubert_weights(777), OW.rmvpe_weights(4321), OW.synth_weights(1234)
    audio = OW.synth_voice(10.0, seed=0).numpy()
    idx = OI.build_ivf(OW.index_vectors(100000, 768, 0).numpy(), None, seed=0, exact_assign=False)
    op = OP.OraclePipeline(48000, XP, 10, 60, 65, hw, rw, sw, OW.V2_48K_CONFIG, noise_seed=3)
    with torch.no_grad():
        ref = op.pipeline(0, audio.copy(), 0, "rmvpe", idx, 0.75, 1, 48000, 0, 0.25, "v2", 0.33)
    tap = op.taps[0]
    T = tap["phone"].shape[1]
    with torch.no_grad():
        ref_wave = OS.synth_infer(sw, OW.V2_48K_CONFIG, tap["phone"], torch.tensor([T]), torch.tensor([0]), op.pitch[:, :T], op.pitchf[:, :T],
                                  tap["noise"][0], tap["noise"][1])[0, 0].numpy()
    a_f = signal.filtfilt(OP.bh, OP.ah, audio)
    audio_pad = np.pad(a_f, (16000 * XP, 16000 * XP), mode="reflect").astype(np.float32)
    return dict(hw=hw, rw=rw, sw=sw, audio=audio, idx=idx, op=op, ref=ref, tap=tap, T=T, ref_wave=ref_wave, audio_pad=audio_pad, OW=OW)


def test_shapes_are_config2(world):
    assert world["tap"]["feats_hubert"].shape[1] == 799 and world["T"] == 1598 and world["ref"].shape[0] == 479040
    assert world["idx"].centroids.shape[0] == 2564 and world["ref_wave"].shape[0] == 767040


def test_hubert_and_retrieval_at_799_frames(world):
    from infer.modules.vc.utils import HubertB200
    from rvc_b200.engine import Index
    tap = world["tap"]
    hub = HubertB200(world["hw"], "cuda:0")
    feats = hub.extract_features(source=torch.from_numpy(world["audio_pad"])[None].cuda(), output_layer=12)[0][0]
    d = (feats.cpu() - tap["feats_hubert"][0]).abs()
    assert d.max().item() < 5e-3 and d.mean().item() < 1e-3, (d.max().item(), d.mean().item())
    gidx = Index.from_oracle_layout(world["idx"])
    _, I = gidx.search_device(feats, 8)
    I = I.cpu().numpy()
    # north_star: bit-exact argmax retrieval index -- the nearest neighbour of every frame is the oracle's
    assert np.array_equal(I[:, 0], tap["ix"][:, 0])
    assert (I == tap["ix"]).mean() >= 0.99            # ranks 2..8: fp16-operand feature noise (3e-4 mean) flips a few near-ties
    # the search itself is bit-exact (D and I) on identical queries, 100 k vectors / 2564 lists
    D2, I2 = gidx.search_device(tap["feats_hubert"][0].cuda(), 8)
    assert np.array_equal(I2.cpu().numpy(), tap["ix"]) and np.array_equal(D2.cpu().numpy(), tap["score"])


def test_synthesizer_waveform_at_T1598(world):
    from rvc_b200.engine import Synth
    tap, op, T = world["tap"], world["op"], world["T"]
    syn = Synth(world["sw"], world["OW"].V2_48K_CONFIG, 768)
    w = syn.infer(tap["phone"][0].cuda(), 0, op.pitch[0, :T].cuda(), op.pitchf[0, :T].cuda(), tap["noise"][0][0].cuda(),
                  tap["noise"][1].reshape(-1).cuda()).cpu().numpy()
    assert w.shape == world["ref_wave"].shape
    err = np.abs(w - world["ref_wave"]).max()
    assert err <= 1e-3, f"waveform max abs err at T=1598: {err}"


def test_rmvpe_f0_at_1601_frames(world):
    from infer.modules.vc.pipeline import Pipeline
    cfg = Cfg()
    cfg.rmvpe_state_dict = world["rw"]
    pipe = Pipeline(48000, cfg)
    pitch, pitchf = world["op"].pitch[0].numpy(), world["op"].pitchf[0].numpy()
    c2, f2 = pipe.f0_gen.calculate(world["audio_pad"], len(pitch), 0, "rmvpe", 3)
    assert (c2[: len(pitch)] == pitch).mean() >= 0.98
    assert np.array_equal(f2[: len(pitchf)] > 0, pitchf > 0)
    both = (f2[: len(pitchf)] > 0) & (pitchf > 0)
    assert np.median(np.abs(f2[: len(pitchf)][both] / pitchf[both] - 1)) < 1e-4


def test_end_to_end_with_shared_pitch_and_noise(world):
    """Pipeline.pipeline (pipeline.py:186-366) on the 10 s utterance with the oracle's pitch track and noise draws:
    the north_star bound, 1e-3 of full scale on the 48 kHz waveform."""
    from infer.modules.vc.pipeline import Pipeline
    from infer.modules.vc.utils import HubertB200
    from rvc.synthesizer import get_synthesizer
    from rvc_b200.engine import Index
    cfg = Cfg()
    cfg.rmvpe_state_dict = world["rw"]
    pipe = Pipeline(48000, cfg)
    hub = HubertB200(world["hw"], "cuda:0")
    net_g, _ = get_synthesizer(world["OW"].synth_cpt(1234, "v2"), "cuda:0")
    net_g.set_noise(*world["tap"]["noise"])
    pitch, pitchf = world["op"].pitch[0].numpy(), world["op"].pitchf[0].numpy()
    out = pipe.pipeline(hub, net_g, 0, world["audio"].copy(), [0, 0, 0], 0, (pitch, pitchf.astype(np.float64)), Index.from_oracle_layout(world["idx"]),
                        0.75, 2, 3, 48000, 0, 0.25, "v2", 0.33)
    assert out.shape == world["ref"].shape
    err = np.abs(out - world["ref"]).max() / 32768.0
    assert err <= 1e-3, f"end-to-end max abs err (full scale) {err}"


def test_config1_shapes_v1_40k_no_index_precomputed_f0():
    """BASELINE config #1 (the reference's own CPU-runnable case): one 10 s utterance, v1 / 40k synthesizer
    (configs/v1/40k.json, upsampling [10, 10, 2, 2]), HuBERT layer 9 + final_proj -> 256-d, no index, f0 passed in pre-computed
    (if_f0 = 2: parselmouth is absent), fp32 config (x_pad = 1): 192 000 padded samples, 599 HuBERT frames, T = 1198,
    479 200 -> 399 200 output samples.  End to end against the fp32 CPU oracle with shared noise."""
    from infer.modules.vc.pipeline import Pipeline
    from infer.modules.vc.utils import HubertB200
    from oracle import pipeline as OP, rmvpe as ORM, weights as OW
    from rvc.synthesizer import get_synthesizer

    class Cfg1:
        x_pad, x_query, x_center, x_max, is_half = 1, 6, 38, 41, False
        device = "cuda:0"
        rmvpe_state_dict = None

    torch.set_num_threads(min(32, max(8, torch.get_num_threads())))
    hw = OW.hubert_weights(777)
    sw = OW.synth_weights(1234, OW.V1_40K_CONFIG, 256)
    audio = OW.synth_voice(10.0, seed=4).numpy()
    p_len = (160000 + 2 * 16000) // 160
    f0 = 110.0 + 220.0 * np.linspace(0.0, 1.0, p_len) ** 2
    f0[(np.arange(p_len) // 100) % 5 == 4] = 0.0                       # unvoiced stretches
    pitch, pitchf = ORM.post_process(f0.copy(), 0)
    op = OP.OraclePipeline(40000, 1, 6, 38, 41, hw, None, sw, OW.V1_40K_CONFIG, noise_seed=5)
    with torch.no_grad():
        ref = op.pipeline(0, audio.copy(), 0, (pitch, pitchf), None, 0.0, 2, 40000, 0, 1.0, "v1", 0.33)
    assert op.taps[0]["feats_hubert"].shape[1:] == (599, 256) and op.taps[0]["phone"].shape[1] == 1198 and ref.shape[0] == 399200
    pipe = Pipeline(40000, Cfg1())
    hub = HubertB200(hw, "cuda:0")
    net_g, cpt = get_synthesizer(OW.synth_cpt(1234, "v1"), "cuda:0")
    net_g.set_noise(*op.taps[0]["noise"])
    out = pipe.pipeline(hub, net_g, 0, audio.copy(), [0, 0, 0], 0, (pitch, np.asarray(pitchf, dtype=np.float64)), "", 0.0, 2, 3, 40000, 0, 1.0,
                        "v1", 0.33)
    assert out.shape == ref.shape
    err = np.abs(out - ref).max() / 32768.0
    print(f"[parity] config #1 (v1/40k, layer 9 + final_proj, no index, given f0): e2e max abs err {err:.3e}")
    assert err < 1.5e-3, err

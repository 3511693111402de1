"""GPU: the training-side extraction drop-ins (infer/modules/train/extract_feature_print.py, extract_f0_print.py)
run on real wav files through the sm_100a HuBERT / RMVPE and write what the reference writes: 3_feature768/*.npy (the vectors
the retrieval index is built from), 2a_f0/*.npy (coarse bins), 2b-f0nsf/*.npy (Hz), checked against the CPU oracle."""
import os

import numpy as np
import pytest
import torch
from scipy.io import wavfile

pytestmark = pytest.mark.gpu


def test_feature_and_f0_extraction_scripts_on_wav_files(tmp_path):
    from oracle import hubert as OH, rmvpe as ORM, weights as OW
    from infer.lib.audio import load_audio
    from infer.modules.train import extract_f0_print as XF0, extract_feature_print as XFE
    from infer.modules.vc.utils import HubertB200
    from rvc.f0 import Generator
    exp = tmp_path / "logs" / "voice"
    wav_dir = exp / "1_16k_wavs"
    wav_dir.mkdir(parents=True)
    for i, sec in enumerate((1.0, 1.7, 0.8)):
        wavfile.write(str(wav_dir / f"{i}_0.wav"), 16000, OW.synth_voice(sec, seed=30 + i).numpy())
    hw, rw = OW.hubert_weights(777), OW.rmvpe_weights(4321)
    logs = []
    # ---- features: two "processes" stride the list like the reference (:110) ----
    model = HubertB200(hw, "cuda:0")
    n = sum(XFE.run(model, str(exp), "v2", 2, part, lambda p: load_audio(p, 16000), logs.append) for part in (0, 1))
    assert n == 3 and sorted(os.listdir(exp / "3_feature768")) == ["0_0.npy", "1_0.npy", "2_0.npy"]
    for name in ("0_0", "1_0", "2_0"):
        got = np.load(exp / "3_feature768" / f"{name}.npy")
        wav = torch.from_numpy(load_audio(str(wav_dir / f"{name}.wav"), 16000))
        with torch.no_grad():
            ref = OH.extract_features(hw, wav[None], 12)[0].numpy()
        assert got.shape == ref.shape and got.dtype == np.float32
        assert np.abs(got - ref).max() < 5e-3 and np.abs(got - ref).mean() < 1e-3
    assert any(m.startswith("all-feature-") for m in logs) and logs[-1] == "all-feature-done"
    # v1: layer 9 + final_proj -> 256-d
    XFE.run(model, str(exp), "v1", 1, 0, lambda p: load_audio(p, 16000), logs.append)
    f1 = np.load(exp / "3_feature256" / "0_0.npy")
    wav = torch.from_numpy(load_audio(str(wav_dir / "0_0.wav"), 16000))
    with torch.no_grad():
        r1 = OH.final_proj(hw, OH.extract_features(hw, wav[None], 9))[0].numpy()
    assert f1.shape == r1.shape and f1.shape[1] == 256 and np.abs(f1 - r1).max() < 5e-3
    # ---- f0 ----
    gen = Generator(rw, False, 0, "cuda:0", 160, 16000)
    jobs = XF0.list_jobs(str(exp))
    assert XF0.run(gen, jobs, "rmvpe", lambda p: load_audio(p, 16000), logs.append) == 3
    for name in ("0_0", "1_0", "2_0"):
        coarse = np.load(exp / "2a_f0" / f"{name}.wav.npy")
        f0 = np.load(exp / "2b-f0nsf" / f"{name}.wav.npy")
        wav = load_audio(str(wav_dir / f"{name}.wav"), 16000)
        rc, rf = ORM.calculate(rw, wav, wav.shape[0] // 160, 0)
        assert coarse.shape == rc.shape and f0.shape == rf.shape
        assert (coarse == rc).mean() >= 0.97 and np.array_equal(f0 > 0, rf > 0)
        both = (f0 > 0) & (rf > 0)
        assert np.median(np.abs(f0[both] / rf[both] - 1)) < 1e-4
    # second run: everything exists already, nothing is rewritten
    assert XF0.run(gen, jobs, "rmvpe", lambda p: load_audio(p, 16000), logs.append) == 0

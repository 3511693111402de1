"""GPU parity: RMVPE (log-mel front end, DeepUnet + BiGRU salience, decode) vs the fp32 CPU oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seconds", [0.31, 2.0])
def test_rmvpe_mel_hidden_f0(seconds):
    from oracle import rmvpe as ORM, weights as OW
    from rvc_b200.engine import Rmvpe
    w = OW.rmvpe_weights(4321)
    wav = OW.synth_voice(seconds, seed=1)
    taps = {}
    with torch.no_grad():
        ORM.compute_f0(w, wav.numpy(), None, 0.03, taps)
    m = Rmvpe(w)
    f0, mel, hid = m.infer(wav.cuda(), 0.03, want_mel=True, want_hidden=True)
    mel_ref = taps["mel"][0]
    assert mel.shape == mel_ref.shape
    # fp32 DFT-as-GEMM vs torch.stft: compare where the mel energy is above the clamp floor
    loud = mel_ref > -9.0
    assert (mel.cpu() - mel_ref)[loud].abs().max().item() < 5e-3
    hid_ref = torch.from_numpy(taps["hidden"])
    herr = (hid.cpu() - hid_ref).abs()
    # fp16 tensor-core operands through ~90 conv layers + GRU: salience is a sigmoid output in (0,1)
    assert herr.max().item() < 3e-2 and herr.mean().item() < 3e-3, (herr.max().item(), herr.mean().item())
    # the decode itself (argmax + local average) is exact given the same salience
    f0_ref = ORM.decode(hid.cpu().numpy().astype(np.float32), 0.03)
    assert np.abs(f0.cpu().numpy() - f0_ref).max() < 1e-3

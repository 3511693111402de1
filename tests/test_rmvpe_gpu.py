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
    # DFT-as-GEMM (split-precision fp16 hi/lo operands, fp32 accumulate) vs torch.stft: compare where the mel energy is above the
    # clamp floor
    loud = mel_ref > -9.0
    assert (mel.cpu() - mel_ref)[loud].abs().max().item() < 5e-3
    hid_ref = torch.from_numpy(taps["hidden"])
    herr = (hid.cpu() - hid_ref).abs()
    # fp16 tensor-core operands through ~90 conv layers + GRU: salience is a sigmoid output in (0,1)
    assert herr.max().item() < 3e-2 and herr.mean().item() < 3e-3, (herr.max().item(), herr.mean().item())
    # the decode itself (argmax + local average) is exact given the same salience
    f0_ref = ORM.decode(hid.cpu().numpy().astype(np.float32), 0.03)
    assert np.abs(f0.cpu().numpy() - f0_ref).max() < 1e-3


@pytest.mark.parametrize("gain", [1.0, 3e-2, 1e-3])
def test_rmvpe_mel_split_precision_is_level_independent(gain):
    """The tensor-core DFT splits signal and basis into fp16 hi + lo halves (three products, fp32 accumulate): its error must stay
    at the fp32 level whatever the input level (the realtime path feeds un-normalised microphone blocks), also for a weak partial
    next to a strong one (the rounding error of the strong partial leaks into every bin).  Judged against a float64 DFT: the GPU
    log-mel may be at most a few times as far from it as the oracle's own float32 torch.stft is."""
    from oracle import rmvpe as ORM, weights as OW
    from rvc_b200.engine import Rmvpe
    w = OW.rmvpe_weights(4321)
    t = torch.arange(16000) / 16000.0
    wav = (0.9 * torch.sin(2 * np.pi * 220.0 * t) + 1e-3 * torch.sin(2 * np.pi * 3301.0 * t)
           + 0.05 * torch.randn(16000, generator=torch.Generator().manual_seed(2)) * (t > 0.5)) * gain
    mel32 = ORM.log_mel(wav[None])[0]
    fft64 = torch.stft(wav.double(), n_fft=1024, hop_length=160, win_length=1024, window=torch.hann_window(1024, dtype=torch.float64),
                       center=True, return_complex=True)
    mel64 = torch.log(torch.clamp(torch.from_numpy(ORM.mel_filterbank()).double() @ fft64.abs(), min=1e-5))
    mel = Rmvpe(w).infer(wav.cuda(), 0.03, want_mel=True, want_hidden=False)[1].cpu().double()
    ok = mel64 > -11.0                                    # above the log clamp floor (log 1e-5 = -11.5)
    e_gpu = (mel - mel64)[ok].abs().max().item()
    e_ref = (mel32.double() - mel64)[ok].abs().max().item()
    print(f"[parity] log-mel vs float64 DFT at gain {gain:g}: GPU split-fp16 {e_gpu:.2e}, oracle float32 stft {e_ref:.2e} "
          f"({int(ok.sum())} of {ok.numel()} bins)")
    assert ok.float().mean().item() > 0.3
    assert e_gpu <= max(4.0 * e_ref, 1e-3), (e_gpu, e_ref)

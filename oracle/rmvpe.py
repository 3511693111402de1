"""Oracle: RMVPE f0 (log-mel -> DeepUnet -> BiGRU -> salience -> Hz) and the f0
post-processing chain, fp32/fp64 CPU restatement.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Reference sites restated:
  MelSpectrogram.forward       rvc/f0/mel.py:51-71 (librosa.filters.mel(htk=True) restated,
                               librosa is not installed: SURVEY §8c)
  STFT.forward (torch.stft)    rvc/f0/stft.py:154-181
  RMVPE._mel2hidden            rvc/f0/rmvpe.py:139-155
  E2E / DeepUnet / BiGRU       rvc/f0/e2e.py:8-67, rvc/f0/deepunet.py:7-217
  RMVPE._decode / local avg    rvc/f0/rmvpe.py:119-137, 157-164
  F0Predictor._resize_f0       rvc/f0/f0.py:69-78
  F0Predictor._interpolate_f0  rvc/f0/f0.py:31-67
  post_process                 rvc/f0/gen.py:10-41
  Generator.calculate("rmvpe") rvc/f0/gen.py:61-141
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

W = Dict[str, torch.Tensor]
N_FFT, HOP, N_MELS, SR = 1024, 160, 128, 16000
CENTS0 = 1997.3794084376191


def mel_filterbank(sr=SR, n_fft=N_FFT, n_mels=N_MELS, fmin=30.0, fmax=8000.0) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=True, norm='slaney') -> f32 [n_mels, 1+n_fft/2]."""
    fftfreqs = np.linspace(0, sr / 2.0, 1 + n_fft // 2)
    lo, hi = 2595.0 * np.log10(1.0 + fmin / 700.0), 2595.0 * np.log10(1.0 + fmax / 700.0)
    mels = np.linspace(lo, hi, n_mels + 2)
    mel_f = 700.0 * (10.0 ** (mels / 2595.0) - 1.0)
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    wts = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        wts[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    wts *= enorm[:, None]
    return wts.astype(np.float32)


def log_mel(wav: torch.Tensor) -> torch.Tensor:
    """wav f32 [B, N] -> log-mel [B, 128, N//160 + 1]  (mel.py:51-71, fp32 path)."""
    win = torch.hann_window(N_FFT, device=wav.device)
    fft = torch.stft(wav, n_fft=N_FFT, hop_length=HOP, win_length=N_FFT, window=win, center=True,
                     return_complex=True)
    mag = torch.sqrt(fft.real.pow(2) + fft.imag.pow(2))
    mel = torch.matmul(torch.from_numpy(mel_filterbank()).to(mag.device), mag)
    return torch.log(torch.clamp(mel, min=1e-5))


def _bn(w: W, p: str, x):
    return F.batch_norm(x, w[p + "running_mean"], w[p + "running_var"], w[p + "weight"], w[p + "bias"], False, 0.0, 1e-5)


def _cbr(w: W, p: str, x):
    y = F.relu(_bn(w, p + "conv.1.", F.conv2d(x, w[p + "conv.0.weight"], None, padding=1)))
    y = F.relu(_bn(w, p + "conv.4.", F.conv2d(y, w[p + "conv.3.weight"], None, padding=1)))
    if p + "shortcut.weight" in w:
        return y + F.conv2d(x, w[p + "shortcut.weight"], w[p + "shortcut.bias"])
    return y + x


_GRU_CACHE: dict = {}


def _gru_module(w: W) -> torch.nn.GRU:
    """torch.nn.GRU carrying the fc.0.gru.* weights (deepunet BiGRU, e2e.py:8-67), built once per weight dict."""
    ref = w["fc.0.gru.weight_ih_l0"]
    key = (id(w), ref.device, ref.dtype)
    if key not in _GRU_CACHE:
        gru = torch.nn.GRU(384, 256, num_layers=1, batch_first=True, bidirectional=True)
        gru.load_state_dict({k[len("fc.0.gru."):]: v for k, v in w.items() if k.startswith("fc.0.gru.")})
        _GRU_CACHE[key] = gru.to(device=ref.device, dtype=ref.dtype).eval().requires_grad_(False)
    return _GRU_CACHE[key]


def e2e_forward(w: W, mel: torch.Tensor, n_blocks=4, en_de=5, inter=4, taps: Optional[dict] = None) -> torch.Tensor:
    """mel [B,128,T] (T % 32 == 0) -> salience [B,T,360]."""
    x = mel.transpose(-1, -2).unsqueeze(1)
    x = _bn(w, "unet.encoder.bn.", x)
    skips = []
    for l in range(en_de):
        for b in range(n_blocks):
            x = _cbr(w, f"unet.encoder.layers.{l}.conv.{b}.", x)
        skips.append(x)
        x = F.avg_pool2d(x, 2)
    if taps is not None:
        taps["enc_out"] = x
    for l in range(inter):
        for b in range(n_blocks):
            x = _cbr(w, f"unet.intermediate.layers.{l}.conv.{b}.", x)
    if taps is not None:
        taps["inter_out"] = x
    for l in range(en_de):
        p = f"unet.decoder.layers.{l}."
        x = F.conv_transpose2d(x, w[p + "conv1.0.weight"], None, stride=2, padding=1, output_padding=1)
        x = F.relu(_bn(w, p + "conv1.1.", x))
        x = torch.cat((x, skips[-1 - l]), dim=1)
        for b in range(n_blocks):
            x = _cbr(w, p + f"conv2.{b}.", x)
    if taps is not None:
        taps["unet_out"] = x
    x = F.conv2d(x, w["cnn.weight"], w["cnn.bias"], padding=1)
    x = x.transpose(1, 2).flatten(-2)                       # [B,T,384]
    if taps is not None:
        taps["gru_in"] = x
    x = _gru_module(w)(x)[0]
    if taps is not None:
        taps["gru_out"] = x
    return torch.sigmoid(F.linear(x, w["fc.1.weight"], w["fc.1.bias"]))


def mel2hidden(w: W, mel: torch.Tensor) -> torch.Tensor:
    n_frames = mel.shape[-1]
    n_pad = 32 * ((n_frames - 1) // 32 + 1) - n_frames
    if n_pad > 0:
        mel = F.pad(mel, (0, n_pad), mode="constant")
    return e2e_forward(w, mel.to(w["cnn.weight"].dtype))[:, :n_frames]


def decode(salience: np.ndarray, thred: float = 0.03) -> np.ndarray:
    """rmvpe.py:119-137,157-164 (vectorised, same arithmetic in float64 like numpy does
    for float32 salience * float64 cents table)."""
    cents_mapping = np.pad(20 * np.arange(360) + CENTS0, (4, 4))
    center = np.argmax(salience, axis=1)
    sal = np.pad(salience, ((0, 0), (4, 4)))
    center = center + 4
    idx = center[:, None] + np.arange(-4, 5)[None, :]
    ts = np.take_along_axis(sal, idx, axis=1)
    tc = cents_mapping[idx]
    devided = np.sum(ts * tc, 1) / np.sum(ts, 1)
    devided[np.max(sal, axis=1) <= thred] = 0
    f0 = 10 * (2 ** (devided / 1200))
    f0[f0 == 10] = 0
    return f0


def resize_f0(x: np.ndarray, target_len: int) -> np.ndarray:
    source = np.array(x)
    source[source < 0.001] = np.nan
    target = np.interp(np.arange(0, len(source) * target_len, len(source)) / target_len,
                       np.arange(0, len(source)), source)
    return np.nan_to_num(target)


def interpolate_f0(f0: np.ndarray) -> np.ndarray:
    """f0.py:31-67, literal loop semantics (in-place on a copy)."""
    data = np.array(f0, dtype=np.float64).reshape(-1)
    ip = data
    n = data.size
    last_value = 0.0
    i = 0
    while i < n:
        if data[i] <= 0.0:
            j = i + 1
            for j in range(i + 1, n):
                if data[j] > 0.0:
                    break
            if j < n - 1:
                if last_value > 0.0:
                    step = (data[j] - data[i - 1]) / float(j - i)
                    for k in range(i, j):
                        ip[k] = data[i - 1] + step * (k - i + 1)
                else:
                    for k in range(i, j):
                        ip[k] = data[j]
            else:
                for k in range(i, n):
                    ip[k] = last_value
        else:
            last_value = data[i]
        i += 1
    return ip


def post_process(f0: np.ndarray, f0_up_key: float, f0_min=50.0, f0_max=1100.0) -> Tuple[np.ndarray, np.ndarray]:
    """gen.py:10-41 without the manual_f0 splice."""
    f0 = np.multiply(f0, pow(2, f0_up_key / 12))
    mel_min = 1127 * math.log(1 + f0_min / 700)
    mel_max = 1127 * math.log(1 + f0_max / 700)
    f0_mel = 1127 * np.log(1 + f0 / 700)
    f0_mel[f0_mel > 0] = (f0_mel[f0_mel > 0] - mel_min) * 254 / (mel_max - mel_min) + 1
    f0_mel[f0_mel <= 1] = 1
    f0_mel[f0_mel > 255] = 255
    return np.rint(f0_mel).astype(np.int32), f0


def compute_f0(w: W, wav: np.ndarray, p_len: Optional[int] = None, thred: float = 0.03,
               taps: Optional[dict] = None) -> np.ndarray:
    """RMVPE.compute_f0 (rmvpe.py:96-117)."""
    if p_len is None:
        p_len = wav.shape[0] // HOP
    mel = log_mel(torch.from_numpy(np.asarray(wav, dtype=np.float32))[None])
    hidden = mel2hidden(w, mel)[0].numpy()
    if taps is not None:
        taps["mel"] = mel
        taps["hidden"] = hidden
    f0 = decode(hidden, thred)
    if taps is not None:
        taps["f0_raw"] = f0.copy()
    return interpolate_f0(resize_f0(f0, p_len))


def calculate(w: W, x: np.ndarray, p_len: Optional[int], f0_up_key: float) -> Tuple[np.ndarray, np.ndarray]:
    """rvc.f0.Generator.calculate(x, p_len, key, "rmvpe", ...) (gen.py:61-141)."""
    return post_process(compute_f0(w, x, p_len, 0.03), f0_up_key)

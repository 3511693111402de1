"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the realtime GUI's spectral gate.

Follows /root/reference/infer/modules/gui/torchgate.py (TorchGate.forward :217-280, _stationary_mask :128-178,
_nonstationary_mask :180-215, _generate_mask_smoothing_filter :72-126) and infer/modules/gui/utils.py (amp_to_db :5-24,
temperature_sigmoid :27-40, linspace :43-70) for the non-DirectML branch (torch.stft / torch.istft, pad_mode="constant").
Pinned against the reference's own class by tests/golden/torchgate_*.npz (tests/golden/make_golden.py: torchgate())."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

_EPS = torch.finfo(torch.float64).eps


def smoothing_filter(sr: int, n_fft: int, hop: int, freq_mask_smooth_hz: Optional[float] = 500, time_mask_smooth_ms: Optional[float] = 50):
    """torchgate.py:72-126 -> [n_f, n_t] outer product of two triangles, normalised to sum 1; None when both widths are 1."""
    if freq_mask_smooth_hz is None and time_mask_smooth_ms is None:
        return None
    nf = 1 if freq_mask_smooth_hz is None else int(freq_mask_smooth_hz / (sr / (n_fft / 2)))
    nt = 1 if time_mask_smooth_ms is None else int(time_mask_smooth_ms / ((hop / sr) * 1000))
    if nf < 1 or nt < 1:
        raise ValueError("smoothing width below one bin / frame")
    if nf == 1 and nt == 1:
        return None

    def tri(n):
        up = torch.linspace(0, 1, n + 2)[:-1]              # linspace(0, 1, n + 1, endpoint=False)
        down = torch.linspace(1, 0, n + 2)
        return torch.cat([up, down])[1:-1]
    f = torch.outer(tri(nf), tri(nt))
    return f / f.sum()


def amp_to_db(x: torch.Tensor, top_db: float = 40.0) -> torch.Tensor:
    """utils.py:5-24: 20 log10(|x| + eps), floored at (max over the LAST axis = time) - top_db."""
    x_db = 20 * torch.log10(x.abs() + _EPS)
    return torch.max(x_db, (x_db.max(-1).values - top_db).unsqueeze(-1))


def _stft(x: torch.Tensor, n_fft: int, hop: int) -> torch.Tensor:
    return torch.stft(x, n_fft=n_fft, hop_length=hop, win_length=n_fft, return_complex=True, pad_mode="constant", center=True,
                      window=torch.hann_window(n_fft))


def torchgate(x: torch.Tensor, xn: Optional[torch.Tensor], sr: int, n_fft: int = 1024, hop: Optional[int] = None,
              nonstationary: bool = False, n_std_thresh_stationary: float = 1.5, n_thresh_nonstationary: float = 1.3,
              temp_coeff_nonstationary: float = 0.1, n_movemean_nonstationary: int = 20, prop_decrease: float = 1.0,
              freq_mask_smooth_hz: Optional[float] = 500, time_mask_smooth_ms: Optional[float] = 50) -> torch.Tensor:
    """x: [B, L] float32, xn: [B, Ln] or None -> [B, hop * (n_frames - 1)]."""
    hop = n_fft // 4 if hop is None else hop
    X = _stft(x, n_fft, hop)                                                   # [B, F, T]
    if nonstationary:
        X_abs = X.abs()
        k = n_movemean_nonstationary
        sm = F.conv1d(X_abs.reshape(-1, 1, X_abs.shape[-1]), torch.ones(1, 1, k, dtype=X_abs.dtype), padding="same").view(X_abs.shape) / k
        ratio = (X_abs - sm) / (sm + 1e-6)
        mask = torch.sigmoid((ratio - n_thresh_nonstationary) / temp_coeff_nonstationary)
    else:
        X_db = amp_to_db(X)
        XN_db = amp_to_db(_stft(xn, n_fft, hop)).to(X_db.dtype) if xn is not None else X_db
        std, mean = torch.std_mean(XN_db, dim=-1)
        mask = X_db > (mean + std * n_std_thresh_stationary).unsqueeze(2)
    mask = prop_decrease * (mask.float() - 1.0) + 1.0
    filt = smoothing_filter(sr, n_fft, hop, freq_mask_smooth_hz, time_mask_smooth_ms)
    if filt is not None:
        mask = F.conv2d(mask.unsqueeze(1), filt[None, None].to(mask.dtype), padding="same").squeeze(1)
    Y = X * mask
    y = torch.istft(Y, n_fft=n_fft, hop_length=hop, win_length=n_fft, center=True, window=torch.hann_window(n_fft))
    return y.to(x.dtype)

"""Drop-in for the reference's realtime spectral gate (infer/modules/gui/torchgate.py: TorchGate) on the B200 library: same
constructor arguments, ``forward(x, xn=None)`` with x [B, L] / xn [B, Ln] device tensors -> [B, hop * (L // hop)], ``.to(device)``.
The STFT, the dB statistics, the mask, its smoothing, the inverse STFT and the overlap-add are CUDA kernels behind
``rvcb_torchgate_apply`` (csrc/torchgate.cu); this file only derives the smoothing filter (torchgate.py:72-126, a handful of
floats) and owns the handle.  There is no CPU path: without the CUDA library the constructor raises."""
from __future__ import annotations

from typing import Optional

import torch

from rvc_b200 import engine


def _linspace(start, stop, num, endpoint=True):
    return torch.linspace(start, stop, num) if endpoint else torch.linspace(start, stop, num + 1)[:-1]      # gui/utils.py:43-70


class TorchGate:
    def __init__(self, sr: int, nonstationary: bool = False, n_std_thresh_stationary: float = 1.5, n_thresh_nonstationary: float = 1.3,
                 temp_coeff_nonstationary: float = 0.1, n_movemean_nonstationary: int = 20, prop_decrease: float = 1.0, n_fft: int = 1024,
                 win_length: Optional[int] = None, hop_length: Optional[int] = None, freq_mask_smooth_hz: Optional[float] = 500,
                 time_mask_smooth_ms: Optional[float] = 50):
        assert 0.0 <= prop_decrease <= 1.0
        self.sr, self.nonstationary, self.prop_decrease = sr, nonstationary, prop_decrease
        self.n_fft = n_fft
        self.win_length = n_fft if win_length is None else win_length
        if self.win_length != n_fft:
            raise NotImplementedError("win_length != n_fft (the reference never passes it: gui.py:869-871)")
        self.hop_length = self.win_length // 4 if hop_length is None else hop_length
        self.n_std_thresh_stationary = n_std_thresh_stationary
        self.temp_coeff_nonstationary = temp_coeff_nonstationary
        self.n_movemean_nonstationary = n_movemean_nonstationary
        self.n_thresh_nonstationary = n_thresh_nonstationary
        self.freq_mask_smooth_hz, self.time_mask_smooth_ms = freq_mask_smooth_hz, time_mask_smooth_ms
        self.smoothing_filter = self._generate_mask_smoothing_filter()
        self._h = None
        self._device_index = 0

    def _generate_mask_smoothing_filter(self):
        if self.freq_mask_smooth_hz is None and self.time_mask_smooth_ms is None:
            return None
        n_grad_freq = 1 if self.freq_mask_smooth_hz is None else int(self.freq_mask_smooth_hz / (self.sr / (self.n_fft / 2)))
        if n_grad_freq < 1:
            raise ValueError(f"freq_mask_smooth_hz needs to be at least {int((self.sr / (self.n_fft / 2)))} Hz")
        n_grad_time = 1 if self.time_mask_smooth_ms is None else int(self.time_mask_smooth_ms / ((self.hop_length / self.sr) * 1000))
        if n_grad_time < 1:
            raise ValueError(f"time_mask_smooth_ms needs to be at least {int((self.hop_length / self.sr) * 1000)} ms")
        if n_grad_time == 1 and n_grad_freq == 1:
            return None
        v_f = torch.cat([_linspace(0, 1, n_grad_freq + 1, endpoint=False), _linspace(1, 0, n_grad_freq + 2)])[1:-1]
        v_t = torch.cat([_linspace(0, 1, n_grad_time + 1, endpoint=False), _linspace(1, 0, n_grad_time + 2)])[1:-1]
        f = torch.outer(v_f, v_t).unsqueeze(0).unsqueeze(0)
        return f / f.sum()

    def to(self, device):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("TorchGate (B200) runs on a CUDA device only")
        self._device_index = dev.index or 0
        return self

    def _handle(self):
        if self._h is None:
            filt = None if self.smoothing_filter is None else self.smoothing_filter[0, 0]
            self._h = engine.TorchGateHandle(self.sr, self.n_fft, self.hop_length, self.nonstationary, self.n_std_thresh_stationary,
                                             self.n_thresh_nonstationary, self.temp_coeff_nonstationary, self.n_movemean_nonstationary,
                                             self.prop_decrease, filt, self._device_index)
        return self._h

    @torch.no_grad()
    def forward(self, x: torch.Tensor, xn: Optional[torch.Tensor] = None) -> torch.Tensor:
        if x.dim() != 2:
            raise ValueError("x must be [batch, samples]")
        h = self._handle()
        rows = [h.apply(x[b].float(), None if xn is None else xn[b].float()) for b in range(x.shape[0])]
        return torch.stack(rows).to(dtype=x.dtype)

    __call__ = forward

"""Drop-in for the reference's realtime spectral gate (infer/modules/gui/torchgate.py: TorchGate) on the B200 library: same
constructor arguments, ``forward(x, xn=None)`` with x [B, L] / xn [B, Ln] device tensors -> [B, hop * (L // hop)], ``.to(device)``.
The STFT, the dB statistics, the mask, its smoothing, the inverse STFT and the overlap-add are CUDA kernels behind
``rvcb_torchgate_apply`` (csrc/torchgate.cu); this file derives the mask-smoothing filter (torchgate.py:72-126, a handful of
floats) and owns the handle.  There is no CPU path: the first call creates the CUDA handle or raises."""
from __future__ import annotations

from typing import Optional

import torch

from rvc_b200 import engine


def _ramp(n: int) -> torch.Tensor:
    """n rising steps k / (n + 1), the peak 1, n falling steps: 2 n + 1 values, the same torch.linspace calls as the reference."""
    return torch.cat([torch.linspace(0, 1, n + 2)[1:-1], torch.linspace(1, 0, n + 2)[:-1]])


def mask_smoothing_filter(sr: int, n_fft: int, hop: int, freq_hz: Optional[float], time_ms: Optional[float]) -> Optional[torch.Tensor]:
    """[2 nf + 1, 2 nt + 1] outer product of two triangles normalised to sum 1 (rows = frequency), or None when there is nothing to
    smooth; nf / nt = the smoothing widths in bins / frames, truncated like the reference does."""
    if freq_hz is None and time_ms is None:
        return None
    bin_hz, frame_ms = sr / (n_fft / 2), (hop / sr) * 1000
    nf = 1 if freq_hz is None else int(freq_hz / bin_hz)
    nt = 1 if time_ms is None else int(time_ms / frame_ms)
    if nf < 1:
        raise ValueError(f"freq_mask_smooth_hz needs to be at least {int(bin_hz)} Hz")
    if nt < 1:
        raise ValueError(f"time_mask_smooth_ms needs to be at least {int(frame_ms)} ms")
    if nf == 1 and nt == 1:
        return None
    f = torch.outer(_ramp(nf), _ramp(nt))
    return f / f.sum()


class TorchGate:
    def __init__(self, sr: int, nonstationary: bool = False, n_std_thresh_stationary: float = 1.5, n_thresh_nonstationary: float = 1.3,
                 temp_coeff_nonstationary: float = 0.1, n_movemean_nonstationary: int = 20, prop_decrease: float = 1.0, n_fft: int = 1024,
                 win_length: Optional[int] = None, hop_length: Optional[int] = None, freq_mask_smooth_hz: Optional[float] = 500,
                 time_mask_smooth_ms: Optional[float] = 50):
        if not 0.0 <= prop_decrease <= 1.0:
            raise AssertionError("prop_decrease must be in [0, 1]")
        if win_length not in (None, n_fft):
            raise NotImplementedError("win_length != n_fft (the reference never passes it: gui.py:869-871)")
        self.sr, self.n_fft, self.win_length = sr, n_fft, n_fft
        self.hop_length = n_fft // 4 if hop_length is None else hop_length
        self.nonstationary, self.prop_decrease = nonstationary, prop_decrease
        self.n_std_thresh_stationary = n_std_thresh_stationary
        self.n_thresh_nonstationary, self.temp_coeff_nonstationary = n_thresh_nonstationary, temp_coeff_nonstationary
        self.n_movemean_nonstationary = n_movemean_nonstationary
        self.freq_mask_smooth_hz, self.time_mask_smooth_ms = freq_mask_smooth_hz, time_mask_smooth_ms
        filt = mask_smoothing_filter(sr, n_fft, self.hop_length, freq_mask_smooth_hz, time_mask_smooth_ms)
        self.smoothing_filter = None if filt is None else filt[None, None]          # [1, 1, rows, cols] like the reference's buffer
        self._handle_obj, self._device_index = None, 0

    def to(self, device):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("TorchGate (B200) runs on a CUDA device only")
        self._device_index = dev.index or 0
        return self

    def _handle(self) -> "engine.TorchGateHandle":
        if self._handle_obj is None:
            filt = None if self.smoothing_filter is None else self.smoothing_filter[0, 0]
            self._handle_obj = engine.TorchGateHandle(self.sr, self.n_fft, self.hop_length, self.nonstationary, self.n_std_thresh_stationary,
                                                      self.n_thresh_nonstationary, self.temp_coeff_nonstationary,
                                                      self.n_movemean_nonstationary, self.prop_decrease, filt, self._device_index)
        return self._handle_obj

    @torch.no_grad()
    def forward(self, x: torch.Tensor, xn: Optional[torch.Tensor] = None) -> torch.Tensor:
        if x.dim() != 2:
            raise ValueError("x must be [batch, samples]")
        h = self._handle()
        rows = [h.apply(x[b].float(), None if xn is None else xn[b].float()) for b in range(x.shape[0])]
        return torch.stack(rows).to(dtype=x.dtype)

    __call__ = forward

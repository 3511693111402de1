"""Drop-in for ``torchaudio.transforms.Resample`` as the realtime GUI uses it (gui.py:851-866: ``tat.Resample(orig_freq, new_freq,
dtype=torch.float32).to(device)``; applied at gui.py:991-1000 and 1008-1009): the windowed-sinc table is built with torchaudio's
own formula (rvc_b200.engine.sinc_resample_kernel), the strided convolution is the CUDA kernel ``rvcb_resample_sinc``."""
from __future__ import annotations

import torch

from rvc_b200 import engine


class Resample:
    def __init__(self, orig_freq: int = 16000, new_freq: int = 16000, resampling_method: str = "sinc_interp_hann",
                 lowpass_filter_width: int = 6, rolloff: float = 0.99, beta=None, *, dtype=torch.float32):
        if resampling_method != "sinc_interp_hann":
            raise NotImplementedError("only sinc_interp_hann (the reference's default) is built")
        self.orig_freq, self.new_freq = int(orig_freq), int(new_freq)
        self.kernel = None
        if self.orig_freq != self.new_freq:
            self.kernel, self.width, self.up, self.down = engine.sinc_resample_kernel(orig_freq, new_freq, lowpass_filter_width, rolloff, dtype)

    def to(self, device):
        if torch.device(device).type != "cuda":
            raise RuntimeError("Resample (B200) runs on a CUDA device only")
        if self.kernel is not None:
            self.kernel = self.kernel.to(device)
        return self

    @torch.no_grad()
    def forward(self, waveform: torch.Tensor) -> torch.Tensor:
        if self.kernel is None:
            return waveform
        if not self.kernel.is_cuda:
            self.kernel = self.kernel.to(waveform.device)
        shape = waveform.shape
        rows = [engine.sinc_resample(r.float(), self.kernel, self.width, self.up, self.down) for r in waveform.reshape(-1, shape[-1])]
        return torch.stack(rows).reshape(shape[:-1] + rows[0].shape).to(waveform.dtype)

    __call__ = forward

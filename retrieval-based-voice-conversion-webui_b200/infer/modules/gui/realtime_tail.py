"""Device-side tail of gui.py's realtime audio callback (gui.py:1024-1087): what happens to ``infer_wav`` between
``self.rvc.infer(...)`` and ``outdata``.  The GUI itself (FreeSimpleGUI, sounddevice, the audio process) is out of scope;
this object carries the callback's block geometry (gui.py:783-855) and its SOLA state so that one realtime block is
``RVC.infer`` + ``RealtimeTail.process`` with a single D2H copy of ``block_frame`` samples at the end.

    tail = RealtimeTail(samplerate=48000, block_time=0.16, crossfade_time=0.05, extra_time=2.5, device="cuda:0")
    y = rvc.infer(input_wav_res, tail.block_frame_16k, tail.skip_head, tail.return_length, "rmvpe")
    out = tail.process(y, input_wav[tail.extra_frame:], rms_mix_rate)        # f32 [block_frame] on the device

The rest of the callback's device side (input rings, TorchGate noise reduction, resamplers) is ``RealtimeBlock`` in
realtime_block.py, which owns one of these; the phase-vocoder cross-fade (use_pv, gui.py:27-48) is the ``use_pv`` flag.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from rvc_b200 import engine


class RealtimeTail:
    def __init__(self, samplerate: int = 48000, block_time: float = 0.25, crossfade_time: float = 0.05, extra_time: float = 2.5,
                 device="cuda:0", use_pv: bool = False):
        self.use_pv = use_pv                       # gui.py:1078-1083: phase-vocoder cross-fade instead of the sin^2 cross-fade
        self.device = torch.device(device if "cuda" in str(device) else "cuda:0")
        self.samplerate = samplerate
        self.zc = samplerate // 100                                                                           # gui.py:783
        self.block_frame = int(np.round(block_time * samplerate / self.zc)) * self.zc
        self.block_frame_16k = 160 * self.block_frame // self.zc
        self.crossfade_frame = int(np.round(crossfade_time * samplerate / self.zc)) * self.zc
        self.sola_buffer_frame = min(self.crossfade_frame, 4 * self.zc)
        self.sola_search_frame = self.zc
        self.extra_frame = int(np.round(extra_time * samplerate / self.zc)) * self.zc
        self.input_frames = self.extra_frame + self.crossfade_frame + self.sola_search_frame + self.block_frame   # len(input_wav)
        self.input_frames_16k = 160 * self.input_frames // self.zc                                                # len(input_wav_res)
        self.skip_head = self.extra_frame // self.zc
        self.return_length = (self.block_frame + self.sola_buffer_frame + self.sola_search_frame) // self.zc
        self.sola_buffer = torch.zeros(self.sola_buffer_frame, device=self.device, dtype=torch.float32)
        self.last_offset: Optional[torch.Tensor] = None

    def reset(self):
        self.sola_buffer.zero_()

    @torch.no_grad()
    def process(self, infer_wav: torch.Tensor, input_wav: Optional[torch.Tensor] = None, rms_mix_rate: float = 1.0,
                want_offset: bool = False) -> torch.Tensor:
        """infer_wav f32[>= block + sola_buffer + sola_search] (device; scaled in place when rms_mix_rate < 1);
        input_wav: the input window from ``extra_frame`` on, at the output rate (needed when rms_mix_rate < 1, gui.py:1025-1028)."""
        need = self.block_frame + self.sola_buffer_frame + self.sola_search_frame
        if infer_wav.shape[0] < need:
            raise ValueError(f"infer_wav has {infer_wav.shape[0]} samples, the tail needs {need}")
        if rms_mix_rate < 1 and input_wav is None:
            raise ValueError("rms_mix_rate < 1 needs input_wav")
        res = engine.rt_tail(infer_wav, input_wav if rms_mix_rate < 1 else None, self.zc, rms_mix_rate, self.sola_buffer, self.block_frame,
                             self.sola_search_frame, want_offset, self.use_pv)
        if want_offset:
            self.last_offset = res[1]
            return res[0]
        return res

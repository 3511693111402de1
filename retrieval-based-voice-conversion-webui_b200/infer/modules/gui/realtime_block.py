"""Device side of gui.py's realtime audio callback (gui.py:783-871 state, :940-1090 one block, ``function == "vc"``): everything
between the host block ``indata`` and ``outdata`` -- rolling input windows, optional input noise gate (TorchGate) with its
cross-fade, resampling to 16 kHz, ``rtrvc.RVC.infer``, resampling back when the model rate differs, optional output noise gate,
volume-envelope mix, SOLA -- runs on the GPU as ONE CUDA graph per block (captured on the second block with the same settings),
with one H2D copy of the block in and one D2H copy of ``block_frame`` samples out.  The GUI itself (FreeSimpleGUI, sounddevice,
the audio process, device enumeration) is out of scope; the response-threshold gate (gui.py:951-966) works on the host block in
numpy exactly where the reference has it; ``use_pv`` selects the phase-vocoder cross-fade (gui.py:27-48, 1078-1083).

    blk = RealtimeBlock(rvc, samplerate=48000, block_time=0.16, crossfade_time=0.05, extra_time=2.5)
    out = blk.process(indata)            # np.float32 [block_frame] -> np.float32 [block_frame]
"""
from __future__ import annotations

import os

import numpy as np
import torch

from rvc_b200 import engine

from .realtime_tail import RealtimeTail
from .resample import Resample
from .torchgate import TorchGate


def _rms_frames(y: np.ndarray, frame_length: int, hop_length: int) -> np.ndarray:
    """librosa.feature.rms(y=y, frame_length=, hop_length=) (center=True, zero padding) -> [1, n]"""
    pad = frame_length // 2
    yp = np.pad(y.astype(np.float32), (pad, pad), mode="constant")
    n = 1 + (len(yp) - frame_length) // hop_length
    idx = np.arange(frame_length)[None, :] + hop_length * np.arange(n)[:, None]
    return np.sqrt(np.mean(np.abs(yp[idx]) ** 2, axis=1, keepdims=True)).T.astype(np.float32)


def response_gate(indata: np.ndarray, rms_buffer: np.ndarray, zc: int, threhold: float) -> np.ndarray:
    """gui.py:951-966: zero the zc-segments whose RMS is under the response threshold (host numpy, as in the reference).  ``rms_buffer``
    (the last 4 zc samples of the previous call) is updated in place; the returned block is 2 zc longer than the input."""
    indata = np.append(rms_buffer, indata)
    rms = _rms_frames(indata, 4 * zc, zc)[:, 2:]
    rms_buffer[:] = indata[-4 * zc:]
    indata = indata[2 * zc - zc // 2:]
    db = 20.0 * np.log10(np.maximum(1e-5, rms))                  # librosa.amplitude_to_db(rms, ref=1.0): amin 1e-5, top_db 80
    db = np.maximum(db, db.max() - 80.0)
    quiet = db[0] < threhold
    for i in range(quiet.shape[0]):
        if quiet[i]:
            indata[i * zc: (i + 1) * zc] = 0
    return indata[zc // 2:]


class RealtimeBlock:
    def __init__(self, rvc, samplerate: int = 48000, block_time: float = 0.25, crossfade_time: float = 0.05, extra_time: float = 2.5,
                 I_noise_reduce: bool = False, O_noise_reduce: bool = False, rms_mix_rate: float = 1.0, threhold: float = -60.0,
                 f0method: str = "rmvpe", device="cuda:0", use_pv: bool = False):
        self.rvc, self.f0method = rvc, f0method
        self.device = torch.device(device if "cuda" in str(device) else "cuda:0")
        self.samplerate = samplerate
        self.I_noise_reduce, self.O_noise_reduce, self.rms_mix_rate, self.threhold = I_noise_reduce, O_noise_reduce, rms_mix_rate, threhold
        t = self.tail = RealtimeTail(samplerate, block_time, crossfade_time, extra_time, self.device, use_pv)
        self.zc, self.block_frame, self.block_frame_16k = t.zc, t.block_frame, t.block_frame_16k
        self.sola_buffer_frame, self.extra_frame = t.sola_buffer_frame, t.extra_frame
        self.skip_head, self.return_length = t.skip_head, t.return_length
        z = lambda n: torch.zeros(n, device=self.device, dtype=torch.float32)                     # noqa: E731
        self.input_wav = z(t.input_frames)                                                        # gui.py:816-823
        self.input_wav_denoise = z(t.input_frames)
        self.input_wav_res = z(t.input_frames_16k)
        self.rms_buffer = np.zeros(4 * self.zc, dtype="float32")
        self.nr_buffer = z(self.sola_buffer_frame)
        self.output_buffer = z(t.input_frames)
        self.fade_in_window = torch.sin(0.5 * np.pi * torch.linspace(0.0, 1.0, steps=self.sola_buffer_frame, device=self.device,
                                                                     dtype=torch.float32)) ** 2
        self.fade_out_window = 1 - self.fade_in_window
        self.resampler = Resample(orig_freq=samplerate, new_freq=16000, dtype=torch.float32).to(self.device)
        self.resampler2 = (Resample(orig_freq=rvc.tgt_sr, new_freq=samplerate, dtype=torch.float32).to(self.device)
                           if rvc.tgt_sr != samplerate else None)
        self.tg = TorchGate(sr=samplerate, n_fft=4 * self.zc, prop_decrease=0.9).to(self.device)  # gui.py:869-871
        self._graphs = {}
        self._host_in = {}
        self._host_out = torch.empty(self.block_frame, dtype=torch.float32).pin_memory()

    # ------------------------------------------------------------------------------------------------------------------
    def _shift(self, buf: torch.Tensor, n: int):
        buf[:-n] = buf[n:].clone()

    @torch.no_grad()
    def _body(self, blk: torch.Tensor) -> torch.Tensor:
        """blk f32[n] (device, n = block_frame, or block_frame + 2 zc behind the threshold gate) -> f32[block_frame] (device).
        No host synchronisation: eager or captured."""
        zc, n = self.zc, blk.shape[0]
        self._shift(self.input_wav, self.block_frame)                                             # gui.py:967-972
        self.input_wav[-n:] = blk
        self._shift(self.input_wav_res, self.block_frame_16k)
        if self.I_noise_reduce:                                                                   # gui.py:974-993
            self._shift(self.input_wav_denoise, self.block_frame)
            x = self.input_wav[-self.sola_buffer_frame - self.block_frame:]
            x = self.tg(x.unsqueeze(0), self.input_wav.unsqueeze(0)).squeeze(0)
            x[: self.sola_buffer_frame] *= self.fade_in_window
            x[: self.sola_buffer_frame] += self.nr_buffer * self.fade_out_window
            self.input_wav_denoise[-self.block_frame:] = x[: self.block_frame]
            self.nr_buffer[:] = x[self.block_frame:]
            self.input_wav_res[-self.block_frame_16k - 160:] = self.resampler(self.input_wav_denoise[-self.block_frame - 2 * zc:])[160:]
        else:
            self.input_wav_res[-160 * (n // zc + 1):] = self.resampler(self.input_wav[-n - 2 * zc:])[160:]
        infer_wav = self.rvc._infer_body(self.input_wav_res, self.block_frame_16k, self.skip_head, self.return_length, self.f0method, 1.0)
        if self.resampler2 is not None:                                                           # gui.py:1008-1009
            infer_wav = self.resampler2(infer_wav)
        if self.O_noise_reduce:                                                                   # gui.py:1015-1023
            self._shift(self.output_buffer, self.block_frame)
            self.output_buffer[-self.block_frame:] = infer_wav[-self.block_frame:]
            infer_wav = self.tg(infer_wav.unsqueeze(0), self.output_buffer.unsqueeze(0)).squeeze(0)
        src = None
        if self.rms_mix_rate < 1:
            src = (self.input_wav_denoise if self.I_noise_reduce else self.input_wav)[self.extra_frame:]
        return self.tail.process(infer_wav.contiguous(), src, self.rms_mix_rate, want_offset=True)

    # ------------------------------------------------------------------------------------------------------------------
    def _gate(self, indata: np.ndarray) -> np.ndarray:
        return response_gate(indata, self.rms_buffer, self.zc, self.threhold)

    @torch.no_grad()
    def process(self, indata: np.ndarray) -> np.ndarray:
        """One block of the callback: mono float32 ``indata`` [block_frame] (host) -> float32 [block_frame] (host)."""
        indata = np.asarray(indata, dtype=np.float32).reshape(-1)
        if indata.shape[0] != self.block_frame:
            raise ValueError(f"expected a block of {self.block_frame} samples, got {indata.shape[0]}")
        if self.threhold > -60:
            indata = self._gate(indata.copy())
        n = indata.shape[0]
        hin = self._host_in.get(n)
        if hin is None:
            hin = self._host_in[n] = torch.empty(n, dtype=torch.float32).pin_memory()
        hin.numpy()[:] = indata
        key = (n, self.I_noise_reduce, self.O_noise_reduce, bool(self.tail.use_pv), float(self.rms_mix_rate), self.f0method, float(self.rvc.f0_up_key),
               float(self.rvc.formant_shift), float(self.rvc.index_rate), id(getattr(self.rvc, "index", None)))
        ent = self._graphs.get(key)
        use_graphs = os.environ.get("RVCB_GRAPHS", "1") != "0" and isinstance(self.f0method, str)
        if use_graphs and ent is not None and "graph" not in ent and not ent.get("failed"):
            try:
                ent["x"] = torch.empty(n, device=self.device, dtype=torch.float32)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    ent["out"] = self._body(ent["x"])
                ent["graph"] = g
            except Exception:
                ent.clear()
                ent["failed"] = True
                torch.cuda.synchronize()
        if not use_graphs or ent is None or "graph" not in ent:
            if use_graphs and ent is None:
                if len(self._graphs) >= 4:
                    self._graphs.pop(next(iter(self._graphs)))
                self._graphs[key] = {}
            out = self._body(hin.to(self.device, non_blocking=True))
        else:
            ent["x"].copy_(hin, non_blocking=True)
            ent["graph"].replay()
            out = ent["out"]
        self._host_out.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self._host_out.numpy().copy()

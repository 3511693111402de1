"""Oracle: the realtime engine ``infer.lib.rtrvc.RVC.infer`` and the per-block tail of gui.py's audio callback, fp32 CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Pinned where the reference can be executed here: ``phase_vocoder``, the SOLA step and the
callback's input side (rings, TorchGate, cross-fade, resampling) are bit-equal to the reference's own statements run from its source
(tests/golden/make_golden.py: phase_vocoder(), callback_pieces(), rtrvc_glue(); tests/test_oracle_golden.py); ``OracleRVC.infer`` is
bit-equal to the reference's own ``RVC.infer`` run from its source on duck-typed components (what reaches ``net_g.infer`` and the pitch
ring, three consecutive blocks).  The envelope mix (librosa) is a restatement.

Reference sites restated:
  RVC.__init__ state (pitch ring of 1024 frames)       infer/lib/rtrvc.py:63-66
  RVC.infer                                            infer/lib/rtrvc.py:134-260
      last HuBERT frame duplicated                        :163
      retrieval on frames >= skip_head // 2 only, skipped when any index is -1   :167-187
      RMVPE on the last f0_extractor_frame samples, pitch ring roll + ``pitch[3:-1]`` write   :199-217
      x2 nearest upsample, protect mix, net_g.infer(skip_head, return_length, return_length2)   :219-250
  gui.py audio callback tail (per block, on ``infer_wav``):
      volume-envelope mix  (librosa.feature.rms(frame 4*zc, hop zc), align_corners interpolation)   gui.py:1024-1056
      SOLA offset search + cross-fade + buffer update                                             gui.py:1057-1087
  fade windows sin^2                                   gui.py:841-855
  whole device side of the audio callback (``OracleCallback``)                                  gui.py:783-871, 940-1090
      response-threshold gate on the host block (librosa rms, zero the quiet zc-segments)          gui.py:951-966
      input ring shifts, input TorchGate + cross-fade with nr_buffer, resample to 16 kHz          gui.py:967-1000
      RVC.infer, resampler2, output TorchGate against output_buffer                               gui.py:1002-1023
librosa.feature.rms is restated (centred frames, zero padding) as in oracle/pipeline.py; TorchGate is oracle/torchgate.py (pinned
to the reference class); the resamplers are torchaudio.transforms.Resample itself (the library the reference calls).  The
phase-vocoder cross-fade (use_pv, gui.py:27-48) is restated operation for operation (``phase_vocoder``).
The formant-shift resampling branch (rtrvc.py:251-260, torchaudio Resample) is outside the oracle: tests use formant = 0.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import hubert as OH, ivf as OI, rmvpe as ORM, synth as OS
from .pipeline import rms_frames


class OracleRVC:
    def __init__(self, hubert_w, rmvpe_w, synth_w, synth_config, index: Optional[OI.IVFFlat], index_rate: float, key: float = 0,
                 version: str = "v2", noise_seed: int = 0):
        self.hw, self.rw, self.sw, self.cfg = hubert_w, rmvpe_w, synth_w, synth_config
        self.index, self.index_rate, self.f0_up_key, self.version = index, index_rate, key, version
        self.big_npy = index.reconstruct_n(0, index.ntotal) if index is not None else None
        self.window = 160
        self.tgt_sr = synth_config[-1] if not isinstance(synth_config[-1], str) else 48000
        self.cache_pitch = torch.zeros(1024, dtype=torch.long)
        self.cache_pitchf = torch.zeros(1024, dtype=torch.float32)
        self.gen = torch.Generator().manual_seed(noise_seed)
        self.taps = []

    @torch.no_grad()
    def infer(self, input_wav: np.ndarray, block_frame_16k: int, skip_head: int, return_length: int, protect: float = 1.0):
        wav = torch.from_numpy(np.asarray(input_wav, dtype=np.float32))
        feats = OH.extract_features(self.hw, wav.view(1, -1), 9 if self.version == "v1" else 12)
        if self.version == "v1":
            feats = OH.final_proj(self.hw, feats)
        feats = torch.cat((feats, feats[:, -1:, :]), 1)                                     # rtrvc.py:163
        tap = {"feats_hubert": feats.clone()}
        if self.index is not None and self.index_rate > 0:
            npy = feats[0][skip_head // 2:].numpy()
            score, ix = self.index.search(npy, k=8)
            tap["ix"] = ix
            if (ix >= 0).all():                                                              # rtrvc.py:173
                npy = OI.blend(npy, score, ix, self.big_npy, self.index_rate)
                feats[0][skip_head // 2:] = torch.from_numpy(npy)
        p_len = wav.shape[0] // self.window
        return_length2 = return_length                                                       # formant shift 0
        f0_extractor_frame = block_frame_16k + 800
        f0_extractor_frame = 5120 * ((f0_extractor_frame - 1) // 5120 + 1) - self.window     # rtrvc.py:199-203
        pitch, pitchf = ORM.calculate(self.rw, wav[-f0_extractor_frame:].numpy(), None, self.f0_up_key)
        pitch, pitchf = torch.from_numpy(pitch.astype(np.int64)), torch.from_numpy(pitchf.astype(np.float32))
        shift = block_frame_16k // self.window
        self.cache_pitch[:-shift] = self.cache_pitch[shift:].clone()
        self.cache_pitchf[:-shift] = self.cache_pitchf[shift:].clone()
        self.cache_pitch[4 - pitch.shape[0]:] = pitch[3:-1]
        self.cache_pitchf[4 - pitch.shape[0]:] = pitchf[3:-1]
        cache_pitch = self.cache_pitch[None, -p_len:]
        cache_pitchf = self.cache_pitchf[None, -p_len:] * return_length2 / return_length
        feats = F.interpolate(feats.permute(0, 2, 1), scale_factor=2).permute(0, 2, 1)[:, :p_len, :]
        upp = self.tgt_sr // 100
        flow_head = max(skip_head - 24, 0)
        n1 = torch.randn(1, self.cfg[2], p_len - flow_head, generator=self.gen)
        n2 = torch.randn(1, return_length * upp, 1, generator=self.gen)
        tap.update(phone=feats.clone(), noise=(n1, n2), pitch=cache_pitch.clone(), pitchf=cache_pitchf.clone())
        self.taps.append(tap)
        out = OS.synth_infer(self.sw, self.cfg, feats, torch.tensor([p_len]), torch.tensor([0]), cache_pitch, cache_pitchf, n1, n2,
                             skip_head=skip_head, return_length=return_length, return_length2=return_length2)
        return out[0, 0].numpy()


def fade_windows(n: int):
    """gui.py:841-855"""
    fade_in = torch.sin(0.5 * np.pi * torch.linspace(0.0, 1.0, steps=n, dtype=torch.float32)) ** 2
    return fade_in, 1 - fade_in


def envelope_mix(infer_wav: torch.Tensor, input_wav: torch.Tensor, zc: int, rms_mix_rate: float) -> torch.Tensor:
    """gui.py:1024-1056 (returns a new tensor; the reference scales ``infer_wav`` in place)."""
    n = infer_wav.shape[0]
    rms1 = torch.from_numpy(rms_frames(input_wav[:n].numpy(), 4 * zc, zc))
    rms1 = F.interpolate(rms1.unsqueeze(0), size=n + 1, mode="linear", align_corners=True)[0, 0, :-1]
    rms2 = torch.from_numpy(rms_frames(infer_wav.numpy(), 4 * zc, zc))
    rms2 = F.interpolate(rms2.unsqueeze(0), size=n + 1, mode="linear", align_corners=True)[0, 0, :-1]
    rms2 = torch.max(rms2, torch.zeros_like(rms2) + 1e-3)
    return infer_wav * torch.pow(rms1 / rms2, torch.tensor(1 - rms_mix_rate))


def phase_vocoder(a: torch.Tensor, b: torch.Tensor, fade_out: torch.Tensor, fade_in: torch.Tensor) -> torch.Tensor:
    """gui.py:27-48, operation for operation (the dtype of a / b decides the precision: float32 in the callback)."""
    window = torch.sqrt(fade_out * fade_in)
    fa = torch.fft.rfft(a * window)
    fb = torch.fft.rfft(b * window)
    absab = torch.abs(fa) + torch.abs(fb)
    n = a.shape[0]
    if n % 2 == 0:
        absab[1:-1] *= 2
    else:
        absab[1:] *= 2
    phia = torch.angle(fa)
    phib = torch.angle(fb)
    deltaphase = phib - phia
    deltaphase = deltaphase - 2 * np.pi * torch.floor(deltaphase / 2 / np.pi + 0.5)
    w = 2 * np.pi * torch.arange(n // 2 + 1).to(a) + deltaphase
    t = torch.arange(n).unsqueeze(-1).to(a) / n
    return a * (fade_out ** 2) + b * (fade_in ** 2) + torch.sum(absab * torch.cos(w * t + phia), -1) * window / n


class SolaTail:
    """State + per-block step of gui.py:1057-1087 (SOLA from DDSP-SVC); ``use_pv`` = the phase-vocoder branch (:1078-1083)."""

    def __init__(self, block_frame: int, sola_buffer_frame: int, sola_search_frame: int, use_pv: bool = False):
        self.use_pv = use_pv
        self.block_frame, self.sola_buffer_frame, self.sola_search_frame = block_frame, sola_buffer_frame, sola_search_frame
        self.sola_buffer = torch.zeros(sola_buffer_frame)
        self.fade_in, self.fade_out = fade_windows(sola_buffer_frame)

    def step(self, infer_wav: torch.Tensor):
        infer_wav = infer_wav.clone()
        conv_input = infer_wav[None, None, : self.sola_buffer_frame + self.sola_search_frame]
        cor_nom = F.conv1d(conv_input, self.sola_buffer[None, None, :])
        cor_den = torch.sqrt(F.conv1d(conv_input ** 2, torch.ones(1, 1, self.sola_buffer_frame)) + 1e-8)
        sola_offset = int(torch.argmax(cor_nom[0, 0] / cor_den[0, 0]))
        infer_wav = infer_wav[sola_offset:]
        if self.use_pv:
            infer_wav[: self.sola_buffer_frame] = phase_vocoder(self.sola_buffer, infer_wav[: self.sola_buffer_frame], self.fade_out, self.fade_in)
        else:
            infer_wav[: self.sola_buffer_frame] *= self.fade_in
            infer_wav[: self.sola_buffer_frame] += self.sola_buffer * self.fade_out
        self.sola_buffer[:] = infer_wav[self.block_frame: self.block_frame + self.sola_buffer_frame]
        return infer_wav[: self.block_frame].clone(), sola_offset


class OracleCallback:
    """gui.py:783-871 (block geometry + state) and :940-1090 (one block), for ``function == "vc"``, mono input, use_pv off.
    ``rvc``: an OracleRVC whose tgt_sr may differ from ``samplerate`` (then resampler2 runs)."""

    def __init__(self, rvc: OracleRVC, samplerate: int = 48000, block_time: float = 0.25, crossfade_time: float = 0.05,
                 extra_time: float = 2.5, I_noise_reduce: bool = False, O_noise_reduce: bool = False, rms_mix_rate: float = 1.0,
                 threhold: float = -60.0):
        import torchaudio.transforms as tat
        from . import torchgate as OT
        self.rvc, self.sr = rvc, samplerate
        self.I_noise_reduce, self.O_noise_reduce, self.rms_mix_rate, self.threhold = I_noise_reduce, O_noise_reduce, rms_mix_rate, threhold
        zc = self.zc = samplerate // 100
        self.block_frame = int(np.round(block_time * samplerate / zc)) * zc
        self.block_frame_16k = 160 * self.block_frame // zc
        self.crossfade_frame = int(np.round(crossfade_time * samplerate / zc)) * zc
        self.sola_buffer_frame = min(self.crossfade_frame, 4 * zc)
        self.sola_search_frame = zc
        self.extra_frame = int(np.round(extra_time * samplerate / zc)) * zc
        self.input_wav = torch.zeros(self.extra_frame + self.crossfade_frame + self.sola_search_frame + self.block_frame)
        self.input_wav_denoise = self.input_wav.clone()
        self.input_wav_res = torch.zeros(160 * self.input_wav.shape[0] // zc)
        self.rms_buffer = np.zeros(4 * zc, dtype="float32")
        self.nr_buffer = torch.zeros(self.sola_buffer_frame)
        self.output_buffer = self.input_wav.clone()
        self.skip_head = self.extra_frame // zc
        self.return_length = (self.block_frame + self.sola_buffer_frame + self.sola_search_frame) // zc
        self.fade_in_window, self.fade_out_window = fade_windows(self.sola_buffer_frame)
        self.resampler = tat.Resample(orig_freq=samplerate, new_freq=16000, dtype=torch.float32)
        self.resampler2 = tat.Resample(orig_freq=rvc.tgt_sr, new_freq=samplerate, dtype=torch.float32) if rvc.tgt_sr != samplerate else None
        self.tg = lambda x, xn: OT.torchgate(x, xn, samplerate, 4 * zc, prop_decrease=0.9)           # gui.py:869-871
        self.tail = SolaTail(self.block_frame, self.sola_buffer_frame, self.sola_search_frame)
        self.last = {}

    @torch.no_grad()
    def block(self, indata: np.ndarray) -> np.ndarray:
        zc = self.zc
        indata = np.asarray(indata, dtype=np.float32).copy()
        if self.threhold > -60:                                                                        # gui.py:951-966
            indata = np.append(self.rms_buffer, indata)
            rms = rms_frames(indata, 4 * zc, zc)[:, 2:]
            self.rms_buffer[:] = indata[-4 * zc:]
            indata = indata[2 * zc - zc // 2:]
            db = 20.0 * np.log10(np.maximum(1e-5, rms))                                                # librosa.amplitude_to_db(ref=1.0):
            db = np.maximum(db, db.max() - 80.0)                                                       # amin 1e-5, top_db 80
            db_threhold = db[0] < self.threhold
            for i in range(db_threhold.shape[0]):
                if db_threhold[i]:
                    indata[i * zc: (i + 1) * zc] = 0
            indata = indata[zc // 2:]
        self.input_wav[: -self.block_frame] = self.input_wav[self.block_frame:].clone()
        self.input_wav[-indata.shape[0]:] = torch.from_numpy(indata)
        self.input_wav_res[: -self.block_frame_16k] = self.input_wav_res[self.block_frame_16k:].clone()
        if self.I_noise_reduce:                                                                        # gui.py:974-993
            self.input_wav_denoise[: -self.block_frame] = self.input_wav_denoise[self.block_frame:].clone()
            input_wav = self.input_wav[-self.sola_buffer_frame - self.block_frame:]
            input_wav = self.tg(input_wav.unsqueeze(0), self.input_wav.unsqueeze(0)).squeeze(0)
            input_wav[: self.sola_buffer_frame] *= self.fade_in_window
            input_wav[: self.sola_buffer_frame] += self.nr_buffer * self.fade_out_window
            self.input_wav_denoise[-self.block_frame:] = input_wav[: self.block_frame]
            self.nr_buffer[:] = input_wav[self.block_frame:]
            self.input_wav_res[-self.block_frame_16k - 160:] = self.resampler(self.input_wav_denoise[-self.block_frame - 2 * zc:])[160:]
        else:
            self.input_wav_res[-160 * (indata.shape[0] // zc + 1):] = self.resampler(self.input_wav[-indata.shape[0] - 2 * zc:])[160:]
        infer_wav = torch.from_numpy(self.rvc.infer(self.input_wav_res.numpy(), self.block_frame_16k, self.skip_head, self.return_length))
        if self.resampler2 is not None:
            infer_wav = self.resampler2(infer_wav)
        if self.O_noise_reduce:                                                                        # gui.py:1015-1023
            self.output_buffer[: -self.block_frame] = self.output_buffer[self.block_frame:].clone()
            self.output_buffer[-self.block_frame:] = infer_wav[-self.block_frame:]
            infer_wav = self.tg(infer_wav.unsqueeze(0), self.output_buffer.unsqueeze(0)).squeeze(0)
        if self.rms_mix_rate < 1:
            src = self.input_wav_denoise if self.I_noise_reduce else self.input_wav
            infer_wav = envelope_mix(infer_wav, src[self.extra_frame:], zc, self.rms_mix_rate)
        self.last = {"infer_wav": infer_wav.clone(), "input_wav_res": self.input_wav_res.clone()}
        out, off = self.tail.step(infer_wav)
        self.last["offset"] = off
        return out.numpy()

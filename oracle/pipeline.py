"""Oracle: ``Pipeline.vc`` / ``Pipeline.pipeline`` control flow over the oracle stages.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Reference sites restated:
  bh, ah (Butterworth-5 HP 48 Hz)   infer/modules/vc/pipeline.py:23
  change_rms                        infer/modules/vc/pipeline.py:26-45 (librosa.feature.rms restated:
                                    centred, zero padded frames; librosa is not installed)
  Pipeline.__init__                 infer/modules/vc/pipeline.py:49-74
  Pipeline.vc                       infer/modules/vc/pipeline.py:76-184
  Pipeline.pipeline                 infer/modules/vc/pipeline.py:186-366
Noise for the synthesizer is drawn from an explicit ``torch.Generator`` so the
product can be fed the very same tensors.

Pinned: ``vc`` and ``pipeline`` are bit-equal to the reference's own methods executed from their source on duck-typed components
(tests/golden/make_golden.py: vc_glue(), pipeline_flow(); tests/test_oracle_golden.py) -- except the faiss / librosa branches, which
cannot run in the build container.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch
import torch.nn.functional as F
from scipy import signal

from . import hubert as OH, ivf as OI, rmvpe as ORM, synth as OS

bh, ah = signal.butter(N=5, Wn=48, btype="high", fs=16000)


def rms_frames(y: np.ndarray, frame_length: int, hop_length: int) -> np.ndarray:
    """librosa.feature.rms(y=y, frame_length, hop_length) (center=True, zero pad) -> [1, n]."""
    pad = frame_length // 2
    yp = np.pad(y.astype(np.float32), (pad, pad), mode="constant")
    n = 1 + (len(yp) - frame_length) // hop_length
    idx = np.arange(frame_length)[None, :] + hop_length * np.arange(n)[:, None]
    power = np.mean(np.abs(yp[idx]) ** 2, axis=1, keepdims=True)
    return np.sqrt(power).T.astype(np.float32)


def change_rms(data1, sr1, data2, sr2, rate):
    rms1 = torch.from_numpy(rms_frames(data1, sr1 // 2 * 2, sr1 // 2))
    rms2 = torch.from_numpy(rms_frames(data2, sr2 // 2 * 2, sr2 // 2))
    rms1 = F.interpolate(rms1.unsqueeze(0), size=data2.shape[0], mode="linear").squeeze()
    rms2 = F.interpolate(rms2.unsqueeze(0), size=data2.shape[0], mode="linear").squeeze()
    rms2 = torch.max(rms2, torch.zeros_like(rms2) + 1e-6)
    data2 = data2 * (torch.pow(rms1, torch.tensor(1 - rate)) * torch.pow(rms2, torch.tensor(rate - 1))).numpy()
    return data2


class OraclePipeline:
    def __init__(self, tgt_sr: int, x_pad: int, x_query: int, x_center: int, x_max: int,
                 hubert_w, rmvpe_w, synth_w, synth_config, noise_seed: int = 0):
        self.sr, self.window = 16000, 160
        self.x_pad = x_pad
        self.t_pad = self.sr * x_pad
        self.t_pad_tgt = tgt_sr * x_pad
        self.t_pad2 = self.t_pad * 2
        self.t_query = self.sr * x_query
        self.t_center = self.sr * x_center
        self.t_max = self.sr * x_max
        self.hw, self.rw, self.sw, self.cfg = hubert_w, rmvpe_w, synth_w, synth_config
        self.gen = torch.Generator().manual_seed(noise_seed)
        self.taps: List[dict] = []

    def draw_noise(self, T: int, upp: int):
        n1 = torch.randn(1, self.cfg[2], T, generator=self.gen)
        n2 = torch.randn(1, T * upp, 1, generator=self.gen)
        return n1, n2

    def vc(self, sid, audio0, pitch, pitchf, index, big_npy, index_rate, version, protect, noise=None):
        feats = OH.extract_features(self.hw, torch.from_numpy(audio0).float().view(1, -1),
                                    9 if version == "v1" else 12)
        if version == "v1":
            feats = OH.final_proj(self.hw, feats)
        feats0 = feats.clone() if (protect < 0.5 and pitch is not None) else None
        tap = {"feats_hubert": feats.clone()}
        if index is not None and big_npy is not None and index_rate != 0:
            npy = feats[0].numpy()
            score, ix = index.search(npy, k=8)
            tap["score"], tap["ix"] = score, ix
            feats = torch.from_numpy(OI.blend(npy, score, ix, big_npy, index_rate)).unsqueeze(0)
        feats = F.interpolate(feats.permute(0, 2, 1), scale_factor=2).permute(0, 2, 1)
        if feats0 is not None:
            feats0 = F.interpolate(feats0.permute(0, 2, 1), scale_factor=2).permute(0, 2, 1)
        p_len = audio0.shape[0] // self.window
        if feats.shape[1] < p_len:
            p_len = feats.shape[1]
            if pitch is not None:
                pitch, pitchf = pitch[:, :p_len], pitchf[:, :p_len]
        if feats0 is not None:
            pitchff = pitchf.clone()
            pitchff[pitchf > 0] = 1
            pitchff[pitchf < 1] = protect
            pitchff = pitchff.unsqueeze(-1)
            feats = feats * pitchff + feats0 * (1 - pitchff)
        tap["phone"] = feats.clone()
        upp = self.cfg[-1] // 100 if not isinstance(self.cfg[-1], str) else int(self.cfg[-1][:-1]) * 10
        if noise is None:
            noise = self.draw_noise(feats.shape[1], upp)
        tap["noise"] = noise
        # net_g.infer receives all feats frames; masks use p_len (synthesizers.py:186-189)
        audio1 = OS.synth_infer(self.sw, self.cfg, feats, torch.tensor([p_len]), sid, pitch, pitchf,
                                noise[0], noise[1])[0, 0].numpy()
        self.taps.append(tap)
        return audio1

    def pipeline(self, sid: int, audio: np.ndarray, f0_up_key, f0_method, index, index_rate, if_f0,
                 tgt_sr, resample_sr, rms_mix_rate, version, protect):
        big_npy = index.reconstruct_n(0, index.ntotal) if index is not None else None
        audio = signal.filtfilt(bh, ah, audio)
        audio_pad = np.pad(audio, (self.window // 2, self.window // 2), mode="reflect")
        opt_ts = []
        if audio_pad.shape[0] > self.t_max:
            audio_sum = np.zeros_like(audio)
            for i in range(self.window):
                audio_sum += np.abs(audio_pad[i: i - self.window])
            for t in range(self.t_center, audio.shape[0], self.t_center):
                seg = audio_sum[t - self.t_query: t + self.t_query]
                opt_ts.append(t - self.t_query + np.where(seg == seg.min())[0][0])
        s = 0
        audio_opt = []
        t = None
        audio_pad = np.pad(audio, (self.t_pad, self.t_pad), mode="reflect")
        p_len = audio_pad.shape[0] // self.window
        sid_t = torch.tensor(sid).unsqueeze(0).long()
        pitch = pitchf = None
        if if_f0:
            if if_f0 == 1:
                pitch, pitchf = ORM.calculate(self.rw, audio_pad.astype(np.float32), p_len, f0_up_key)
            else:
                pitch, pitchf = f0_method
            pitch = torch.tensor(pitch[:p_len]).unsqueeze(0).long()
            pitchf = torch.tensor(pitchf[:p_len].astype(np.float32)).unsqueeze(0).float()
            self.pitch, self.pitchf = pitch, pitchf
        W_ = self.window
        for t in opt_ts:
            t = t // W_ * W_
            audio_opt.append(self.vc(sid_t, audio_pad[s: t + self.t_pad2 + W_].astype(np.float32),
                                     pitch[:, s // W_: (t + self.t_pad2) // W_] if if_f0 else None,
                                     pitchf[:, s // W_: (t + self.t_pad2) // W_] if if_f0 else None,
                                     index, big_npy, index_rate, version, protect)[self.t_pad_tgt: -self.t_pad_tgt])
            s = t
        audio_opt.append(self.vc(sid_t, audio_pad[t:].astype(np.float32),
                                 (pitch[:, t // W_:] if t is not None else pitch) if if_f0 else None,
                                 (pitchf[:, t // W_:] if t is not None else pitchf) if if_f0 else None,
                                 index, big_npy, index_rate, version, protect)[self.t_pad_tgt: -self.t_pad_tgt])
        audio_opt = np.concatenate(audio_opt)
        if rms_mix_rate != 1:
            audio_opt = change_rms(audio, 16000, audio_opt, tgt_sr, rms_mix_rate)
        if tgt_sr != resample_sr >= 16000:
            # pipeline.py:351-354 calls librosa.resample (soxr_hq), which is not installed in the build container and whose
            # arithmetic lives in a C library outside /root/reference: PARITY UNPINNED for this optional branch.  The stand-in is
            # torchaudio's windowed-sinc resampler (sinc_interp_hann defaults), the library the reference itself uses for its other
            # resamplers (gui.py:851-866); the product runs the same table as a CUDA kernel.
            import torchaudio
            audio_opt = torchaudio.functional.resample(torch.from_numpy(audio_opt.astype(np.float32)), tgt_sr, resample_sr).numpy()
        audio_max = np.abs(audio_opt).max() / 0.99
        max_int16 = 32768
        if audio_max > 1:
            max_int16 /= audio_max
        return audio_opt * max_int16

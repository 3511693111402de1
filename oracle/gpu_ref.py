"""Reference GPU path: the oracle's functional PyTorch modules run EAGER on the GPU (cuDNN / cuBLAS kernels), fp16 or fp32,
with the reference's host round trips preserved.  SURVEY.md §8d(ii): "the same PyTorch modules on the same B200 in fp16 and
fp32, eager" -- the bar every sm_100a kernel of the product has to beat.

TEST / BENCH INFRASTRUCTURE (see oracle/__init__.py): used by ``bench.py --impl reference_gpu``, ``tools/ref_gpu_stages.py``
and ``tools/parity_report.py`` only.  Nothing of the product imports it.

Control flow restated (single-chunk utterance, as config #2):
  host filtfilt + reflect pad                    infer/modules/vc/pipeline.py:221,241
  RMVPE: mel fp32 on the device, cast to half before the log, E2E net, salience -> HOST, numpy decode   rvc/f0/rmvpe.py:96-164
  f0 post-processing on the host (numba in the reference)                                             rvc/f0/gen.py:10-41
  HuBERT on the device, features -> HOST for the faiss search, blend in numpy, back to the device      pipeline.py:102-138
  x2 nearest upsample, protect, net_g.infer on the device, waveform -> HOST                            pipeline.py:140-174
  change_rms + int16 scaling on the host                                                              pipeline.py:349-360
faiss is absent: the CPU search stand-in is a torch sgemm L2 search over the probed list (same semantics as the oracle IVF,
BLAS arithmetic), so that the retrieval leg costs what a BLAS-backed CPU library costs rather than the oracle's serial
lane-order emulation.
"""
from __future__ import annotations

import time
from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F
from scipy import signal

from . import hubert as OH, ivf as OI, pipeline as OP, rmvpe as ORM, synth as OS


def _to(w: Dict[str, torch.Tensor], device, dtype):
    return {k: (v.to(device=device, dtype=dtype) if v.is_floating_point() else v.to(device)) for k, v in w.items()}


class BlasIVF:
    """IVF-Flat nprobe=1 search with BLAS distances (||q||^2 - 2 q.c + ||c||^2), CPU: what faiss-cpu does for nq >= 20."""

    def __init__(self, idx: OI.IVFFlat):
        self.idx = idx
        self.c = torch.from_numpy(idx.centroids)
        self.cn = (self.c * self.c).sum(1)
        self.v = torch.from_numpy(idx.vectors)
        self.lists = [torch.from_numpy(np.asarray(idx.list_ids[idx.list_off[l]:idx.list_off[l + 1]])) for l in range(len(idx.list_off) - 1)]

    def search(self, x: np.ndarray, k: int = 8):
        q = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        coarse = (self.cn[None, :] - 2.0 * (q @ self.c.t())).argmin(1)
        D = np.full((q.shape[0], k), 3.4028235e38, dtype=np.float32)
        I = np.full((q.shape[0], k), -1, dtype=np.int64)
        order = torch.argsort(coarse)
        cs = coarse[order]
        bounds = torch.nonzero(cs[1:] != cs[:-1]).flatten() + 1
        starts = [0] + bounds.tolist() + [len(cs)]
        for a, b in zip(starts[:-1], starts[1:]):
            ids = self.lists[int(cs[a])]
            rows = order[a:b]
            if len(ids) == 0:
                continue
            v = self.v[ids]
            d = (q[rows] * q[rows]).sum(1, keepdim=True) - 2.0 * (q[rows] @ v.t()) + (v * v).sum(1)[None, :]
            kk = min(k, len(ids))
            dv, di = torch.topk(d, kk, dim=1, largest=False)
            D[rows.numpy()[:, None], np.arange(kk)[None, :]] = dv.numpy()
            I[rows.numpy()[:, None], np.arange(kk)[None, :]] = ids[di].numpy()
        return D, I


class GpuReference:
    def __init__(self, hw, rw, sw, cfg, idx: OI.IVFFlat, device="cuda:0", half=True, x_pad=3):
        self.dev = torch.device(device)
        self.dt = torch.float16 if half else torch.float32
        self.half = half
        self.hw, self.rw, self.sw = _to(hw, self.dev, self.dt), _to(rw, self.dev, self.dt), _to(sw, self.dev, self.dt)
        self.cfg = cfg
        self.idx = BlasIVF(idx) if idx is not None else None
        self.big = idx.reconstruct_n(0, idx.ntotal) if idx is not None else None
        self.t_pad = 16000 * x_pad
        self.t_pad_tgt = int(cfg[-1]) * x_pad if not isinstance(cfg[-1], str) else 48000 * x_pad
        self.upp = (cfg[-1] if not isinstance(cfg[-1], str) else 48000) // 100
        self.gen = torch.Generator(device=self.dev).manual_seed(0)
        self.stage_ms: Dict[str, float] = {}

    # ---- stages (each returns device tensors; host round trips are where the reference has them) ----
    def f0(self, audio_pad: np.ndarray, p_len: int, key: float = 0):
        wav = torch.from_numpy(audio_pad).float().to(self.dev)[None]
        mel = ORM.log_mel(wav) if not self.half else self._log_mel_half(wav)
        hidden = ORM.mel2hidden(self.rw, mel)[0].float().cpu().numpy()           # rmvpe.py:108-112 (host round trip)
        f0 = ORM.interpolate_f0(ORM.resize_f0(ORM.decode(hidden, 0.03), p_len))
        return ORM.post_process(f0, key)

    def _log_mel_half(self, wav):
        win = torch.hann_window(ORM.N_FFT, device=wav.device)
        fft = torch.stft(wav, n_fft=ORM.N_FFT, hop_length=ORM.HOP, win_length=ORM.N_FFT, window=win, center=True, return_complex=True)
        mag = torch.sqrt(fft.real.pow(2) + fft.imag.pow(2))
        mel = torch.matmul(torch.from_numpy(ORM.mel_filterbank()).to(mag.device), mag).half()          # mel.py:68-70
        return torch.log(torch.clamp(mel, min=1e-5))

    def hubert(self, audio0: np.ndarray):
        src = torch.from_numpy(audio0).to(self.dev, self.dt).view(1, -1)                                # pipeline.py:78-88
        return OH.extract_features(self.hw, src, 12)

    def retrieve(self, feats: torch.Tensor, index_rate: float):
        npy = feats[0].float().cpu().numpy()                                                             # pipeline.py:118
        score, ix = self.idx.search(npy, 8)
        out = OI.blend(npy, score, ix, self.big, index_rate)
        return torch.from_numpy(out).unsqueeze(0).to(self.dev, self.dt), ix                              # pipeline.py:135-138

    def synth(self, feats, p_len, pitch, pitchf, noise=None, taps=None):
        T = feats.shape[1]
        if noise is None:
            n1 = torch.randn(1, self.cfg[2], T, device=self.dev, dtype=self.dt, generator=self.gen)
            n2 = torch.randn(1, T * self.upp, 1, device=self.dev, dtype=self.dt, generator=self.gen)
        else:
            n1, n2 = noise[0].to(self.dev, self.dt), noise[1].to(self.dev, self.dt)
        return OS.synth_infer(self.sw, self.cfg, feats, torch.tensor([p_len], device=self.dev), torch.tensor([0], device=self.dev),
                              pitch, pitchf, n1, n2, taps=taps)

    # ---- the whole single-chunk utterance, host numpy in -> host float (int16 range) out ----
    @torch.no_grad()
    def convert(self, audio: np.ndarray, index_rate=0.75, protect=0.33, rms_mix_rate=0.25, noise=None, pitch_override=None, timed=False):
        ev = []

        def mark(name):
            if timed:
                torch.cuda.synchronize(self.dev)
                ev.append((name, time.perf_counter()))
        mark("start")
        a = signal.filtfilt(OP.bh, OP.ah, audio)
        audio_pad = np.pad(a, (self.t_pad, self.t_pad), mode="reflect").astype(np.float32)
        p_len = audio_pad.shape[0] // 160
        mark("host_pre")
        if pitch_override is None:
            pitch, pitchf = self.f0(audio_pad, p_len)
        else:
            pitch, pitchf = pitch_override
        pitch = torch.tensor(np.asarray(pitch)[:p_len], device=self.dev).unsqueeze(0).long()
        pitchf = torch.tensor(np.asarray(pitchf)[:p_len].astype(np.float32), device=self.dev).unsqueeze(0)
        mark("rmvpe_f0")
        feats = self.hubert(audio_pad)
        mark("hubert")
        feats0 = feats.clone() if protect < 0.5 else None
        ix = None
        if self.idx is not None and index_rate != 0:
            feats, ix = self.retrieve(feats, index_rate)
        mark("retrieval_cpu")
        feats = F.interpolate(feats.permute(0, 2, 1), scale_factor=2).permute(0, 2, 1)
        if feats0 is not None:
            feats0 = F.interpolate(feats0.permute(0, 2, 1), scale_factor=2).permute(0, 2, 1)
        pl = min(audio_pad.shape[0] // 160, feats.shape[1])
        pitch, pitchf = pitch[:, :pl], pitchf[:, :pl]
        if feats0 is not None:
            pf = pitchf.clone()
            pf[pitchf > 0] = 1
            pf[pitchf < 1] = protect
            pf = pf.unsqueeze(-1).to(self.dt)
            feats = feats * pf + feats0 * (1 - pf)
        wav = self.synth(feats, pl, pitch, pitchf, noise)[0, 0].float().cpu().numpy()                     # pipeline.py:172-174
        mark("synth")
        wav = wav[self.t_pad_tgt: -self.t_pad_tgt]
        if rms_mix_rate != 1:
            wav = OP.change_rms(a, 16000, wav, self.upp * 100, rms_mix_rate)
        amax = np.abs(wav).max() / 0.99
        out = wav * (32768 / amax if amax > 1 else 32768)
        mark("host_post")
        if timed:
            self.stage_ms = {n: (t - ev[i][1]) * 1e3 for i, (n, t) in enumerate(ev[1:])}
        self.last_ix = ix
        return out

"""Drop-in for ``infer.lib.rtrvc.RVC`` (infer/lib/rtrvc.py:19-274): the realtime engine gui.py drives
(one rolling window per block).  Same constructor arguments, attributes (tgt_sr, if_f0, version), setters
and ``infer(input_wav, block_frame_16k, skip_head, return_length, f0method, protect)``."""
from __future__ import annotations

import os
from pathlib import Path
from typing import Optional, Union

import numpy as np
import torch

from rvc.f0 import Generator
from rvc.synthesizer import get_synthesizer, load_synthesizer
from rvc_b200 import engine, faiss_io
from infer.modules.vc.utils import load_hubert


class RVC:
    def __init__(self, key, formant, pth_path, index_path, index_rate, n_cpu: int = os.cpu_count(), device: str = "cuda:0",
                 use_jit: bool = False, is_half: bool = False, is_dml: bool = False, hubert_model=None, rmvpe_state_dict=None):
        self.device = torch.device(device if "cuda" in str(device) else "cuda:0")
        self.f0_up_key = key
        self.formant_shift = formant
        self.sr = 16000
        self.window = 160
        self.n_cpu = n_cpu
        self.is_half = is_half
        self.index_path = index_path
        self.index_rate = index_rate
        if index_rate > 0:
            self._load_index()
        self.pth_path = pth_path
        self.cache_pitch = torch.zeros(1024, device=self.device, dtype=torch.long)
        self.cache_pitchf = torch.zeros(1024, device=self.device, dtype=torch.float32)
        self.resample_kernel = {}
        self._side = torch.cuda.Stream(device=self.device)
        self._graphs = {}
        self.f0_gen = Generator(rmvpe_state_dict or Path(os.environ.get("rmvpe_root", "assets/rmvpe")), is_half, 0, self.device,
                                self.window, self.sr)
        self.hubert = hubert_model if hubert_model is not None else load_hubert(self.device, is_half)
        if isinstance(pth_path, dict):
            self.net_g, cpt = get_synthesizer(pth_path, self.device)
        else:
            self.net_g, cpt = load_synthesizer(pth_path, self.device)
        self.tgt_sr = cpt["config"][-1]
        self.if_f0 = cpt.get("f0", 1)
        self.version = cpt.get("version", "v1")

    def _load_index(self):
        if isinstance(self.index_path, engine.Index):
            self.index = self.index_path
        else:
            self.index = engine.Index.from_oracle_layout(faiss_io.read_index(self.index_path), self.device.index or 0)
        self.big_npy = self.index.vectors

    def set_key(self, new_key): self.f0_up_key = new_key
    def set_formant(self, new_formant): self.formant_shift = new_formant

    def set_index_rate(self, new_index_rate):
        if new_index_rate > 0 and self.index_rate <= 0:
            self._load_index()
        self.index_rate = new_index_rate

    @torch.no_grad()
    def infer(self, input_wav: torch.Tensor, block_frame_16k: int, skip_head: int, return_length: int, f0method: Union[tuple, str],
              protect: float = 1.0) -> torch.Tensor:
        """rtrvc.py:131-246.  A realtime session calls this with the same shapes block after block: from the second block with
        the same (shapes, settings) on, the whole block -- both streams, the pitch-cache roll, the noise draws -- is one CUDA
        graph replay (RVCB_GRAPHS=0 turns that off)."""
        wav_dev = input_wav.float().to(self.device)
        if wav_dev.dim() == 2:
            wav_dev = wav_dev.mean(-1)
        key = None
        if isinstance(f0method, str) and os.environ.get("RVCB_GRAPHS", "1") != "0" and getattr(self.net_g, "accepts_host_scalars", False):
            key = (int(wav_dev.shape[0]), int(block_frame_16k), int(skip_head), int(return_length), f0method, float(protect),
                   float(self.f0_up_key), float(self.formant_shift), float(self.index_rate), id(getattr(self, "index", None)))
        ent = self._graphs.get(key) if key is not None else None
        if ent is not None and "graph" not in ent and not ent.get("failed"):
            try:
                ent["x"] = torch.empty_like(wav_dev)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    ent["out"] = self._infer_body(ent["x"], block_frame_16k, skip_head, return_length, f0method, protect)
                ent["graph"] = g
                ent["refs"] = (self.net_g, self.hubert, getattr(self, "index", None))     # device pointers are baked into the graph
            except Exception:                           # capture is an optimisation only: stay eager for this key
                ent.clear()
                ent["failed"] = True
                torch.cuda.synchronize()
        if ent is None or "graph" not in ent:
            if key is not None and ent is None:
                if len(self._graphs) >= 4:
                    self._graphs.pop(next(iter(self._graphs)))
                self._graphs[key] = {}
            return self._infer_body(wav_dev, block_frame_16k, skip_head, return_length, f0method, protect)
        ent["x"].copy_(wav_dev, non_blocking=True)
        ent["graph"].replay()
        return ent["out"].clone()

    def _infer_body(self, wav_dev, block_frame_16k, skip_head, return_length, f0method, protect):
        try:
            return self._infer_body_impl(wav_dev, block_frame_16k, skip_head, return_length, f0method, protect)
        except Exception:
            engine.set_grid_cap(0)          # never leave the process-wide launch cap behind
            raise

    def _infer_body_impl(self, wav_dev: torch.Tensor, block_frame_16k: int, skip_head: int, return_length: int, f0method, protect: float):
        input_wav = wav_dev
        capturing = torch.cuda.is_current_stream_capturing()
        # content features + retrieval do not depend on f0: run them on a side stream while RMVPE runs on the main one
        cur = torch.cuda.current_stream()
        self._side.wait_stream(cur)
        # the offline pipeline caps the two front branches at half the SMs each (side by side); on the short realtime window the
        # cap measured slower (p50 3.64 vs 3.48 ms per block): RVCB_RT_FRONT_CAP opts in
        cap_prev = engine.set_grid_cap(int(os.environ.get("RVCB_RT_FRONT_CAP", "0")))
        with torch.cuda.stream(self._side):
            logits = self.hubert.extract_features(source=wav_dev.view(1, -1), padding_mask=None, output_layer=9 if self.version == "v1" else 12)
            feats = self.hubert.final_proj(logits[0]) if self.version == "v1" else logits[0]
            feats = torch.cat((feats, feats[:, -1:, :]), 1)[0]                       # rtrvc.py:163
            feats0 = feats.clone() if (protect < 0.5 and self.if_f0 == 1) else None
            try:
                if hasattr(self, "index") and self.index_rate > 0:
                    tail = feats[skip_head // 2:]
                    D, I = self.index.search_device(tail, 8)
                    # rtrvc.py:173 ``if (ix >= 0).all()``: selected on the device so the host never waits for HuBERT here
                    blended = self.index.blend_device(tail, D, I, self.index_rate)
                    feats[skip_head // 2:] = torch.where((I >= 0).all(), blended, tail)
            except Exception:
                pass
            feats_ready = torch.cuda.Event()
            feats_ready.record(self._side)
        p_len = input_wav.shape[0] // self.window
        factor = pow(2, self.formant_shift / 12)
        return_length2 = int(np.ceil(return_length * factor))
        cache_pitch = cache_pitchf = None
        pitch = pitchf = None
        if isinstance(f0method, tuple):
            pitch, pitchf = f0method
            pitch = torch.tensor(pitch, device=self.device).unsqueeze(0).long()
            pitchf = torch.tensor(pitchf, device=self.device).unsqueeze(0).float()
            cache_pitch, cache_pitchf = pitch[:, -p_len:], pitchf[:, -p_len:] * return_length2 / return_length
        elif self.if_f0 == 1:
            f0_extractor_frame = block_frame_16k + 800
            if f0method == "rmvpe":
                f0_extractor_frame = 5120 * ((f0_extractor_frame - 1) // 5120 + 1) - self.window
            wav_f0 = input_wav[-f0_extractor_frame:]
            if f0method == "rmvpe" and torch.is_tensor(wav_f0) and wav_f0.is_cuda:
                # RMVPE + f0 post-processing stay on the device: no host round trip inside the block
                pitch, pitchf = self.f0_gen.calculate_device(wav_f0, wav_f0.shape[0] // self.window, self.f0_up_key - self.formant_shift)
            else:
                c, f = self.f0_gen.calculate(wav_f0, None, self.f0_up_key - self.formant_shift, f0method, None)
                pitch = torch.from_numpy(c).long().to(self.device)
                pitchf = torch.from_numpy(np.asarray(f)).float().to(self.device)
            shift = block_frame_16k // self.window
            self.cache_pitch[:-shift] = self.cache_pitch[shift:].clone()
            self.cache_pitchf[:-shift] = self.cache_pitchf[shift:].clone()
            self.cache_pitch[4 - pitch.shape[0]:] = pitch[3:-1]
            self.cache_pitchf[4 - pitch.shape[0]:] = pitchf[3:-1]
            cache_pitch = self.cache_pitch[None, -p_len:]
            cache_pitchf = self.cache_pitchf[None, -p_len:] * return_length2 / return_length
        engine.set_grid_cap(cap_prev)
        cur.wait_event(feats_ready)
        if not capturing:
            feats.record_stream(cur)
            if feats0 is not None:
                feats0.record_stream(cur)
        use_protect = protect < 0.5 and pitch is not None and pitchf is not None and feats0 is not None
        pf = None
        if use_protect:
            # The reference builds the voiced / unvoiced mask from this block's fresh f0 (rtrvc.py:226-236), which covers only
            # the last f0_extractor_frame samples while the features cover the whole rolling window: with "rmvpe" its shapes do
            # not match and it raises (gui.py never passes protect < 0.5).  Here the mask comes from the pitch ring, which is
            # aligned with the window frame for frame (unscaled by the formant factor); a (pitch, pitchf) tuple is used as given.
            pf = torch.ones(p_len, device=self.device)
            src = pitchf.reshape(-1) if isinstance(f0method, tuple) else self.cache_pitchf[-p_len:]
            n = min(p_len, src.shape[0])
            pf[p_len - n:] = src[-n:]
        phone = engine.upsample_protect(feats, feats0 if use_protect else None, pf, p_len, protect if use_protect else 1.0)
        out = self.net_g.infer(phone.unsqueeze(0), torch.tensor([p_len]), torch.tensor([0]),     # host scalars: no stream sync
                               pitch=cache_pitch, pitchf=cache_pitchf, skip_head=skip_head, return_length=return_length,
                               return_length2=return_length2).squeeze(1).float()
        upp_res = int(np.floor(factor * self.tgt_sr // 100))
        if upp_res != self.tgt_sr // 100:
            from torchaudio.transforms import Resample
            if upp_res not in self.resample_kernel:
                self.resample_kernel[upp_res] = Resample(orig_freq=upp_res, new_freq=self.tgt_sr // 100, dtype=torch.float32).to(self.device)
            out = self.resample_kernel[upp_res](out[:, : return_length * upp_res])
        return out.squeeze()

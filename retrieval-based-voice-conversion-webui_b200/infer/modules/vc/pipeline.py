"""Drop-in for ``infer.modules.vc.pipeline.Pipeline`` (infer/modules/vc/pipeline.py:48-366).

Same constructor, ``vc`` and ``pipeline`` signatures, ``times`` accounting and return conventions; the
stages between the host DSP run on the B200 without the reference's four host round-trips per chunk
(pipeline.py:118, 135-138, 172-174; rmvpe.py:109): HuBERT features, IVF-Flat search + blend, x2 upsample
+ protect mix, f0 and the synthesizer all stay on the device; one D2H copy returns the chunk's waveform.
"""
from __future__ import annotations

import os
import threading
import traceback
import logging
from pathlib import Path
from time import time

import numpy as np
import torch
import torch.nn.functional as F
from scipy import signal

from rvc.f0 import Generator
from rvc_b200 import engine, faiss_io

logger = logging.getLogger(__name__)

# Launch-side work that touches process-wide state is serialised between the conversion lanes of VC.vc_multi (one thread each):
# torch's CUDA generator (the noise draws inside a captured graph register offset increments with it -- a capture in one thread while
# another thread replays or draws eagerly raises "Offset increment outside graph capture"), graph capture itself, and the library's
# grid cap.  Only the host-side ENQUEUE is under the lock (~0.1 ms for a graph replay); the utterances still overlap on the device.
_LAUNCH_LOCK = threading.RLock()

bh, ah = signal.butter(N=5, Wn=48, btype="high", fs=16000)     # pipeline.py:23
zi_h = signal.lfilter_zi(bh, ah)                                # filtfilt's edge state, a constant of the filter
sos_h, sos_zi_h = engine.highpass_sos_from_ba(bh, ah)            # the same filter as second-order sections (device path)


def _rms(y: np.ndarray, frame_length: int, hop_length: int) -> np.ndarray:
    """librosa.feature.rms (center=True, zero padding) -> [1, n_frames].  frame_length == 2*hop_length here
    (pipeline.py:28-31), so every frame is two adjacent hop blocks: O(N) block sums instead of a gather."""
    pad = frame_length // 2
    yp = np.pad(y.astype(np.float32), (pad, pad), mode="constant")
    n = 1 + (len(yp) - frame_length) // hop_length
    if frame_length == 2 * hop_length:
        nb = n + 1
        blocks = (yp[: nb * hop_length].astype(np.float64) ** 2).reshape(nb, hop_length).sum(axis=1)
        power = (blocks[:-1] + blocks[1:]) / frame_length
        return np.sqrt(power)[None, :].astype(np.float32)
    idx = np.arange(frame_length)[None, :] + hop_length * np.arange(n)[:, None]
    return np.sqrt(np.mean(np.abs(yp[idx]) ** 2, axis=1, keepdims=True)).T.astype(np.float32)


def change_rms(data1, sr1, data2, sr2, rate):     # pipeline.py:26-45
    rms1 = torch.from_numpy(_rms(data1, sr1 // 2 * 2, sr1 // 2))
    rms2 = torch.from_numpy(_rms(data2, sr2 // 2 * 2, sr2 // 2))
    rms1 = F.interpolate(rms1.unsqueeze(0), size=data2.shape[0], mode="linear").squeeze()
    rms2 = F.interpolate(rms2.unsqueeze(0), size=data2.shape[0], mode="linear").squeeze()
    rms2 = torch.max(rms2, torch.zeros_like(rms2) + 1e-6)
    data2 *= (torch.pow(rms1, torch.tensor(1 - rate)) * torch.pow(rms2, torch.tensor(rate - 1))).numpy()
    return data2


class PendingResult:
    """A device-resident result whose device-to-host copy (into pinned memory, on the stream that produced it) has been queued but not
    waited for: ``result()`` waits and returns the numpy array.  Lets one host thread keep several utterances in flight."""

    def __init__(self, pipe: "Pipeline", dev: torch.Tensor):
        n = dev.numel()
        buf = pipe._out_pinned.get(dev.dtype)
        if buf is None or buf.numel() < n:
            buf = pipe._out_pinned[dev.dtype] = torch.empty(max(n, 1 << 19), dtype=dev.dtype, pin_memory=True)
        self._host = buf[:n]
        self._host.copy_(dev.reshape(-1), non_blocking=True)
        self._ev = torch.cuda.Event()
        self._ev.record(torch.cuda.current_stream())

    def result(self) -> np.ndarray:
        self._ev.synchronize()
        return self._host.numpy().copy()          # the pinned buffer is reused by this lane's next utterance


class Pipeline(object):
    def __init__(self, tgt_sr, config):
        self.x_pad, self.x_query, self.x_center, self.x_max, self.is_half = (
            config.x_pad, config.x_query, config.x_center, config.x_max, config.is_half)
        self.sr = 16000
        self.window = 160
        self.t_pad = self.sr * self.x_pad
        self.t_pad_tgt = tgt_sr * self.x_pad
        self.t_pad2 = self.t_pad * 2
        self.t_query = self.sr * self.x_query
        self.t_center = self.sr * self.x_center
        self.t_max = self.sr * self.x_max
        self.device = torch.device(config.device if "cuda" in str(config.device) else "cuda:0")
        rmvpe_root = getattr(config, "rmvpe_state_dict", None) or Path(os.environ.get("rmvpe_root", "assets/rmvpe"))
        self.f0_gen = Generator(rmvpe_root, self.is_half, self.x_pad, self.device, self.window, self.sr)
        self._index_cache = {}
        self._side = torch.cuda.Stream(device=self.device)
        self._resamplers = {}          # (tgt_sr, resample_sr) -> device sinc table of the resample_sr branch
        self._out_pinned = {}          # dtype -> pinned result buffer of PendingResult (one utterance in flight per Pipeline)
        self._defer_d2h = False
        # RMVPE ends in the serial BiGRU and is the longer front branch: it gets a high-priority stream so that its CTAs are placed
        # first whenever SMs free up, and HuBERT + retrieval fill the rest (RVCB_F0_PRIO=0: RMVPE stays on the caller's stream)
        self._f0_stream = torch.cuda.Stream(device=self.device, priority=-1) if os.environ.get("RVCB_F0_PRIO", "1") != "0" else None
        self._prefetched = None
        self._pinned = None          # reusable pinned staging buffer for the H2D copy of the utterance
        self._graphs = {}            # (length, settings) -> captured CUDA graph of the device-resident utterance path

    # -----------------------------------------------------------------------------------------
    def _features(self, model, audio0, index, big_npy, index_rate, version):
        """HuBERT features (+ IVF-Flat blend) of one chunk: (blended [T_h, C], unblended [T_h, C])"""
        feats = audio0.float() if torch.is_tensor(audio0) else torch.from_numpy(np.ascontiguousarray(audio0, dtype=np.float32))
        if feats.dim() == 2:
            feats = feats.mean(-1)
        assert feats.dim() == 1, feats.dim()
        feats = feats.view(1, -1)
        with torch.no_grad():
            logits = model.extract_features(source=feats.to(self.device, non_blocking=True), padding_mask=None,   # all-False, pipeline.py:100
                                            output_layer=9 if version == "v1" else 12)
            feats = model.final_proj(logits[0]) if version == "v1" else logits[0]
        f0 = feats[0]
        f = f0
        if index is not None and big_npy is not None and index_rate != 0:
            if isinstance(index, engine.Index):
                D, I = index.search_device(f, 8)                       # IVF-Flat nprobe=1, k=8 on the device
                f = index.blend_device(f, D, I, index_rate)
            else:   # foreign index object (e.g. real faiss): the reference's host path, pipeline.py:118-138
                npy = f.cpu().numpy().astype("float32")
                try:
                    score, ix = index.search(npy, k=8)
                except Exception:
                    raise Exception("index mistatch")
                weight = np.square(1 / score)
                weight /= weight.sum(axis=1, keepdims=True)
                npy = np.sum(big_npy[ix] * np.expand_dims(weight, axis=2), axis=1)
                f = torch.from_numpy(npy.astype(np.float32)).to(self.device) * index_rate + (1 - index_rate) * f
        return f, f0

    def vc(self, model, net_g, sid, audio0, pitch, pitchf, times, index, big_npy, index_rate, version, protect):
        """Reference signature (pipeline.py:76): returns the chunk's waveform as a host float32 array."""
        return self._vc_dev(model, net_g, sid, audio0, pitch, pitchf, times, index, big_npy, index_rate, version, protect).cpu().numpy()

    def _vc_dev(self, model, net_g, sid, audio0, pitch, pitchf, times, index, big_npy, index_rate, version, protect, trim=False):
        """Device form of ``vc``.  trim=True returns the chunk WITHOUT its x_pad context, i.e. what the callers keep after
        ``audio1[t_pad_tgt : -t_pad_tgt]`` (pipeline.py:295): the synthesizer then runs its local stages (flow, decoder) only over
        the kept frames plus their receptive-field margins -- the kept samples are bit-identical (rvcb_synth_infer_keep)."""
        t0 = time()
        if self._prefetched is not None and self._prefetched[0] is audio0:
            _, f, f_raw, ev = self._prefetched          # computed on the side stream while RMVPE was running
            self._prefetched = None
            torch.cuda.current_stream().wait_event(ev)
            if not torch.cuda.is_current_stream_capturing():
                f.record_stream(torch.cuda.current_stream())
                f_raw.record_stream(torch.cuda.current_stream())
        else:
            f, f_raw = self._features(model, audio0, index, big_npy, index_rate, version)
        use_protect = protect < 0.5 and pitch is not None and pitchf is not None
        feats0 = f_raw if use_protect else None
        t1 = time()
        p_len = audio0.shape[0] // self.window
        if 2 * f.shape[0] < p_len:
            p_len = 2 * f.shape[0]
            if pitch is not None and pitchf is not None:
                pitch = pitch[:, :p_len]
                pitchf = pitchf[:, :p_len]
        # x2 nearest upsample + protect mix (pipeline.py:140-160) in one kernel; all 2*T_h frames go to net_g
        T2 = 2 * f.shape[0]
        pf_full = None
        if use_protect:
            pf_full = torch.ones(T2, device=self.device)
            pf_full[:pitchf.shape[1]] = pitchf[0, :T2]
        phone = engine.upsample_protect(f, feats0, pf_full, T2, protect if use_protect else 1.0)
        with torch.no_grad():
            T = phone.shape[0]
            if pitch is not None and pitch.shape[1] < T:      # masks use p_len; frames past it are ignored downstream
                phone = phone[: pitch.shape[1]]
                T = phone.shape[0]
            host_ok = getattr(net_g, "accepts_host_scalars", False)      # a device tensor here costs a stream sync (.item())
            lengths = torch.tensor([T]) if host_ok else torch.tensor([T], device=self.device)
            pad_f = self.t_pad_tgt // getattr(net_g, "upp", 0) if getattr(net_g, "upp", 0) else 0
            can_keep = (trim and host_ok and pad_f > 0 and T > 2 * pad_f and os.environ.get("RVCB_TRIM", "1") != "0"
                        and pad_f * net_g.upp == self.t_pad_tgt)
            if can_keep:
                audio1 = net_g.infer(phone.unsqueeze(0), lengths, sid, pitch=None if pitch is None else pitch[:, :T],
                                     pitchf=None if pitchf is None else pitchf[:, :T], keep_head=pad_f, keep_length=T - 2 * pad_f)[0, 0]
            else:
                audio1 = net_g.infer(phone.unsqueeze(0), lengths, sid, pitch=None if pitch is None else pitch[:, :T],
                                     pitchf=None if pitchf is None else pitchf[:, :T])[0, 0]
                if trim:
                    audio1 = audio1[self.t_pad_tgt: -self.t_pad_tgt]
        t2 = time()
        times[0] += t1 - t0
        times[2] += t2 - t1
        return audio1

    # -----------------------------------------------------------------------------------------
    def _load_index(self, file_index):
        """Replaces faiss.read_index + reconstruct_n on EVERY call (pipeline.py:213-215) with a cached
        device-resident index."""
        if isinstance(file_index, engine.Index):
            return file_index, file_index.vectors
        key = (file_index, os.path.getmtime(file_index))
        if key not in self._index_cache:
            layout = faiss_io.read_index(file_index)
            self._index_cache = {key: engine.Index.from_oracle_layout(layout, self.device.index or 0)}
        ix = self._index_cache[key]
        return ix, ix.vectors

    def _stage_h2d(self, audio: np.ndarray) -> torch.Tensor:
        """One pinned-memory H2D copy of the utterance (float32), asynchronous on the current stream.  Two staging buffers
        alternate, each guarded by the event of the last copy issued from it: back-to-back calls without a host sync never
        overwrite a buffer whose queued copy has not run yet."""
        n = int(audio.shape[0])
        if self._pinned is None or self._pinned[0][0].numel() < n:
            cap = max(n, 1 << 18)
            self._pinned = [[torch.empty(cap, dtype=torch.float32, pin_memory=True), None] for _ in range(2)]
            self._pin_i = 0
        self._pin_i ^= 1
        slot = self._pinned[self._pin_i]
        if slot[1] is not None:
            slot[1].synchronize()
        stage = slot[0][:n]
        stage.numpy()[:] = audio          # casts float64 -> float32 if needed
        x = stage.to(self.device, non_blocking=True)
        slot[1] = torch.cuda.Event()
        slot[1].record(torch.cuda.current_stream())
        return x

    def _epilogue_dev(self, out, tgt_sr, resample_sr, a16, rms_mix_rate):
        """pipeline.py:349-360 on the device: change_rms, the optional resample_sr branch, peak normalisation to the int16 range.
        The reference resamples with librosa (soxr), which is not installed here: the windowed-sinc resampler of torchaudio
        (``torchaudio.functional.resample`` defaults) stands in, as a CUDA kernel (rvcb_resample_sinc) -- not numerically equal to
        the reference on this optional branch."""
        out = out.contiguous()
        if not (tgt_sr != resample_sr >= 16000):
            return engine.post_mix(out, tgt_sr, a16, rms_mix_rate)
        if rms_mix_rate != 1:
            out = engine.rms_mix(out, tgt_sr, a16, rms_mix_rate)
        key = (int(tgt_sr), int(resample_sr))
        if key not in self._resamplers:
            k, width, up, down = engine.sinc_resample_kernel(tgt_sr, resample_sr)
            self._resamplers[key] = (k.to(self.device), width, up, down)
        out = engine.sinc_resample(out, *self._resamplers[key])
        return engine.post_mix(out, resample_sr, a16, 1.0)            # rate 1: no mix, peak scaling only

    def _dev_body(self, x, model, net_g, sid, times, f0_up_key, index, big_npy, index_rate, if_f0, tgt_sr, rms_mix_rate, version,
                  protect, as_int16, resample_sr=0):
        """Device-resident body of one single-chunk utterance: x f32[n] (device) -> waveform (device).  No host
        synchronisation anywhere, so it can run eagerly or be captured once into a CUDA graph and replayed."""
        capturing = torch.cuda.is_current_stream_capturing()
        a16 = engine.sosfiltfilt(sos_h, sos_zi_h, 3 * max(len(ah), len(bh)), x)
        audio_pad = engine.reflect_pad(a16, self.t_pad)
        p_len = audio_pad.numel() // self.window
        pitch = pitchf = None
        self._prefetched = None
        if if_f0 == 1:
            cur = torch.cuda.current_stream()
            fork = torch.cuda.Event()
            fork.record(cur)                       # the side stream depends on the padded audio only, not on RMVPE
            # the two branches are independent.  HuBERT + retrieval launch with their persistent grids capped at half the SMs
            # (RVCB_FRONT_CAP), RMVPE -- the longer branch, on the high-priority stream -- uncapped (RVCB_F0_CAP, 0 = no cap), so
            # that kernels of the two streams run side by side instead of taking turns at the whole chip.  Measured (whole step,
            # profiles/README.md r2k): both capped, no priority 7.56 ms; priority + RMVPE uncapped 7.43 ms; nothing capped 7.73 ms
            cap_prev = engine.set_grid_cap(int(os.environ.get("RVCB_F0_CAP", "0")))
            try:
                # f0 first: RMVPE has the fewer launches, so both branches are in flight sooner when launching eagerly
                if self._f0_stream is not None:
                    self._f0_stream.wait_event(fork)
                    with torch.cuda.stream(self._f0_stream):
                        pitch, pitchf = self.f0_gen.calculate_device(audio_pad, p_len, f0_up_key)
                        f0_done = torch.cuda.Event()
                        f0_done.record(self._f0_stream)
                    cur.wait_event(f0_done)
                    if not capturing:
                        audio_pad.record_stream(self._f0_stream)
                        pitch.record_stream(cur)
                        pitchf.record_stream(cur)
                else:
                    pitch, pitchf = self.f0_gen.calculate_device(audio_pad, p_len, f0_up_key)
                pitch, pitchf = pitch.unsqueeze(0), pitchf.unsqueeze(0)
                engine.set_grid_cap(engine.front_branch_cap())
                self._side.wait_event(fork)
                with torch.cuda.stream(self._side):
                    f, f_raw = self._features(model, audio_pad, index, big_npy, index_rate, version)
                    ev = torch.cuda.Event()
                    ev.record(self._side)
            finally:
                engine.set_grid_cap(cap_prev)
            if not capturing:
                audio_pad.record_stream(self._side)
            self._prefetched = (audio_pad, f, f_raw, ev)
        out = self._vc_dev(model, net_g, sid, audio_pad, pitch, pitchf, times, index, big_npy, index_rate, version, protect, trim=True)
        out = self._epilogue_dev(out, tgt_sr, resample_sr, a16, rms_mix_rate)
        return engine.f32_to_i16(out) if as_int16 else out

    def _pipeline_single_dev(self, model, net_g, sid, audio, times, f0_up_key, index, big_npy, index_rate, if_f0, tgt_sr,
                             rms_mix_rate, version, protect, as_int16=False, resample_sr=0):
        """One-chunk utterance with NO host round trip between the input copy and the result copy: high-pass filtfilt
        (pipeline.py:221), reflect padding (:241), RMVPE, f0 post-processing (rvc/f0/gen.py:10-41), HuBERT + retrieval on a
        side stream, synthesizer, RMS mix + scaling (:349-360) and the int16 cast (modules.py:181) all run on the device.
        The second time the same (length, settings) comes in, the ~510 launches are captured into a CUDA graph and replayed
        from then on (RVCB_GRAPHS=0 turns that off).  Returns a device tensor (float32 in the int16 range, or int16)."""
        t0 = time()
        host_ok = getattr(net_g, "accepts_host_scalars", False)
        sid_t = torch.tensor(sid).unsqueeze(0).long() if host_ok else torch.tensor(sid, device=self.device).unsqueeze(0).long()
        args = (model, net_g, sid_t, times, f0_up_key, index, big_npy, index_rate, if_f0, tgt_sr, rms_mix_rate, version, protect, as_int16,
                int(resample_sr))
        x = self._stage_h2d(audio)
        key = None
        if host_ok and os.environ.get("RVCB_GRAPHS", "1") != "0" and (index is None or isinstance(index, engine.Index)):
            key = (int(audio.shape[0]), id(model), id(net_g), int(sid), float(f0_up_key), id(index), float(index_rate), int(if_f0),
                   int(tgt_sr), float(rms_mix_rate), str(version), float(protect), bool(as_int16), int(resample_sr))
        with _LAUNCH_LOCK:
            ent = self._graphs.get(key) if key is not None else None
            if ent is not None and "graph" not in ent and not ent.get("failed"):
                # second sighting: capture (arenas and kernels are warm from the first run)
                try:
                    ent["x"] = torch.empty_like(x)
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, capture_error_mode="thread_local"):       # other vc_multi lanes keep running
                        ent["out"] = self._dev_body(ent["x"], *args)
                    ent["graph"] = g
                except Exception:                           # capture is an optimisation only: keep launching eagerly for this key
                    logger.warning("CUDA graph capture failed, staying eager:\n%s", traceback.format_exc())
                    ent.clear()
                    ent["failed"] = True
                    torch.cuda.synchronize()
            if key is None or ent is None or "graph" not in ent:
                if key is not None and ent is None:
                    if len(self._graphs) >= 4:
                        self._graphs.pop(next(iter(self._graphs)))
                    self._graphs[key] = {}
                out = self._dev_body(x, *args)
            else:
                if "refs" not in ent:
                    ent["refs"] = (model, net_g, index, big_npy)     # the graph bakes their device pointers in: keep them alive (and
                                                                     # their id()s, which are part of the key, unrecyclable)
                ent["x"].copy_(x, non_blocking=True)
                ent["graph"].replay()
                out = ent["out"]
                times[2] += time() - t0          # launch time only: the replay is asynchronous like the eager path
        return out

    def pipeline(self, model, net_g, sid, audio, times, f0_up_key, f0_method, file_index, index_rate, if_f0, filter_radius, tgt_sr,
                 resample_sr, rms_mix_rate, version, protect, f0_file=None):
        index = big_npy = None
        if index_rate != 0 and (isinstance(file_index, engine.Index) or (file_index != "" and os.path.exists(file_index))):
            try:
                index, big_npy = self._load_index(file_index)
            except Exception:
                traceback.print_exc()
                index = big_npy = None
        single = audio.shape[0] + 2 * (self.window // 2) <= self.t_max          # pipeline.py:224: no silence-point chunking
        if (single and if_f0 in (0, 1) and (if_f0 == 0 or f0_method == "rmvpe") and not hasattr(f0_file, "name")
                and not getattr(self, "_force_host", False)):
            as_i16 = getattr(self, "_want_int16", False)
            out = self._pipeline_single_dev(model, net_g, sid, audio, times, f0_up_key, index, big_npy, index_rate, if_f0, tgt_sr,
                                            rms_mix_rate, version, protect, as_int16=as_i16, resample_sr=resample_sr)
            if getattr(self, "_defer_d2h", False):
                return PendingResult(self, out)          # VC.vc_multi: the copy back is queued, the caller collects it later
            return out.cpu().numpy()
        with _LAUNCH_LOCK:          # the reference's control flow with eager launches (noise draws): one lane at a time
            return self._pipeline_host_flow(model, net_g, sid, audio, times, f0_up_key, f0_method, index, big_npy, index_rate, if_f0,
                                            filter_radius, tgt_sr, resample_sr, rms_mix_rate, version, protect, f0_file)

    def _pipeline_host_flow(self, model, net_g, sid, audio, times, f0_up_key, f0_method, index, big_npy, index_rate, if_f0, filter_radius,
                            tgt_sr, resample_sr, rms_mix_rate, version, protect, f0_file):
        """pipeline.py:221-366 with the reference's own control flow (silence-point chunking, f0 files, given f0): host DSP
        prologue, device compute per chunk, device epilogue."""
        audio = engine.host_filtfilt(bh, ah, zi_h, audio)       # == signal.filtfilt(bh, ah, audio), bit for bit
        audio_pad = np.pad(audio, (self.window // 2, self.window // 2), mode="reflect")
        opt_ts = []
        if audio_pad.shape[0] > self.t_max:
            # pipeline.py:224-227 sums 160 shifted copies; a cumulative sum gives the same window sums
            cs = np.concatenate([[0.0], np.cumsum(np.abs(audio_pad))])
            audio_sum = cs[self.window: self.window + audio.shape[0]] - cs[: audio.shape[0]]
            for t in range(self.t_center, audio.shape[0], self.t_center):
                seg = audio_sum[t - self.t_query: t + self.t_query]
                opt_ts.append(t - self.t_query + int(np.argmin(seg)))
        s = 0
        audio_opt = []
        t = None
        t1 = time()
        audio_pad = np.pad(audio, (self.t_pad, self.t_pad), mode="reflect")
        p_len = audio_pad.shape[0] // self.window
        inp_f0 = None
        if hasattr(f0_file, "name"):
            try:
                with open(f0_file.name, "r") as f:
                    raw_lines = f.read()
                    if len(raw_lines) > 0:
                        inp_f0 = np.array([[float(i) for i in line.split(",")] for line in raw_lines.strip("\n").split("\n")], dtype="float32")
            except Exception:
                traceback.print_exc()
        sid = torch.tensor(sid, device=self.device).unsqueeze(0).long()
        pitch, pitchf = None, None
        self._prefetched = None
        if not opt_ts and if_f0 == 1:
            # one chunk: content features + retrieval do not depend on f0 -> run them on a side stream under RMVPE
            self._side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._side):
                f, f_raw = self._features(model, audio_pad, index, big_npy, index_rate, version)
                ev = torch.cuda.Event()
                ev.record(self._side)
            self._prefetched = (audio_pad, f, f_raw, ev)
        if if_f0:
            if if_f0 == 1:
                pitch, pitchf = self.f0_gen.calculate(audio_pad, p_len, f0_up_key, f0_method, filter_radius, inp_f0)
            elif if_f0 == 2:
                pitch, pitchf = f0_method
            pitch = pitch[:p_len]
            pitchf = pitchf[:p_len].astype(np.float32)
            pitch = torch.tensor(pitch, device=self.device).unsqueeze(0).long()
            pitchf = torch.tensor(pitchf, device=self.device).unsqueeze(0).float()
        t2 = time()
        times[1] += t2 - t1
        W = self.window
        for t in opt_ts:
            t = t // W * W
            audio_opt.append(self._vc_dev(model, net_g, sid, audio_pad[s: t + self.t_pad2 + W],
                                     pitch[:, s // W: (t + self.t_pad2) // W] if if_f0 else None,
                                     pitchf[:, s // W: (t + self.t_pad2) // W] if if_f0 else None,
                                     times, index, big_npy, index_rate, version, protect, trim=True))
            s = t
        audio_opt.append(self._vc_dev(model, net_g, sid, audio_pad if t is None else audio_pad[t:],
                                 (pitch[:, t // W:] if t is not None else pitch) if if_f0 else None,
                                 (pitchf[:, t // W:] if t is not None else pitchf) if if_f0 else None,
                                 times, index, big_npy, index_rate, version, protect, trim=True))
        audio_dev = audio_opt[0] if len(audio_opt) == 1 else torch.cat(audio_opt)
        # RMS-envelope mix (+ resample_sr branch) + peak normalisation on the device (pipeline.py:349-360), then ONE D2H copy
        a16 = torch.from_numpy(np.ascontiguousarray(audio, dtype=np.float32)).to(self.device, non_blocking=True)
        return self._epilogue_dev(audio_dev, tgt_sr, resample_sr, a16, rms_mix_rate).cpu().numpy()

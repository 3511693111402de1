"""Drop-in for ``infer.modules.vc.modules.VC`` (infer/modules/vc/modules.py:18-266): same method names,
arguments, Gradio-shaped returns and error convention (exceptions inside vc_single become the info string,
modules.py:196-199).  Organised as small helpers around the three public methods; under torchrun ``vc_multi``
strides the file list over ranks (one process / stream per GPU, no collective)."""
from __future__ import annotations

import logging
import os
import traceback
from typing import Any, Dict, Iterator, List, Optional, Tuple

import numpy as np
import torch

from infer.lib.audio import load_audio, save_audio
from rvc.synthesizer import get_synthesizer, load_synthesizer
from .pipeline import Pipeline
from .utils import get_index_path_from_model, load_hubert

logger = logging.getLogger(__name__)


def _gr(**fields) -> Dict[str, Any]:
    """A Gradio component update, as the WebUI expects it (``{"__type__": "update", ...}``)."""
    fields["__type__"] = "update"
    return fields


def _unquote(path: str) -> str:
    """What the reference does with pasted paths: strip blanks, quotes and newlines at both ends (modules.py:150-158, 216-217)."""
    return path.strip(" ").strip('"').strip("\n").strip('"').strip(" ")


class VC:
    def __init__(self, config):
        self.config = config
        self.hubert_model = None
        self.pipeline: Optional[Pipeline] = None
        self.net_g = None
        self.cpt: Optional[dict] = None
        self.n_spk = self.tgt_sr = self.version = self.if_f0 = None
        self._loaded_sid = None
        self._lanes: List["VC"] = []          # extra conversion lanes of vc_multi (own handles / streams / graphs each)
        self._index_clones: Dict[tuple, tuple] = {}     # (id(index object), lane) -> (index object, its per-lane clone)

    # ------------------------------------------------------------------ model selection (modules.py:32-117)
    def _protect_updates(self, extra: tuple) -> Tuple[Dict[str, Any], Dict[str, Any]]:
        """The two "protect" sliders: hidden for no-f0 models; they keep the caller's values when there are any."""
        show = self.if_f0 != 0
        keep = show and bool(extra)
        return (_gr(visible=show, value=extra[0] if keep else 0.5), _gr(visible=show, value=extra[1] if keep else 0.33))

    def _unload(self) -> None:
        if self.hubert_model is None:
            return
        logger.info("Clean model cache")
        self.hubert_model = self.net_g = self.n_spk = self.tgt_sr = None
        self._lanes.clear()
        torch.cuda.empty_cache()

    def _load(self, sid) -> str:
        """Build the synthesizer container and the pipeline for ``sid`` (a file under $weight_root, or the dict a .pth holds)."""
        self._loaded_sid = sid
        self._lanes.clear()
        self._index_clones.clear()
        if isinstance(sid, dict):
            self.net_g, self.cpt = get_synthesizer(sid, self.config.device)
            name = sid.get("name", "in-memory")
        else:
            person = f'{os.getenv("weight_root")}/{sid}'
            logger.info(f"Loading: {person}")
            self.net_g, self.cpt = load_synthesizer(person, self.config.device)
            name = sid
        cfg = self.cpt["config"]
        cfg[-3] = self.cpt["weight"]["emb_g.weight"].shape[0]        # speakers actually present in the checkpoint
        self.tgt_sr, self.n_spk = cfg[-1], cfg[-3]
        self.if_f0 = self.cpt.get("f0", 1)
        self.version = self.cpt.get("version", "v1")
        self.net_g = self.net_g.half() if self.config.is_half else self.net_g.float()
        self.pipeline = Pipeline(self.tgt_sr, self.config)
        return name

    def get_vc(self, sid, *to_return_protect):
        logger.info("Get sid: " + str(sid))
        protect0, protect1 = self._protect_updates(to_return_protect)      # from the PREVIOUS model's f0 flag, like the reference
        if sid == "" or sid == []:
            self._unload()
            if not to_return_protect:
                return _gr(visible=True, maximum=0)
            return (_gr(visible=False), protect0, protect1, _gr(value=to_return_protect[2]), _gr(value=to_return_protect[3]),
                    _gr(value=""))
        name = self._load(sid)
        speakers = _gr(visible=True, maximum=self.n_spk)
        if not to_return_protect:
            return speakers
        index = _gr(value=get_index_path_from_model(name))
        return speakers, protect0, protect1, index, index, _gr(value=self.cpt.get("info", ""))

    # ------------------------------------------------------------------ one utterance (modules.py:119-199)
    @staticmethod
    def _pick_index(file_index, file_index2):
        """Text box first (``trained`` -> ``added``), then the dropdown, else no index; a device-resident Index passes through."""
        if file_index is not None and not isinstance(file_index, str) and not hasattr(file_index, "name"):
            return file_index
        if file_index:
            if hasattr(file_index, "name"):
                file_index = str(file_index.name)
            return _unquote(file_index).replace("trained", "added")
        return file_index2 if file_index2 else ""

    def _convert(self, sid, audio: np.ndarray, f0_up_key: int, f0_file, f0_method, file_index, index_rate, filter_radius, resample_sr,
                 rms_mix_rate, protect):
        peak = np.abs(audio).max() / 0.95
        if peak > 1:
            np.divide(audio, peak, audio)
        if self.hubert_model is None:
            self.hubert_model = load_hubert(self.config.device, self.config.is_half)
        times = [0, 0, 0]
        self.pipeline._want_int16 = True        # the device path casts like .astype(np.int16) before its single D2H copy
        try:
            wav = self.pipeline.pipeline(self.hubert_model, self.net_g, sid, audio, times, f0_up_key, f0_method, file_index, index_rate,
                                         self.if_f0, filter_radius, self.tgt_sr, resample_sr, rms_mix_rate, self.version, protect, f0_file)
        finally:
            self.pipeline._want_int16 = False
        out_sr = resample_sr if self.tgt_sr != resample_sr >= 16000 else self.tgt_sr
        if hasattr(wav, "result"):                # PendingResult (vc_multi lanes): the copy back is still in flight
            return out_sr, wav, times
        return out_sr, wav.astype(np.int16, copy=False), times

    def vc_single(self, sid, input_audio_path, f0_up_key, f0_file, f0_method, file_index, file_index2, index_rate, filter_radius,
                  resample_sr, rms_mix_rate, protect):
        if input_audio_path is None:
            return "You need to upload an audio", None
        if hasattr(input_audio_path, "name"):
            input_audio_path = str(input_audio_path.name)
        f0_up_key = int(f0_up_key)
        try:
            audio = input_audio_path if isinstance(input_audio_path, np.ndarray) else load_audio(input_audio_path, 16000)
            audio = np.array(audio, dtype=np.float32)
            file_index = self._pick_index(file_index, file_index2)
            out_sr, wav, times = self._convert(sid, audio, f0_up_key, f0_file, f0_method, file_index, index_rate, filter_radius,
                                               resample_sr, rms_mix_rate, protect)
            used = not isinstance(file_index, str) or os.path.exists(file_index)
            index_info = "Index: %s." % file_index if used else "Index not used."
            return "Success.\n%s\nTime: npy: %.2fs, f0: %.2fs, infer: %.2fs." % (index_info, *times), (out_sr, wav)
        except Exception as e:
            logger.warning(traceback.format_exc())
            return str(e), None

    # ------------------------------------------------------------------ a folder of utterances (modules.py:201-266)
    @staticmethod
    def _inputs(dir_path: str, uploads) -> List[str]:
        try:
            if dir_path != "":
                return [os.path.join(dir_path, name) for name in os.listdir(dir_path)]
        except Exception:
            traceback.print_exc()
        return [p.name for p in uploads]

    def _lane(self, i: int) -> "VC":
        """Lane 0 is this object; lane i > 0 is a private copy of the loaded models (own synthesizer container, HuBERT handle,
        Pipeline with its RMVPE, streams and captured graphs): the library's workspaces are per handle, so utterances that are in
        flight at the same time must not share handles."""
        if i == 0:
            return self
        while len(self._lanes) < i:
            lane = VC(self.config)
            lane._load(self._loaded_sid)
            if self.hubert_model is None:
                self.hubert_model = load_hubert(self.config.device, self.config.is_half)
            lane.hubert_model = self.hubert_model.clone() if hasattr(self.hubert_model, "clone") else load_hubert(self.config.device,
                                                                                                               self.config.is_half)
            self._lanes.append(lane)
        return self._lanes[i - 1]

    def _index_clone(self, index, lane_id: int):
        """Per-lane handle of a device-resident Index object, created once (the captured graphs are keyed by the handle's identity)."""
        key = (id(index), lane_id)
        if key not in self._index_clones:
            if len(self._index_clones) >= 8:
                self._index_clones.clear()
            self._index_clones[key] = (index, index.clone())          # the original is kept alive so its id() cannot be recycled
        return self._index_clones[key][1]

    def _submit(self, lane: "VC", stream, sid, path, f0_up_key, f0_method, file_index, index_rate, filter_radius, resample_sr, rms_mix_rate,
                protect):
        """First half of ``vc_single`` on ``lane`` / ``stream``: decode, enqueue the whole utterance and its copy back, do NOT wait.
        Returns ``finish() -> (info, opt)`` (the second half: wait, format the info string)."""
        try:
            audio = np.array(load_audio(path, 16000), dtype=np.float32)
            lane.pipeline._defer_d2h = True
            try:
                with torch.cuda.stream(stream):
                    out_sr, wav, times = lane._convert(sid, audio, int(f0_up_key), None, f0_method, file_index, index_rate, filter_radius,
                                                       resample_sr, rms_mix_rate, protect)
            finally:
                lane.pipeline._defer_d2h = False
        except Exception as e:
            logger.warning(traceback.format_exc())
            msg = str(e)
            return lambda: (msg, None)

        def finish():
            try:
                w = wav.result() if hasattr(wav, "result") else wav
                used = not isinstance(file_index, str) or os.path.exists(file_index)
                index_info = "Index: %s." % file_index if used else "Index not used."
                return ("Success.\n%s\nTime: npy: %.2fs, f0: %.2fs, infer: %.2fs." % (index_info, *times),
                        (out_sr, w.astype(np.int16, copy=False)))
            except Exception as e:
                logger.warning(traceback.format_exc())
                return str(e), None
        return finish

    def vc_multi(self, sid, dir_path, opt_root, paths, f0_up_key, f0_method, file_index, file_index2, index_rate, filter_radius,
                 resample_sr, rms_mix_rate, protect, format1) -> Iterator[str]:
        """modules.py:201-266.  The reference converts the files one after the other.  With RVCB_LANES > 1 (default 1 = the serial
        loop; opt-in) this host thread keeps that many utterances in flight on the GPU: file i is enqueued on lane i mod L (own handle set,
        stream and captured graph) and collected, in input order, just before its lane is needed again -- no host threads, so nothing
        contends for the GIL and the noise draws stay single-threaded.  Results equal the serial loop's; the log lines keep the input
        order.  Under torchrun the list is additionally strided over ranks (one process per GPU)."""
        try:
            dir_path, opt_root = _unquote(dir_path), _unquote(opt_root)
            os.makedirs(opt_root, exist_ok=True)
            todo = self._inputs(dir_path, paths)
            # under torchrun the list is strided over ranks exactly like extract_feature_print.py:110
            rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
            todo = todo[rank::world]
            n_lanes = max(1, min(int(os.environ.get("RVCB_LANES", "1")), len(todo)))
            log: List[str] = []

            def record(path: str, info: str, opt) -> str:
                if "Success" in info:
                    try:
                        out_sr, wav = opt
                        save_audio("%s/%s.%s" % (opt_root, os.path.basename(path), format1), wav, out_sr, f32=True)
                    except Exception:
                        info += traceback.format_exc()
                log.append("%s->%s" % (os.path.basename(path), info))
                return "\n".join(log)

            if n_lanes == 1 or self.net_g is None:
                for path in todo:
                    info, opt = self.vc_single(sid, path, f0_up_key, None, f0_method, file_index, file_index2, index_rate, filter_radius,
                                               resample_sr, rms_mix_rate, protect)
                    yield record(path, info, opt)
            else:
                # the first file goes through lane 0 alone and synchronously: one-time kernel attribute setup and arena sizing
                info, opt = self.vc_single(sid, todo[0], f0_up_key, None, f0_method, file_index, file_index2, index_rate, filter_radius,
                                           resample_sr, rms_mix_rate, protect)
                yield record(todo[0], info, opt)
                device = torch.device(self.config.device if "cuda" in str(self.config.device) else "cuda:0")
                picked = self._pick_index(file_index, file_index2)
                lanes = [self._lane(i) for i in range(n_lanes)]
                streams = [torch.cuda.Stream(device=device) for _ in range(n_lanes)]
                # a device-resident Index object is one handle (one search workspace): every further lane gets its own
                indexes = [picked if (i == 0 or isinstance(picked, str) or not hasattr(picked, "clone")) else self._index_clone(picked, i)
                           for i in range(n_lanes)]
                pending: List[Optional[tuple]] = [None] * n_lanes
                for i, path in enumerate(todo[1:]):
                    k = i % n_lanes
                    if pending[k] is not None:                      # the oldest utterance in flight: collect it, its lane is free again
                        p_path, p_finish = pending[k]
                        yield record(p_path, *p_finish())
                    pending[k] = (path, self._submit(lanes[k], streams[k], sid, path, f0_up_key, f0_method, indexes[k], index_rate,
                                                     filter_radius, resample_sr, rms_mix_rate, protect))
                n_sub = len(todo) - 1
                for j in range(n_lanes):                            # drain in submission order
                    k = (n_sub + j) % n_lanes
                    if pending[k] is not None:
                        p_path, p_finish = pending[k]
                        pending[k] = None
                        yield record(p_path, *p_finish())
            yield "\n".join(log)
        except Exception:
            yield traceback.format_exc()

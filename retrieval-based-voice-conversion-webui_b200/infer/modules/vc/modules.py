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
        torch.cuda.empty_cache()

    def _load(self, sid) -> str:
        """Build the synthesizer container and the pipeline for ``sid`` (a file under $weight_root, or the dict a .pth holds)."""
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

    def vc_multi(self, sid, dir_path, opt_root, paths, f0_up_key, f0_method, file_index, file_index2, index_rate, filter_radius,
                 resample_sr, rms_mix_rate, protect, format1) -> Iterator[str]:
        try:
            dir_path, opt_root = _unquote(dir_path), _unquote(opt_root)
            os.makedirs(opt_root, exist_ok=True)
            todo = self._inputs(dir_path, paths)
            # the reference is a serial loop on one device; under torchrun the list is strided over ranks exactly like
            # extract_feature_print.py:110
            rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
            log: List[str] = []
            for path in todo[rank::world]:
                info, opt = self.vc_single(sid, path, f0_up_key, None, f0_method, file_index, file_index2, index_rate, filter_radius,
                                           resample_sr, rms_mix_rate, protect)
                if "Success" in info:
                    try:
                        out_sr, wav = opt
                        save_audio("%s/%s.%s" % (opt_root, os.path.basename(path), format1), wav, out_sr, f32=True)
                    except Exception:
                        info += traceback.format_exc()
                log.append("%s->%s" % (os.path.basename(path), info))
                yield "\n".join(log)
            yield "\n".join(log)
        except Exception:
            yield traceback.format_exc()

"""Drop-in for ``infer.modules.vc.modules.VC`` (infer/modules/vc/modules.py:18-266): same method names,
arguments, Gradio-shaped returns and error convention (exceptions inside vc_single become the info string,
modules.py:196-199)."""
from __future__ import annotations

import logging
import os
import traceback

import numpy as np
import torch

from infer.lib.audio import load_audio, save_audio
from rvc.synthesizer import get_synthesizer, load_synthesizer
from .pipeline import Pipeline
from .utils import get_index_path_from_model, load_hubert

logger = logging.getLogger(__name__)


class VC:
    def __init__(self, config):
        self.n_spk = None
        self.tgt_sr = None
        self.net_g = None
        self.pipeline = None
        self.cpt = None
        self.version = None
        self.if_f0 = None
        self.hubert_model = None
        self.config = config

    def get_vc(self, sid, *to_return_protect):
        logger.info("Get sid: " + str(sid))
        to_return_protect0 = {"visible": self.if_f0 != 0,
                              "value": (to_return_protect[0] if self.if_f0 != 0 and to_return_protect else 0.5), "__type__": "update"}
        to_return_protect1 = {"visible": self.if_f0 != 0,
                              "value": (to_return_protect[1] if self.if_f0 != 0 and to_return_protect else 0.33), "__type__": "update"}
        if sid == "" or sid == []:
            if self.hubert_model is not None:
                logger.info("Clean model cache")
                self.hubert_model = self.net_g = self.n_spk = self.tgt_sr = None
                torch.cuda.empty_cache()
            return (({"visible": False, "__type__": "update"}, to_return_protect0, to_return_protect1,
                     {"value": to_return_protect[2], "__type__": "update"}, {"value": to_return_protect[3], "__type__": "update"},
                     {"value": "", "__type__": "update"}) if to_return_protect else {"visible": True, "maximum": 0, "__type__": "update"})
        if isinstance(sid, dict):                 # in-memory checkpoint (tests / bench): same dict a .pth holds
            self.net_g, self.cpt = get_synthesizer(sid, self.config.device)
            sid = sid.get("name", "in-memory")
        else:
            person = f'{os.getenv("weight_root")}/{sid}'
            logger.info(f"Loading: {person}")
            self.net_g, self.cpt = load_synthesizer(person, self.config.device)
        self.tgt_sr = self.cpt["config"][-1]
        self.cpt["config"][-3] = self.cpt["weight"]["emb_g.weight"].shape[0]
        self.if_f0 = self.cpt.get("f0", 1)
        self.version = self.cpt.get("version", "v1")
        self.net_g = self.net_g.half() if self.config.is_half else self.net_g.float()
        self.pipeline = Pipeline(self.tgt_sr, self.config)
        n_spk = self.cpt["config"][-3]
        index = {"value": get_index_path_from_model(sid), "__type__": "update"}
        return (({"visible": True, "maximum": n_spk, "__type__": "update"}, to_return_protect0, to_return_protect1, index, index,
                 {"value": self.cpt.get("info", ""), "__type__": "update"}) if to_return_protect
                else {"visible": True, "maximum": n_spk, "__type__": "update"})

    def vc_single(self, sid, input_audio_path, f0_up_key, f0_file, f0_method, file_index, file_index2, index_rate, filter_radius,
                  resample_sr, rms_mix_rate, protect):
        if input_audio_path is None:
            return "You need to upload an audio", None
        elif hasattr(input_audio_path, "name"):
            input_audio_path = str(input_audio_path.name)
        f0_up_key = int(f0_up_key)
        try:
            audio = input_audio_path if isinstance(input_audio_path, np.ndarray) else load_audio(input_audio_path, 16000)
            audio = np.array(audio, dtype=np.float32)
            audio_max = np.abs(audio).max() / 0.95
            if audio_max > 1:
                np.divide(audio, audio_max, audio)
            times = [0, 0, 0]
            if self.hubert_model is None:
                self.hubert_model = load_hubert(self.config.device, self.config.is_half)
            if file_index is not None and not isinstance(file_index, str) and not hasattr(file_index, "name"):
                pass                                    # device-resident rvc_b200.engine.Index object
            elif file_index:
                if hasattr(file_index, "name"):
                    file_index = str(file_index.name)
                file_index = file_index.strip(" ").strip('"').strip("\n").strip('"').strip(" ").replace("trained", "added")
            elif file_index2:
                file_index = file_index2
            else:
                file_index = ""
            self.pipeline._want_int16 = True        # the device path casts like .astype(np.int16) before its single D2H copy
            try:
                audio_opt = self.pipeline.pipeline(self.hubert_model, self.net_g, sid, audio, times, f0_up_key, f0_method, file_index,
                                                   index_rate, self.if_f0, filter_radius, self.tgt_sr, resample_sr, rms_mix_rate,
                                                   self.version, protect, f0_file).astype(np.int16, copy=False)
            finally:
                self.pipeline._want_int16 = False
            tgt_sr = resample_sr if self.tgt_sr != resample_sr >= 16000 else self.tgt_sr
            index_info = ("Index: %s." % file_index if (not isinstance(file_index, str) or os.path.exists(file_index)) else "Index not used.")
            return ("Success.\n%s\nTime: npy: %.2fs, f0: %.2fs, infer: %.2fs." % (index_info, *times), (tgt_sr, audio_opt))
        except Exception as e:
            info = traceback.format_exc()
            logger.warning(info)
            return str(e), None

    def vc_multi(self, sid, dir_path, opt_root, paths, f0_up_key, f0_method, file_index, file_index2, index_rate, filter_radius,
                 resample_sr, rms_mix_rate, protect, format1):
        try:
            dir_path = dir_path.strip(" ").strip('"').strip("\n").strip('"').strip(" ")
            opt_root = opt_root.strip(" ").strip('"').strip("\n").strip('"').strip(" ")
            os.makedirs(opt_root, exist_ok=True)
            try:
                paths = [os.path.join(dir_path, name) for name in os.listdir(dir_path)] if dir_path != "" else [p.name for p in paths]
            except Exception:
                traceback.print_exc()
                paths = [p.name for p in paths]
            # the reference is a serial loop on one device (modules.py:235-263); under torchrun the list is
            # strided over ranks exactly like extract_feature_print.py:110 (one stream per GPU, no collective)
            rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
            infos = []
            for path in paths[rank::world]:
                info, opt = self.vc_single(sid, path, f0_up_key, None, f0_method, file_index, file_index2, index_rate, filter_radius,
                                           resample_sr, rms_mix_rate, protect)
                if "Success" in info:
                    try:
                        tgt_sr, audio_opt = opt
                        save_audio("%s/%s.%s" % (opt_root, os.path.basename(path), format1), audio_opt, tgt_sr, f32=True)
                    except Exception:
                        info += traceback.format_exc()
                infos.append("%s->%s" % (os.path.basename(path), info))
                yield "\n".join(infos)
            yield "\n".join(infos)
        except Exception:
            yield traceback.format_exc()

"""Drop-in for ``infer/modules/train/extract_feature_print.py`` (the script that fills ``3_feature{256,768}/`` of an
experiment folder -- the vectors the retrieval index is later built from): same command line
(``device n_part i_part i_gpu exp_dir version is_half``), same inputs (``<exp_dir>/1_16k_wavs/*.wav``), same outputs
(one ``.npy`` of shape [T_h, 256 | 768] per file), same log lines in ``<exp_dir>/extract_f0_feature.log``.
The features come from the sm_100a HuBERT (librvcb200) through the same duck-typed model object the inference path uses;
the file list is strided ``[i_part::n_part]`` like the reference (:110), one process per GPU, no collective.

    python -m infer.modules.train.extract_feature_print cuda 2 0 0 logs/my-voice v2 True
"""
from __future__ import annotations

import os
import re
import sys
import traceback
from typing import Callable, List, Optional

import numpy as np
import torch


def feature_dir(exp_dir: str, version: str) -> str:
    return "%s/3_feature256" % exp_dir if version == "v1" else "%s/3_feature768" % exp_dir


def extract_features(model, wav: np.ndarray, version: str) -> np.ndarray:
    """One utterance -> [T_h, 256 | 768] float32 (reference :118-137: layer 9 + final_proj for v1, layer 12 for v2)."""
    feats = torch.from_numpy(np.ascontiguousarray(wav, dtype=np.float32))
    assert feats.dim() == 1, feats.dim()
    feats = feats.view(1, -1)
    with torch.no_grad():
        logits = model.extract_features(source=feats, padding_mask=torch.zeros(feats.shape, dtype=torch.bool),
                                        output_layer=9 if version == "v1" else 12)
        out = model.final_proj(logits[0]) if version == "v1" else logits[0]
    return out.squeeze(0).float().cpu().numpy()


def run(model, exp_dir: str, version: str, n_part: int, i_part: int, load_wav: Callable[[str], np.ndarray], log: Callable[[str], None]) -> int:
    """The per-process loop (reference :109-150).  Returns the number of files written."""
    wav_dir, out_dir = "%s/1_16k_wavs" % exp_dir, feature_dir(exp_dir, version)
    os.makedirs(out_dir, exist_ok=True)
    todo: List[str] = sorted(os.listdir(wav_dir))[i_part::n_part]
    if not todo:
        log("no-feature-todo")
        return 0
    log("all-feature-%s" % len(todo))
    every, written = max(1, len(todo) // 10), 0          # at most ten progress lines
    for idx, name in enumerate(todo):
        try:
            if not name.endswith(".wav"):
                continue
            out_path = "%s/%s" % (out_dir, name.replace("wav", "npy"))
            if os.path.exists(out_path):
                continue
            feats = extract_features(model, load_wav("%s/%s" % (wav_dir, name)), version)
            if np.isnan(feats).sum() == 0:
                np.save(out_path, feats, allow_pickle=False)
                written += 1
            else:
                log("%s-contains nan" % name)
            if idx % every == 0:
                log("now-%s,all-%s,%s,%s" % (len(todo), idx, name, feats.shape))
        except Exception:
            log(traceback.format_exc())
    log("all-feature-done")
    return written


def main(argv: Optional[List[str]] = None) -> int:
    argv = list(sys.argv if argv is None else argv)
    if len(argv) != 8:
        return 0
    _device, n_part, i_part, i_gpu, exp_dir, version, is_half = argv[1], int(argv[2]), int(argv[3]), argv[4], argv[5], argv[6], argv[7].lower() == "true"
    # CUDA_VISIBLE_DEVICES wants bare indices; callers pass "0", "cuda:0", "cuda:0-cuda:1", ... (reference :16-23)
    os.environ["CUDA_VISIBLE_DEVICES"] = re.sub(r"cuda:", "", str(i_gpu)).replace("-", ",")
    with open("%s/extract_f0_feature.log" % exp_dir, "a+") as f:
        def log(msg: str) -> None:
            print(msg)
            f.write("%s\n" % msg)
            f.flush()
        log(" ".join(argv))
        log("exp_dir: " + exp_dir)
        model_path = "assets/hubert/hubert_base.pt"
        log("load model(s) from {}".format(model_path))
        if not os.access(model_path, os.F_OK):
            log("Error: Extracting is shut down because %s does not exist." % model_path)
            return 0
        from infer.lib.audio import load_audio
        from infer.modules.vc.utils import load_hubert
        model = load_hubert("cuda:0", is_half)
        log("move model to cuda")
        run(model, exp_dir, version, n_part, i_part, lambda p: load_audio(p, 16000), log)
    return 0


if __name__ == "__main__":
    sys.exit(main())

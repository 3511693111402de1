"""Drop-in for ``infer/modules/train/extract_f0_print.py`` (fills ``2a_f0/`` with the coarse mel bins and ``2b-f0nsf/`` with the
f0 track in Hz of every ``<exp_dir>/1_16k_wavs/*.wav``): same command line (``exp_dir n_p f0method device is_half``), same
output names (``np.save`` appends ``.npy`` to ``<name>.wav``), same log lines.  f0 comes from the sm_100a RMVPE through
``rvc.f0.Generator`` -- the object the inference path uses; only ``rmvpe`` is available (the CPU estimators are out of scope).
With a GPU method the reference runs one worker (``n_p = 1``, :112-116); under torchrun the list is strided over ranks instead.

    python -m infer.modules.train.extract_f0_print logs/my-voice 1 rmvpe cuda:0 True
"""
from __future__ import annotations

import os
import sys
import traceback
from pathlib import Path
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

HOP, FS = 160, 16000


def list_jobs(exp_dir: str) -> List[Tuple[str, str, str]]:
    """(input wav, coarse output stem, f0 output stem) for every usable file (reference :121-131)."""
    inp_root, opt_root1, opt_root2 = "%s/1_16k_wavs" % exp_dir, "%s/2a_f0" % exp_dir, "%s/2b-f0nsf" % exp_dir
    os.makedirs(opt_root1, exist_ok=True)
    os.makedirs(opt_root2, exist_ok=True)
    jobs = []
    for name in sorted(os.listdir(inp_root)):
        inp_path = "%s/%s" % (inp_root, name)
        if "spec" in inp_path:
            continue
        jobs.append((inp_path, "%s/%s" % (opt_root1, name), "%s/%s" % (opt_root2, name)))
    return jobs


def run(f0_gen, jobs: Sequence[Tuple[str, str, str]], f0_method: str, load_wav: Callable[[str], np.ndarray], log: Callable[[str], None]) -> int:
    """The worker loop (reference FeatureInput.go, :64-97).  Returns the number of utterances written."""
    if len(jobs) == 0:
        log("no-f0-todo")
        return 0
    log("todo-f0-%s" % len(jobs))
    every, written = max(len(jobs) // 5, 1), 0           # at most five progress lines per worker
    for idx, (inp_path, coarse_stem, f0_stem) in enumerate(jobs):
        try:
            if idx % every == 0:
                log("f0ing,now-%s,all-%s,-%s" % (idx, len(jobs), inp_path))
            if os.path.exists(coarse_stem + ".npy") and os.path.exists(f0_stem + ".npy"):
                continue
            x = load_wav(inp_path)
            coarse, f0 = f0_gen.calculate(x, x.shape[0] // HOP, 0, f0_method, None)
            np.save(f0_stem, f0, allow_pickle=False)          # nsf
            np.save(coarse_stem, coarse, allow_pickle=False)  # ori
            written += 1
        except Exception:
            log("f0fail-%s-%s-%s" % (idx, inp_path, traceback.format_exc()))
    return written


def main(argv: Optional[List[str]] = None) -> int:
    argv = list(sys.argv if argv is None else argv)
    exp_dir, _n_p, f0_method, device, is_half = argv[1], int(argv[2]), argv[3], argv[4], argv[5] == "True"
    with open("%s/extract_f0_feature.log" % exp_dir, "a+") as f:
        def log(msg: str) -> None:
            print(msg)
            f.write("%s\n" % msg)
            f.flush()
        log(" ".join(argv))
        if "cuda" in device:
            log("WARN: use 1 thread since GPU is used.")
        from infer.lib.audio import load_audio
        from rvc.f0 import Generator
        f0_gen = Generator(Path(os.environ["rmvpe_root"]), is_half, 0, device if "cuda" in device else "cuda:0", HOP, FS)
        rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
        run(f0_gen, list_jobs(exp_dir)[rank::world], f0_method, lambda p: load_audio(p, FS), log)
    return 0


if __name__ == "__main__":
    sys.exit(main())

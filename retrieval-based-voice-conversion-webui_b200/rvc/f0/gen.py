"""Drop-in for ``rvc.f0.Generator`` (rvc/f0/gen.py:44-141) for the hot path's f0 method: "rmvpe".
The RMVPE network, mel front end and salience decode run in sm_100a kernels (librvcb200).  ``calculate_device`` keeps the O(T)
resize / gap-fill / mel-quantise post-processing on the device too (rvcb_f0_post, bit-equal to the host form); ``calculate``
returns host arrays like the reference and runs that post-processing on the host (f0post.py; needed for manual f0 curves).
The CPU third-party estimators (pm, dio, harvest, crepe, fcpe) are out of scope."""
from __future__ import annotations

from pathlib import Path
from typing import Optional, Tuple, Union

import numpy as np
import torch

from rvc_b200 import f0post
from rvc_b200.engine import Rmvpe, f0_post as engine_f0_post


def post_process(tf0, f0, f0_up_key, manual_x_pad, f0_mel_min=None, f0_mel_max=None, manual_f0=None):
    """Same positional signature as the numba function at rvc/f0/gen.py:10-41."""
    return f0post.post_process(f0, f0_up_key, tf0, manual_x_pad, manual_f0)


class Generator(object):
    def __init__(self, rmvpe_root: Union[Path, str, dict], is_half: bool, x_pad: int, device="cuda:0", window=160, sr=16000):
        self.rmvpe_root = rmvpe_root        # directory holding rmvpe.pt, or an E2E state_dict
        self.is_half = is_half
        self.x_pad = x_pad
        self.device = torch.device(device if "cuda" in str(device) else "cuda:0")
        self.window = window
        self.sr = sr

    def _rmvpe(self) -> Rmvpe:
        if not hasattr(self, "rmvpe"):
            if isinstance(self.rmvpe_root, dict):
                sd = self.rmvpe_root
            else:
                sd = torch.load(str(Path(self.rmvpe_root) / "rmvpe.pt"), map_location="cpu", weights_only=True)
            self.rmvpe = Rmvpe(sd, self.device.index or 0)
        return self.rmvpe

    def compute_f0_rmvpe(self, x, p_len: Optional[int], thred: float = 0.03) -> np.ndarray:
        """RMVPE.compute_f0 (rvc/f0/rmvpe.py:96-117): one D2H of n_frames floats, like the reference's
        ``hidden.cpu()`` but 360x smaller."""
        wav = x if torch.is_tensor(x) else torch.from_numpy(np.asarray(x, dtype=np.float32))
        wav = wav.to(self.device, dtype=torch.float32)
        if p_len is None:
            p_len = wav.shape[0] // self.window
        f0, _, _ = self._rmvpe().infer(wav, thred)
        return f0post.interpolate_f0(f0post.resize_f0(f0.cpu().numpy().astype(np.float64), p_len))

    def calculate_device(self, wav: torch.Tensor, p_len: int, f0_up_key: float, thred: float = 0.03):
        """``calculate`` for f0_method "rmvpe" without a manual curve, entirely on the device: RMVPE, then the resize / gap
        fill / key shift / mel quantisation kernel.  Returns device tensors (pitch int64 [p_len], pitchf float32 [p_len])."""
        f0, _, _ = self._rmvpe().infer(wav, thred)
        return engine_f0_post(f0, p_len, f0_up_key)

    def calculate(self, x, p_len: Optional[int], f0_up_key: int, f0_method: str, filter_radius, manual_f0=None) -> Tuple[np.ndarray, np.ndarray]:
        if f0_method != "rmvpe":
            raise ValueError(f"f0 method {f0_method} has not yet been supported")
        f0 = self.compute_f0_rmvpe(x, p_len, 0.03)
        return f0post.post_process(f0, f0_up_key, self.sr // self.window, self.x_pad, manual_f0)

"""Host-side f0 post-processing (O(T) work on <= a few thousand frames; the reference does the
same on the CPU with pure-python loops / numba):

  F0Predictor._resize_f0        rvc/f0/f0.py:69-78
  F0Predictor._interpolate_f0   rvc/f0/f0.py:31-67   (run-based rewrite of the frame loop)
  post_process                  rvc/f0/gen.py:10-41
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np


def resize_f0(x: np.ndarray, target_len: int) -> np.ndarray:
    source = np.array(x, dtype=np.float64)
    source[source < 0.001] = np.nan
    target = np.interp(np.arange(0, len(source) * target_len, len(source)) / target_len, np.arange(0, len(source)), source)
    return np.nan_to_num(target)


def interpolate_f0(f0: np.ndarray) -> np.ndarray:
    """Fill unvoiced runs: interior run after a voiced frame -> linear ramp to the next voiced value;
    leading run -> copy the next voiced value; a run whose next voiced frame is the last frame (or that
    reaches the end) -> hold the last voiced value.  Same results as the reference's in-place frame loop."""
    data = np.array(f0, dtype=np.float64).reshape(-1)
    n = data.size
    i = 0
    last_value = 0.0
    while i < n:
        if data[i] > 0.0:
            last_value = data[i]
            i += 1
            continue
        j = i + 1
        while j < n and not data[j] > 0.0:
            j += 1
        # reference: `for j in range(i+1, n): if data[j] > 0: break` leaves j = n-1 when nothing is found
        # (or i+1 when the range is empty)
        jj = j if j < n else (n - 1 if i + 1 < n else i + 1)
        if jj < n - 1:
            if last_value > 0.0:
                step = (data[jj] - data[i - 1]) / float(jj - i)
                k = np.arange(i, jj)
                data[i:jj] = data[i - 1] + step * (k - i + 1)
            else:
                data[i:jj] = data[jj]
            # frames i..jj-1 are now voiced; the loop in the reference continues at i+1 and sees them voiced
            last_value = data[jj - 1] if jj > i else last_value
            i = jj
        else:
            data[i:n] = last_value
            i = n
    return data


def post_process(f0: np.ndarray, f0_up_key: float, tf0: int = 100, manual_x_pad: int = 0, manual_f0: Optional[np.ndarray] = None,
                 f0_min: float = 50.0, f0_max: float = 1100.0) -> Tuple[np.ndarray, np.ndarray]:
    f0 = np.multiply(f0, pow(2, f0_up_key / 12))
    if manual_f0 is not None:
        manual_f0 = np.asarray(manual_f0)
        delta_t = np.round((manual_f0[:, 0].max() - manual_f0[:, 0].min()) * tf0 + 1).astype("int16")
        replace_f0 = np.interp(list(range(delta_t)), manual_f0[:, 0] * 100, manual_f0[:, 1])
        shape = f0[manual_x_pad * tf0: manual_x_pad * tf0 + len(replace_f0)].shape[0]
        f0[manual_x_pad * tf0: manual_x_pad * tf0 + len(replace_f0)] = replace_f0[:shape]
    mel_min = 1127 * math.log(1 + f0_min / 700)
    mel_max = 1127 * math.log(1 + f0_max / 700)
    f0_mel = 1127 * np.log(1 + f0 / 700)
    f0_mel[f0_mel > 0] = (f0_mel[f0_mel > 0] - mel_min) * 254 / (mel_max - mel_min) + 1
    f0_mel[f0_mel <= 1] = 1
    f0_mel[f0_mel > 255] = 255
    return np.rint(f0_mel).astype(np.int32), f0

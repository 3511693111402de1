"""Host audio I/O for the drop-in front door (reference: infer/lib/audio.py:49-159 uses PyAV, which is not a
dependency here; host I/O is out of the hot path -- SURVEY §2.1).  WAV via scipy, resample via polyphase."""
from __future__ import annotations

from math import gcd

import numpy as np
from scipy import signal
from scipy.io import wavfile


def load_audio(file: str, sr: int) -> np.ndarray:
    """-> mono float32 at ``sr`` (same contract as load_audio(file, 16000), infer/lib/audio.py:78)."""
    file = str(file).strip(" ").strip('"').strip("\n").strip('"').strip(" ")
    src_sr, x = wavfile.read(file)
    if x.dtype == np.int16:
        x = x.astype(np.float32) / 32768.0
    elif x.dtype == np.int32:
        x = x.astype(np.float32) / 2147483648.0
    elif x.dtype == np.uint8:
        x = (x.astype(np.float32) - 128.0) / 128.0
    else:
        x = x.astype(np.float32)
    if x.ndim == 2:
        x = x.mean(axis=1)
    if src_sr != sr:
        g = gcd(int(src_sr), int(sr))
        x = signal.resample_poly(x, sr // g, src_sr // g).astype(np.float32)
    return np.ascontiguousarray(x, dtype=np.float32)


def save_audio(path: str, audio: np.ndarray, sr: int, f32: bool = False, format: str = "wav") -> None:
    a = np.asarray(audio)
    if f32 and a.dtype != np.int16:
        a = a.astype(np.float32)
    wavfile.write(path, sr, a)

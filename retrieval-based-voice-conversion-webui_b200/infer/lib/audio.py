"""Host audio I/O for the drop-in front door (reference: infer/lib/audio.py:49-159 uses PyAV, which is not a
dependency here; host I/O is out of the hot path -- SURVEY §2.1).  WAV via scipy, resample via polyphase."""
from __future__ import annotations

import math
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
    """infer/lib/audio.py:35-56.  f32=True writes a float32 WAV of the samples AS THEY ARE -- vc_multi passes the int16-range result,
    so the file holds float32 values in +-32768, exactly like the reference's ``wavfile.write(buf, sr, wav.astype(np.float32))``;
    f32=False writes 16-bit PCM from floats in [-1, 1].  Only the wav container is built (other formats go through PyAV there)."""
    a = np.asarray(audio)
    if format != "wav":
        raise NotImplementedError("only wav output is built (the reference transcodes other formats with PyAV)")
    if f32:
        a = a.astype(np.float32)
    elif a.dtype != np.int16:
        am = int(math.ceil(float(np.abs(a).max())) * 32768) if a.size else 0          # float_to_int16, audio.py:29-32
        a = np.multiply(a, 32767 * 32768 // am).astype(np.int16) if am else np.zeros(a.shape, np.int16)
    wavfile.write(path, sr, a)

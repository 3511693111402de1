"""Thin Python owners of the librvcb200 handles.  PyTorch is used only for device memory,
streams and host<->device copies (plumbing); all arithmetic happens inside the C ABI calls."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib


def _stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def set_grid_cap(max_ctas: int) -> int:
    """Cap the persistent GEMM / kNN grids launched from now on (0 = all SMs); returns the previous cap."""
    return _lib.lib().rvcb_set_grid_cap(int(max_ctas))


def front_branch_cap() -> int:
    """CTAs per kernel while the two front branches are launched (RVCB_FRONT_CAP, default: half of the 148 SMs)."""
    import os
    return int(os.environ.get("RVCB_FRONT_CAP", "74"))


def _chk_dev(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (the hot path has no CPU fallback)")
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


class Weights:
    """Host fp32 tensors keyed by the reference's state_dict names -> rvcb_weights handle."""

    def __init__(self, state_dict: Dict[str, torch.Tensor]):
        _lib.lib()
        self.h = C.c_void_p()
        _lib.check(_lib.lib().rvcb_weights_create(C.byref(self.h)))
        for k, v in state_dict.items():
            if not torch.is_tensor(v) or not (v.is_floating_point()):
                continue
            a = np.ascontiguousarray(v.detach().cpu().float().numpy())
            shape = (C.c_int64 * max(a.ndim, 1))(*a.shape)
            _lib.check(_lib.lib().rvcb_weights_add(self.h, k.encode(), a.ctypes.data_as(C.c_void_p), a.ndim, shape))

    def close(self):
        if self.h:
            _lib.lib().rvcb_weights_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class Hubert:
    """Replaces the fairseq HubertModel object for the calls the pipeline makes."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device: int = 0):
        _lib.init(device)
        self.h = C.c_void_p()
        with Weights(state_dict) as w:
            _lib.check(_lib.lib().rvcb_hubert_create(w.h, C.byref(self.h)))
        self.device = torch.device("cuda", device)

    @staticmethod
    def num_frames(n_samples: int) -> int:
        return _lib.lib().rvcb_hubert_num_frames(n_samples)

    def extract(self, wav: torch.Tensor, output_layer: int = 12) -> torch.Tensor:
        """wav f32 [n] (device) -> f32 [T_h, 768] (device)"""
        wav = _chk_dev(wav.reshape(-1), torch.float32, "wav")
        T = self.num_frames(wav.numel())
        if T < 1:
            raise RuntimeError("hubert: input too short")
        out = torch.empty(T, 768, device=wav.device, dtype=torch.float32)
        nf = C.c_int(0)
        _lib.check(_lib.lib().rvcb_hubert_extract_features(self.h, _p(wav), wav.numel(), output_layer, _p(out), C.byref(nf), _stream_ptr()))
        return out

    def final_proj(self, x: torch.Tensor) -> torch.Tensor:
        x = _chk_dev(x.reshape(-1, 768), torch.float32, "x")
        out = torch.empty(x.shape[0], 256, device=x.device, dtype=torch.float32)
        _lib.check(_lib.lib().rvcb_hubert_final_proj(self.h, _p(x), x.shape[0], _p(out), _stream_ptr()))
        return out

    def __del__(self):
        try:
            if self.h:
                _lib.lib().rvcb_hubert_destroy(self.h)
        except Exception:
            pass


class Index:
    """IVF-Flat index on the device; duck-types the faiss object the pipeline uses
    (.search / .reconstruct_n / .ntotal) with device tensors in and out."""

    def __init__(self, centroids: np.ndarray, vectors: np.ndarray, list_off: np.ndarray, list_ids: np.ndarray, device: int = 0):
        _lib.init(device)
        self.centroids = np.ascontiguousarray(centroids, dtype=np.float32)
        self.vectors = np.ascontiguousarray(vectors, dtype=np.float32)
        lo = np.ascontiguousarray(list_off, dtype=np.int64)
        li = np.ascontiguousarray(list_ids, dtype=np.int64)
        self.ntotal, self.d = self.vectors.shape
        self.nlist = self.centroids.shape[0]
        self._lists = (lo, li)
        self.device = torch.device("cuda", device)
        self.h = C.c_void_p()
        _lib.check(_lib.lib().rvcb_index_create(self.centroids.ctypes.data_as(C.c_void_p), self.nlist,
                                                self.vectors.ctypes.data_as(C.c_void_p), self.ntotal, self.d,
                                                lo.ctypes.data_as(C.c_void_p), li.ctypes.data_as(C.c_void_p), C.byref(self.h)))

    @classmethod
    def from_oracle_layout(cls, idx, device: int = 0) -> "Index":
        """idx: any object with .centroids, .vectors, .list_off, .list_ids (e.g. oracle.ivf.IVFFlat or the
        .index reader)."""
        return cls(idx.centroids, idx.vectors, idx.list_off, idx.list_ids, device)

    def clone(self) -> "Index":
        """A second device handle over the same host arrays (the search workspace is per handle: utterances in flight at the same
        time -- VC.vc_multi lanes -- must not share one)."""
        return Index(self.centroids, self.vectors, self._lists[0], self._lists[1], self.device.index or 0)

    def reconstruct_n(self, i0: int, n: int) -> np.ndarray:
        return self.vectors[i0:i0 + n]

    def search_device(self, q: torch.Tensor, k: int = 8) -> Tuple[torch.Tensor, torch.Tensor]:
        q = _chk_dev(q.reshape(-1, self.d), torch.float32, "q")
        D = torch.empty(q.shape[0], k, device=q.device, dtype=torch.float32)
        I = torch.empty(q.shape[0], k, device=q.device, dtype=torch.int64)
        _lib.check(_lib.lib().rvcb_index_search(self.h, _p(q), q.shape[0], k, _p(D), _p(I), _stream_ptr()))
        return D, I

    def search(self, x, k: int = 8):
        """faiss-style: numpy in, numpy out (host buffers cross PCIe inside this call)."""
        q = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(self.device)
        D, I = self.search_device(q, k)
        return D.cpu().numpy(), I.cpu().numpy()

    def blend_device(self, feats: torch.Tensor, D: torch.Tensor, I: torch.Tensor, index_rate: float) -> torch.Tensor:
        feats = _chk_dev(feats.reshape(-1, self.d), torch.float32, "feats")
        out = torch.empty_like(feats)
        _lib.check(_lib.lib().rvcb_index_blend(self.h, _p(feats), feats.shape[0], D.shape[1], _p(D), _p(I), float(index_rate), _p(out),
                                               _stream_ptr()))
        return out

    def __del__(self):
        try:
            if self.h:
                _lib.lib().rvcb_index_destroy(self.h)
        except Exception:
            pass


def knn_bruteforce_top1(db: torch.Tensor, q: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    _lib.init(db.device.index or 0)
    db = _chk_dev(db, torch.float32, "db")
    q = _chk_dev(q, torch.float32, "q")
    D = torch.empty(q.shape[0], device=q.device, dtype=torch.float32)
    I = torch.empty(q.shape[0], device=q.device, dtype=torch.int64)
    _lib.check(_lib.lib().rvcb_knn_bruteforce_top1(_p(db), db.shape[0], db.shape[1], _p(q), q.shape[0], _p(D), _p(I), _stream_ptr()))
    return D, I


class FlatIndex:
    """Exact brute-force L2 top-1 over a device-resident fp32 database, tensor-core short list for query batches
    (rvcb_flat_*): results bit-identical to ``knn_bruteforce_top1``."""

    def __init__(self, db: torch.Tensor):
        _lib.init(db.device.index or 0)
        self.db = _chk_dev(db, torch.float32, "db")          # kept alive: the handle reads it
        self.h = C.c_void_p()
        _lib.check(_lib.lib().rvcb_flat_create(_p(self.db), self.db.shape[0], self.db.shape[1], C.byref(self.h)))

    def search(self, q: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        q = _chk_dev(q.reshape(-1, self.db.shape[1]), torch.float32, "q")
        D = torch.empty(q.shape[0], device=q.device, dtype=torch.float32)
        I = torch.empty(q.shape[0], device=q.device, dtype=torch.int64)
        _lib.check(_lib.lib().rvcb_flat_search_top1(self.h, _p(q), q.shape[0], _p(D), _p(I), _stream_ptr()))
        return D, I

    def __del__(self):
        try:
            if self.h:
                _lib.lib().rvcb_flat_destroy(self.h)
        except Exception:
            pass


def upsample_protect(feats: torch.Tensor, feats0: Optional[torch.Tensor], pitchf: Optional[torch.Tensor], T: int, protect: float) -> torch.Tensor:
    feats = _chk_dev(feats, torch.float32, "feats")
    T_h, Cc = feats.shape
    out = torch.empty(T, Cc, device=feats.device, dtype=torch.float32)
    f0 = None if feats0 is None else _chk_dev(feats0, torch.float32, "feats0")
    pf = None if pitchf is None else _chk_dev(pitchf.reshape(-1), torch.float32, "pitchf")
    _lib.check(_lib.lib().rvcb_upsample_protect(_p(feats), _p(f0), T_h, Cc, _p(pf), T, float(protect), _p(out), _stream_ptr()))
    return out


def post_mix(wav: torch.Tensor, tgt_sr: int, audio16k: torch.Tensor, rms_mix_rate: float) -> torch.Tensor:
    """In place on ``wav`` (device f32 [n]): RMS-envelope mix with the 16 kHz input + peak normalisation to the int16 range."""
    wav = _chk_dev(wav.reshape(-1), torch.float32, "wav")
    a = _chk_dev(audio16k.reshape(-1), torch.float32, "audio16k")
    scratch = torch.empty(a.numel() // 8000 + wav.numel() // (tgt_sr // 2) + 16, device=wav.device, dtype=torch.float64)
    _lib.check(_lib.lib().rvcb_post_mix(_p(wav), wav.numel(), int(tgt_sr), _p(a), a.numel(), float(rms_mix_rate), _p(scratch), _stream_ptr()))
    return wav


def rms_mix(wav: torch.Tensor, tgt_sr: int, audio16k: torch.Tensor, rms_mix_rate: float) -> torch.Tensor:
    """In place on ``wav``: the RMS-envelope mix alone (pipeline.py:349-350), no peak scaling (rvcb_rms_mix)."""
    wav = _chk_dev(wav.reshape(-1), torch.float32, "wav")
    a = _chk_dev(audio16k.reshape(-1), torch.float32, "audio16k")
    scratch = torch.empty(a.numel() // 8000 + wav.numel() // (tgt_sr // 2) + 16, device=wav.device, dtype=torch.float64)
    _lib.check(_lib.lib().rvcb_rms_mix(_p(wav), wav.numel(), int(tgt_sr), _p(a), a.numel(), float(rms_mix_rate), _p(scratch), _stream_ptr()))
    return wav


def rt_tail(infer_wav: torch.Tensor, input_wav: Optional[torch.Tensor], zc: int, rms_mix_rate: float, sola_buffer: torch.Tensor,
            block_frame: int, sola_search_frame: int, want_offset: bool = False, use_pv: bool = False):
    """gui.py:1024-1087 on the device: envelope mix (in place on ``infer_wav``) + SOLA.  Updates ``sola_buffer`` in place and
    returns the output block (and the offset tensor if asked)."""
    y = _chk_dev(infer_wav.reshape(-1), torch.float32, "infer_wav")
    buf = sola_buffer
    if not (buf.is_cuda and buf.dtype == torch.float32 and buf.is_contiguous()):
        raise RuntimeError("sola_buffer must be a contiguous float32 CUDA tensor (it is updated in place)")
    x = None if input_wav is None else _chk_dev(input_wav.reshape(-1), torch.float32, "input_wav")
    n = y.numel()
    out = torch.empty(block_frame, device=y.device, dtype=torch.float32)
    nb = int(buf.numel())
    scratch = torch.empty(2 * (n // zc + 1) + sola_search_frame + 8 + (3 * (nb // 2 + 1) + nb if use_pv else 0), device=y.device,
                          dtype=torch.float32)
    off = torch.empty(1, device=y.device, dtype=torch.int32) if want_offset else None
    _lib.check(_lib.lib().rvcb_rt_tail_pv(_p(y), n, _p(x), int(zc), float(rms_mix_rate), _p(buf), int(block_frame), nb,
                                          int(sola_search_frame), int(bool(use_pv)), _p(out), _p(scratch), _p(off), _stream_ptr()))
    return (out, off) if want_offset else out


class TorchGateHandle:
    """rvcb_torchgate_*: the realtime GUI's spectral gate on the device (fp32).  ``filt``: HOST float32 [rows, cols] smoothing filter
    or None."""

    def __init__(self, sr: int, n_fft: int, hop: int, nonstationary: bool, n_std_thresh_stationary: float, n_thresh_nonstationary: float,
                 temp_coeff_nonstationary: float, n_movemean_nonstationary: int, prop_decrease: float, filt: Optional[torch.Tensor],
                 device_index: int = 0):
        _lib.init(device_index)
        self.h = C.c_void_p()
        f = None if filt is None else filt.detach().to("cpu", torch.float32).contiguous()
        _lib.check(_lib.lib().rvcb_torchgate_create(int(sr), int(n_fft), int(hop), int(bool(nonstationary)), float(n_std_thresh_stationary),
                                                    float(n_thresh_nonstationary), float(temp_coeff_nonstationary), int(n_movemean_nonstationary),
                                                    float(prop_decrease), None if f is None else f.data_ptr(), 0 if f is None else f.shape[0],
                                                    0 if f is None else f.shape[1], C.byref(self.h)))
        self.hop = int(hop)

    def apply(self, x: torch.Tensor, xn: Optional[torch.Tensor]) -> torch.Tensor:
        x = _chk_dev(x.reshape(-1), torch.float32, "x")
        n = x.numel()
        noise = None if xn is None else _chk_dev(xn.reshape(-1), torch.float32, "xn")
        y = torch.empty(self.hop * (n // self.hop), device=x.device, dtype=torch.float32)
        _lib.check(_lib.lib().rvcb_torchgate_apply(self.h, _p(x), n, _p(noise), 0 if noise is None else noise.numel(), _p(y), _stream_ptr()))
        return y

    def __del__(self):
        try:
            if self.h:
                _lib.lib().rvcb_torchgate_destroy(self.h)
        except Exception:
            pass


def sinc_resample_kernel(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99, dtype=torch.float32):
    """The [new/gcd, 2*width + orig/gcd] windowed-sinc table of torchaudio.transforms.Resample(orig, new, dtype=dtype) with its
    defaults (sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99), same operation order as torchaudio's
    _get_sinc_resample_kernel so the table is equal to the one gui.py:851-866 builds.  Returns (kernel [up, kw] CPU, width, up, down)."""
    import math
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base_freq = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base_freq)
    idx_dtype = dtype if dtype is not None else torch.float64
    idx = torch.arange(-width, width + orig, dtype=idx_dtype)[None, None] / orig
    t = torch.arange(0, -new, -1, dtype=dtype)[:, None, None] / new + idx
    t *= base_freq
    t = t.clamp_(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t *= math.pi
    scale = base_freq / orig
    kernels = torch.where(t == 0, torch.tensor(1.0).to(t), t.sin() / t)
    kernels *= window * scale
    if dtype is None:
        kernels = kernels.to(dtype=torch.float32)
    return kernels[:, 0].contiguous().float(), width, new, orig


def sinc_resample(x: torch.Tensor, kernel: torch.Tensor, width: int, up: int, down: int) -> torch.Tensor:
    """rvcb_resample_sinc: x f32[n] (device) -> f32[ceil(up * n / down)]; ``kernel`` f32[up, kw] on the device."""
    x = _chk_dev(x.reshape(-1), torch.float32, "x")
    k = _chk_dev(kernel, torch.float32, "kernel")
    n = x.numel()
    n_out = -((-up * n) // down)
    out = torch.empty(n_out, device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().rvcb_resample_sinc(_p(x), n, _p(k), int(up), int(down), int(k.shape[1]), int(width), _p(out), n_out, _stream_ptr()))
    return out


def host_filtfilt(b: np.ndarray, a: np.ndarray, zi: np.ndarray, x: np.ndarray) -> np.ndarray:
    """scipy.signal.filtfilt(b, a, x) (defaults) on the host, in C (rvcb_host_filtfilt); float32 / float64 in, float64 out."""
    b = np.ascontiguousarray(b, dtype=np.float64); a = np.ascontiguousarray(a, dtype=np.float64)
    zi = np.ascontiguousarray(zi, dtype=np.float64)
    x = np.asarray(x).reshape(-1)
    is32 = x.dtype == np.float32
    x = np.ascontiguousarray(x, dtype=np.float32 if is32 else np.float64)
    y = np.empty(x.shape[0], dtype=np.float64)
    _lib.check(_lib.lib().rvcb_host_filtfilt(b.ctypes.data, a.ctypes.data, zi.ctypes.data, int(b.shape[0]), x.ctypes.data, int(is32),
                                             int(x.shape[0]), y.ctypes.data))
    return y


def highpass_sos_from_ba(b: np.ndarray, a: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Second-order sections of an odd-order Butterworth high-pass given in (b, a) form, for the device filter: the zeros are
    the N-fold zero at z = 1 (known exactly; root finding would smear it by eps^(1/N)), the poles are the roots of ``a``
    polished by Newton steps in extended precision.  Returns (sos [ns,6], zi [ns,2]) with zi = scipy.signal.sosfilt_zi(sos)."""
    from scipy import signal
    order = len(a) - 1
    al = np.array(a, dtype=np.longdouble)
    r = np.roots(a).astype(np.clongdouble)
    for _ in range(50):
        r = r - np.polyval(al, r) / np.polyval(np.polyder(al), r)
    r = np.array(r, dtype=np.complex128)
    cplx = sorted([z for z in r if z.imag > 1e-9], key=abs)
    real = [z.real for z in r if abs(z.imag) <= 1e-9]
    if 2 * len(cplx) + len(real) != order or len(real) > 1:
        raise ValueError("unexpected pole pattern")
    sos = [[1.0, -2.0, 1.0, 1.0, -2.0 * z.real, abs(z) ** 2] for z in cplx] + [[1.0, -1.0, 0.0, 1.0, -p, 0.0] for p in real]
    sos = np.array(sos, dtype=np.float64)
    sos[0, :3] *= b[0]
    return sos, np.ascontiguousarray(signal.sosfilt_zi(sos), dtype=np.float64)


def sosfiltfilt(sos: np.ndarray, zi: np.ndarray, edge: int, x: torch.Tensor) -> torch.Tensor:
    """Device forward-backward IIR over second-order sections (float64 arithmetic): x f32[n] on the device -> f32[n]."""
    sos = np.ascontiguousarray(sos, dtype=np.float64); zi = np.ascontiguousarray(zi, dtype=np.float64)
    x = _chk_dev(x.reshape(-1), torch.float32, "x")
    y = torch.empty_like(x)
    scratch = torch.empty(x.numel() + 2 * edge + 8 + 4 * 1024 + 16, device=x.device, dtype=torch.float64)
    _lib.check(_lib.lib().rvcb_sosfiltfilt(sos.ctypes.data, zi.ctypes.data, int(sos.shape[0]), int(edge), _p(x), x.numel(), _p(y),
                                           _p(scratch), _stream_ptr()))
    return y


def reflect_pad(x: torch.Tensor, pad: int) -> torch.Tensor:
    x = _chk_dev(x.reshape(-1), torch.float32, "x")
    out = torch.empty(x.numel() + 2 * pad, device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().rvcb_reflect_pad(_p(x), x.numel(), int(pad), _p(out), _stream_ptr()))
    return out


def f32_to_i16(x: torch.Tensor) -> torch.Tensor:
    x = _chk_dev(x.reshape(-1), torch.float32, "x")
    out = torch.empty(x.numel(), device=x.device, dtype=torch.int16)
    _lib.check(_lib.lib().rvcb_f32_to_i16(_p(x), x.numel(), _p(out), _stream_ptr()))
    return out


def f0_post(f0: torch.Tensor, p_len: int, f0_up_key: float, f0_min: float = 50.0, f0_max: float = 1100.0):
    """Device-resident resize + gap fill + key shift + mel quantisation: (pitch i64[p_len], pitchf f32[p_len])."""
    f0 = _chk_dev(f0.reshape(-1), torch.float32, "f0")
    pitch = torch.empty(p_len, device=f0.device, dtype=torch.int64)
    pitchf = torch.empty(p_len, device=f0.device, dtype=torch.float32)
    scratch = torch.empty(2 * p_len, device=f0.device, dtype=torch.float64)
    _lib.check(_lib.lib().rvcb_f0_post(_p(f0), f0.numel(), int(p_len), float(pow(2, f0_up_key / 12)), float(f0_min), float(f0_max),
                                      _p(pitch), _p(pitchf), _p(scratch), _stream_ptr()))
    return pitch, pitchf


class Rmvpe:
    def __init__(self, state_dict: Dict[str, torch.Tensor], device: int = 0):
        _lib.init(device)
        self.h = C.c_void_p()
        with Weights(state_dict) as w:
            _lib.check(_lib.lib().rvcb_rmvpe_create(w.h, C.byref(self.h)))
        self.device = torch.device("cuda", device)

    @staticmethod
    def num_frames(n_samples: int) -> int:
        return _lib.lib().rvcb_rmvpe_num_frames(n_samples)

    def infer(self, wav: torch.Tensor, thred: float = 0.03, want_mel: bool = False, want_hidden: bool = False):
        wav = _chk_dev(wav.reshape(-1), torch.float32, "wav")
        nf = self.num_frames(wav.numel())
        mel = torch.empty(128, nf, device=wav.device) if want_mel else None
        hid = torch.empty(nf, 360, device=wav.device) if want_hidden else None
        f0 = torch.empty(nf, device=wav.device)
        n = C.c_int(0)
        _lib.check(_lib.lib().rvcb_rmvpe_infer(self.h, _p(wav), wav.numel(), float(thred), _p(mel), _p(hid), _p(f0), C.byref(n), _stream_ptr()))
        return f0, mel, hid

    def __del__(self):
        try:
            if self.h:
                _lib.lib().rvcb_rmvpe_destroy(self.h)
        except Exception:
            pass


def synth_config_struct(config, encoder_dim: int) -> "_lib.SynthConfig":
    (_spec, _seg, inter, hidden, filt, n_heads, n_layers, ksz, _pd, _rb, rb_k, rb_d, up_rates, up_init, up_k, n_spk, gin, sr) = config
    if isinstance(sr, str):
        sr = {"32k": 32000, "40k": 40000, "48k": 48000}[sr]
    c = _lib.SynthConfig()
    c.inter_channels, c.hidden_channels, c.filter_channels = inter, hidden, filt
    c.n_heads, c.n_layers, c.kernel_size = n_heads, n_layers, ksz
    c.n_resblock_kernels = len(rb_k)
    for i, k in enumerate(rb_k):
        c.resblock_kernel_sizes[i] = k
        for j, d in enumerate(rb_d[i]):
            c.resblock_dilations[i][j] = d
    c.n_upsamples = len(up_rates)
    for i, (u, k) in enumerate(zip(up_rates, up_k)):
        c.upsample_rates[i] = u
        c.upsample_kernel_sizes[i] = k
    c.upsample_initial_channel, c.spk_embed_dim, c.gin_channels, c.sr, c.encoder_dim = up_init, n_spk, gin, sr, encoder_dim
    return c


class Synth:
    def __init__(self, state_dict: Dict[str, torch.Tensor], config, encoder_dim: int = 768, device: int = 0):
        _lib.init(device)
        self.cfg = synth_config_struct(config, encoder_dim)
        self.upp = 1
        for i in range(self.cfg.n_upsamples):
            self.upp *= self.cfg.upsample_rates[i]
        self.inter = self.cfg.inter_channels
        self.h = C.c_void_p()
        with Weights(state_dict) as w:
            _lib.check(_lib.lib().rvcb_synth_create(C.byref(self.cfg), w.h, C.byref(self.h)))
        self.device = torch.device("cuda", device)

    def infer(self, phone: torch.Tensor, sid: int, pitch: Optional[torch.Tensor], pitchf: Optional[torch.Tensor], noise_prior: torch.Tensor,
              noise_src: Optional[torch.Tensor], skip_head: Optional[int] = None, return_length: Optional[int] = None,
              return_length2: Optional[int] = None) -> torch.Tensor:
        """pitch / pitchf / noise_src are None for no-f0 models (``cpt["f0"] == 0``)."""
        phone = _chk_dev(phone.reshape(-1, phone.shape[-1]), torch.float32, "phone")
        T = phone.shape[0]
        pitch = None if pitch is None else _chk_dev(pitch.reshape(-1), torch.int64, "pitch")
        pitchf = None if pitchf is None else _chk_dev(pitchf.reshape(-1), torch.float32, "pitchf")
        noise_prior = _chk_dev(noise_prior.reshape(self.inter, -1), torch.float32, "noise_prior")
        noise_src = None if noise_src is None else _chk_dev(noise_src.reshape(-1), torch.float32, "noise_src")
        T_dec = T if return_length is None else int(return_length)
        T_out = T_dec if return_length2 is None else int(return_length2)
        out = torch.empty(T_out * self.upp, device=phone.device, dtype=torch.float32)
        n = C.c_int(0)
        _lib.check(_lib.lib().rvcb_synth_infer(self.h, _p(phone), T, int(sid), _p(pitch), _p(pitchf), _p(noise_prior), _p(noise_src),
                                               -1 if skip_head is None else int(skip_head),
                                               -1 if return_length is None else int(return_length),
                                               -1 if return_length2 is None else int(return_length2), _p(out), C.byref(n), _stream_ptr()))
        return out[: n.value]

    def infer_keep(self, phone: torch.Tensor, sid: int, pitch: Optional[torch.Tensor], pitchf: Optional[torch.Tensor], noise_prior: torch.Tensor,
                   noise_src: Optional[torch.Tensor], keep_head: int, keep_length: int) -> torch.Tensor:
        """``infer(...)[keep_head*upp : (keep_head+keep_length)*upp]`` bit for bit, with the flow / decoder restricted to the kept frames
        plus their receptive-field margins (rvcb_synth_infer_keep).  noise tensors are the full-length ones of ``infer``."""
        phone = _chk_dev(phone.reshape(-1, phone.shape[-1]), torch.float32, "phone")
        T = phone.shape[0]
        pitch = None if pitch is None else _chk_dev(pitch.reshape(-1), torch.int64, "pitch")
        pitchf = None if pitchf is None else _chk_dev(pitchf.reshape(-1), torch.float32, "pitchf")
        noise_prior = _chk_dev(noise_prior.reshape(self.inter, -1), torch.float32, "noise_prior")
        noise_src = None if noise_src is None else _chk_dev(noise_src.reshape(-1), torch.float32, "noise_src")
        if noise_prior.shape[1] != T or (noise_src is not None and noise_src.numel() != T * self.upp):
            raise RuntimeError("infer_keep takes the full-length noise tensors of the untrimmed call")
        out = torch.empty(int(keep_length) * self.upp, device=phone.device, dtype=torch.float32)
        n = C.c_int(0)
        _lib.check(_lib.lib().rvcb_synth_infer_keep(self.h, _p(phone), T, int(sid), _p(pitch), _p(pitchf), _p(noise_prior), _p(noise_src),
                                                    int(keep_head), int(keep_length), _p(out), C.byref(n), _stream_ptr()))
        return out[: n.value]

    def __del__(self):
        try:
            if self.h:
                _lib.lib().rvcb_synth_destroy(self.h)
        except Exception:
            pass

"""Drop-in for infer/modules/vc/utils.py:24-36 (load_hubert, index discovery)."""
import os
import pathlib

import torch

from rvc_b200.engine import Hubert


def get_index_path_from_model(sid):
    return next((f for f in [str(pathlib.Path(root, name))
                             for path in [os.getenv("outside_index_root"), os.getenv("index_root")] if path
                             for root, _, files in os.walk(path, topdown=False) for name in files
                             if name.endswith(".index") and "trained" not in name] if sid.split(".")[0] in f), "")


class _Permissive:
    """Stand-in for classes pickled inside a fairseq checkpoint (fairseq is not a dependency here)."""
    def __init__(self, *a, **k): pass
    def __setstate__(self, s): self.__dict__.update(s if isinstance(s, dict) else {})


# Globals a tensor checkpoint legitimately needs.  Everything else pickled inside a fairseq checkpoint (Dictionary, argparse
# Namespace, omegaconf nodes, ...) is mapped to an inert stub UNCONDITIONALLY: find_class never imports on behalf of the file,
# so a crafted hubert_base.pt cannot reach os.system / subprocess / builtins.eval (the reference's inference-side load uses
# torch safe_globals for the same reason, infer/modules/vc/utils.py:24-36).
_ALLOWED_GLOBALS = {
    ("collections", "OrderedDict"),
    ("torch._utils", "_rebuild_tensor_v2"), ("torch._utils", "_rebuild_tensor"), ("torch._utils", "_rebuild_parameter"),
    ("torch._utils", "_rebuild_parameter_with_state"), ("torch._tensor", "_rebuild_from_type_v2"),
    ("torch", "Size"), ("torch", "device"), ("torch", "dtype"),
    ("torch.serialization", "_get_layout"),
    ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
    ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"),
    ("numpy", "ndarray"), ("numpy", "dtype"),
    ("_codecs", "encode"),
}
_ALLOWED_TORCH_NAMES = {n for n in dir(torch) if n.endswith("Storage")} | {
    "float16", "float32", "float64", "bfloat16", "int8", "int16", "int32", "int64", "uint8", "bool"}


def _load_fairseq_state_dict(path: str) -> dict:
    import importlib
    import pickle

    class U(pickle.Unpickler):
        def find_class(self, module, name):
            if (module, name) in _ALLOWED_GLOBALS or (module == "torch" and name in _ALLOWED_TORCH_NAMES):
                return getattr(importlib.import_module(module), name)
            return type(name, (_Permissive,), {})

    class PM:
        Unpickler = U
        load = staticmethod(lambda f, **k: U(f, **k).load())
        __name__ = "pickle"
    ckpt = torch.load(path, map_location="cpu", weights_only=False, pickle_module=PM)
    return ckpt["model"] if isinstance(ckpt, dict) and "model" in ckpt else ckpt


class HubertB200:
    """Duck-type of the fairseq HubertModel object for the calls at pipeline.py:102-110 / rtrvc.py:154-162."""

    def __init__(self, state_dict: dict, device="cuda:0"):
        self.device = torch.device(device if "cuda" in str(device) else "cuda:0")
        self._state = state_dict                       # kept for clone(): a second lane needs its own handle (arenas are per handle)
        self._m = Hubert({k: v.float() for k, v in state_dict.items() if torch.is_tensor(v) and v.is_floating_point()},
                         self.device.index or 0)

    def half(self): return self
    def float(self): return self
    def eval(self): return self
    def to(self, *a, **k): return self

    def clone(self) -> "HubertB200":
        """A second handle over the same weights (own workspace): one per concurrent lane of ``VC.vc_multi``."""
        return HubertB200(self._state, self.device)

    @torch.no_grad()
    def extract_features(self, source, padding_mask=None, mask=False, output_layer=None):
        """source [B, N]; padding_mask bool [B, N] (True = padding, trailing).  B = 1 without padding is the reference callers'
        case (pipeline.py:100-110) and the hot path.  B > 1 is the batched front door: every utterance is run over
        its own valid samples and the results are stacked, zero-padded to the longest -- frame for frame what B = 1 calls return
        (the kernels themselves are B = 1: independent utterances overlap as separate streams / graphs, not as packed GEMMs).
        Returns (features [B, T_max, 768], frame padding mask [B, T_max] or None)."""
        if source.dim() != 2:
            raise ValueError("source must be [B, N]")
        layer = 12 if output_layer is None else int(output_layer)
        B = source.shape[0]
        lens = [source.shape[1]] * B
        if padding_mask is not None and bool(padding_mask.any()):
            pm = padding_mask.to("cpu")
            for b in range(B):
                n_valid = int((~pm[b]).sum())
                if bool(pm[b, :n_valid].any()):
                    raise ValueError("padding_mask must mark trailing padding only")
                lens[b] = n_valid
        if B == 1 and lens[0] == source.shape[1]:
            return self._m.extract(source[0].to(self.device), layer).unsqueeze(0), None
        outs = [self._m.extract(source[b, : lens[b]].to(self.device), layer) for b in range(B)]
        T = max(o.shape[0] for o in outs)
        feats = torch.zeros(B, T, outs[0].shape[1], device=self.device)
        fmask = torch.ones(B, T, dtype=torch.bool, device=self.device)
        for b, o in enumerate(outs):
            feats[b, : o.shape[0]] = o
            fmask[b, : o.shape[0]] = False
        return feats, (fmask if bool(fmask.any()) else None)

    @torch.no_grad()
    def final_proj(self, x):
        return self._m.final_proj(x.reshape(-1, 768)).view(*x.shape[:-1], 256)


def load_hubert(device, is_half, path: str = "assets/hubert/hubert_base.pt"):
    return HubertB200(_load_fairseq_state_dict(path), device)

"""Drop-in for the reference's ``rvc/synthesizer.py`` (get_synthesizer / load_synthesizer,
rvc/synthesizer.py:10-35): same names, arguments and returned ``(net_g, cpt)`` pair, but ``net_g`` is a
thin container around the librvcb200 synthesizer handle -- ``net_g.infer(...)`` keeps the signature of
``SynthesizerTrnMsNSFsid.infer`` (rvc/layers/synthesizers.py:159-170) and runs entirely in hand-written
sm_100a kernels.  There is no PyTorch / CPU fallback: construction fails if the CUDA library is missing.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional

import torch

from rvc_b200.engine import Synth


def _fold_weight_norm(sd):
    """Accept legacy ``weight_g/weight_v`` and ``parametrizations.weight.original{0,1}`` keys (SURVEY
    Appendix C); the C side folds them (csrc/weights.cuh effective_weight), so just pass through as fp32."""
    return {k: v.float() for k, v in sd.items() if torch.is_tensor(v) and v.is_floating_point()}


class SynthesizerB200:
    """Container with the ``nn.Module`` surface the callers touch (.infer/.half/.float/.eval/.to)."""

    def __init__(self, cpt: dict, device="cuda:0"):
        dev = torch.device(device if "cuda" in str(device) else "cuda:0")
        version = cpt.get("version", "v1")
        self.encoder_dim = 256 if version == "v1" else 768
        self.use_f0 = cpt.get("f0", 1) == 1
        self.accepts_host_scalars = True     # phone_lengths / sid may be CPU tensors: reading them then costs no device sync
        self.config = list(cpt["config"])
        self.device = dev
        self._synth = Synth(_fold_weight_norm(cpt["weight"]), self.config, self.encoder_dim, dev.index or 0)
        self.upp = self._synth.upp
        self.inter = self._synth.inter
        self._noise = []

    # nn.Module look-alikes -------------------------------------------------------------------
    def half(self): return self
    def float(self): return self
    def eval(self): return self
    def to(self, *a, **k): return self
    def remove_weight_norm(self): return None

    def set_noise(self, noise_prior: torch.Tensor, noise_src: torch.Tensor):
        """Parity hook (precedent: rvc/onnx/synthesizer.py:66-80 takes ``rnd`` as an input): the next
        ``infer`` consumes these tensors instead of drawing randn (queued: one entry per upcoming ``infer`` call)."""
        self._noise.append((noise_prior, noise_src))

    @torch.no_grad()
    def infer(self, phone: torch.Tensor, phone_lengths: torch.Tensor, sid: torch.Tensor, pitch: Optional[torch.Tensor] = None,
              pitchf: Optional[torch.Tensor] = None, skip_head: Optional[int] = None, return_length: Optional[int] = None,
              return_length2: Optional[int] = None, keep_head: Optional[int] = None, keep_length: Optional[int] = None) -> torch.Tensor:
        """keep_head / keep_length (not in the reference signature; used by the drop-in Pipeline): return only
        ``infer(...)[:, :, keep_head*upp : (keep_head+keep_length)*upp]`` -- bit for bit, see rvcb_synth_infer_keep."""
        if phone.dim() != 3:
            raise ValueError("phone must be [B, T, C]")
        if phone.shape[0] != 1:
            # batched front door: sequence_mask semantics (rvc/layers/utils.py:58-65) by running every utterance over
            # its own phone_lengths[b] frames; outputs are stacked and zero-padded to the longest.  The reference callers are B = 1.
            if skip_head is not None or return_length is not None or return_length2 is not None:
                raise ValueError("the realtime arguments are B = 1 only")
            outs = []
            for b in range(phone.shape[0]):
                Tb = int(phone_lengths.reshape(-1)[b])
                outs.append(self.infer(phone[b: b + 1, :Tb], torch.tensor([Tb]), sid.reshape(-1)[b: b + 1] if sid.numel() > 1 else sid,
                                       None if pitch is None else pitch[b: b + 1, :Tb], None if pitchf is None else pitchf[b: b + 1, :Tb])[0, 0])
            n = max(o.shape[0] for o in outs)
            y = torch.zeros(len(outs), 1, n, device=self.device)
            for b, o in enumerate(outs):
                y[b, 0, : o.shape[0]] = o
            return y
        T = phone.shape[1]
        if int(phone_lengths.reshape(-1)[0]) != T:
            raise ValueError("phone_lengths must equal the number of phone frames")
        if self.use_f0 and (pitch is None or pitchf is None):
            raise ValueError("f0 model needs pitch and pitchf")
        if skip_head is not None and return_length is not None:
            flow_head = max(int(skip_head) - 24, 0)
            Tf, Td = T - flow_head, int(return_length)
        else:
            skip_head = return_length = None
            Tf, Td = T, T
        if self._noise:
            n1, n2 = self._noise.pop(0)
        else:   # synthesizers.py:180,188 randn_like(m_p); generators.py:160,192 rand(1,1,1) + randn_like(sine_waves)
            n1 = torch.randn(1, self.inter, Tf, device=self.device)
            n2 = None
            if self.use_f0:
                torch.rand(1, 1, 1, device=self.device)
                n2 = torch.randn(1, Td * self.upp, 1, device=self.device)
        if keep_head is not None and keep_length is not None and skip_head is None:
            out = self._synth.infer_keep(phone[0].to(self.device), int(sid.reshape(-1)[0]),
                                         pitch.reshape(-1)[:T].to(self.device) if self.use_f0 else None,
                                         pitchf.reshape(-1)[:T].to(self.device) if self.use_f0 else None, n1.to(self.device),
                                         None if n2 is None else n2.to(self.device), int(keep_head), int(keep_length))
            return out.view(1, 1, -1)
        out = self._synth.infer(phone[0].to(self.device), int(sid.reshape(-1)[0]),
                                pitch.reshape(-1)[:T].to(self.device) if self.use_f0 else None,
                                pitchf.reshape(-1)[:T].to(self.device) if self.use_f0 else None, n1.to(self.device),
                                None if n2 is None else n2.to(self.device), skip_head, return_length, return_length2)
        return out.view(1, 1, -1)


def get_synthesizer(cpt: OrderedDict, device=torch.device("cuda:0")):
    cpt["config"][-3] = cpt["weight"]["emb_g.weight"].shape[0]
    net_g = SynthesizerB200(cpt, device)
    return net_g, cpt


def load_synthesizer(pth_path, device=torch.device("cuda:0")):
    return get_synthesizer(torch.load(pth_path, map_location=torch.device("cpu"), weights_only=True), device)

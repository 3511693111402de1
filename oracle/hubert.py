"""Oracle: HuBERT-base ``extract_features`` (fairseq semantics), fp32 CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The arithmetic lives in a third-party dependency that is absent from
/root/reference: ``fairseq @ git+https://github.com/fumiama/fairseq.git``,
no version / commit pin (requirements/main.txt:7).  This file restates the
published HuBERT-base architecture and anchors on the reference's call sites:

  call sites         infer/modules/vc/pipeline.py:102-110, infer/lib/rtrvc.py:154-162,
                     infer/modules/train/extract_feature_print.py:127-140
  encoder loop       rvc/hubert.py:27-91 (the repo's own re-statement of fairseq
                     TransformerEncoder.extract_features: pos_conv add, LN
                     (layer_norm_first=False), pad-to-multiple-of-2, layer loop with
                     early exit at ``tgt_layer``, undo pad)
  conv feature ext.  [(512,10,5)] + [(512,3,2)]*4 + [(512,2,2)]*2, no bias, GroupNorm(512,512)
                     after conv0 only ("default" extractor mode), exact-erf GELU
  pos_conv           Conv1d(768,768,k=128,pad=64,groups=16), weight_norm(dim=2),
                     SamePad drops the last frame, GELU
  layers             post-LN: x = LN(x + MHA(x)); x = LN(x + fc2(gelu(fc1(x))))
                     12 heads x 64, q scaled by 64**-0.5 after the bias add

Parity status: UNPINNED by reference tests for synthetic weights (the only
fixtures, logs/mute/3_feature{256,768}/mute.npy, need the real hubert_base.pt).
Cross-checked against ``transformers.HubertModel(HubertConfig())`` -- an
independent implementation of the same architecture -- in
tests/test_oracle_vs_reference.py.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from .weights import HUBERT_CONV

W = Dict[str, torch.Tensor]


def pos_conv_weight(w: W) -> torch.Tensor:
    """weight_norm with dim=2: w = g * v / ||v||, norm over dims (0,1) per tap."""
    v, g = w["encoder.pos_conv.0.weight_v"], w["encoder.pos_conv.0.weight_g"]
    return g * v / v.norm(dim=(0, 1), keepdim=True)


def conv_feature_extractor(w: W, wav: torch.Tensor) -> torch.Tensor:
    """wav [B, N] -> [B, 512, T_h]"""
    x = wav.unsqueeze(1).to(w["feature_extractor.conv_layers.0.0.weight"].dtype)
    for i, (_c, _k, s) in enumerate(HUBERT_CONV):
        x = F.conv1d(x, w[f"feature_extractor.conv_layers.{i}.0.weight"], None, stride=s)
        if i == 0:
            x = F.group_norm(x, 512, w["feature_extractor.conv_layers.0.2.weight"],
                             w["feature_extractor.conv_layers.0.2.bias"], 1e-5)
        x = F.gelu(x)
    return x


def encoder_layer(w: W, p: str, x: torch.Tensor, n_heads: int = 12) -> torch.Tensor:
    """x [B,T,768] post-LN transformer layer."""
    B, T, C = x.shape
    hd = C // n_heads
    q = F.linear(x, w[p + "self_attn.q_proj.weight"], w[p + "self_attn.q_proj.bias"]) * hd ** -0.5
    k = F.linear(x, w[p + "self_attn.k_proj.weight"], w[p + "self_attn.k_proj.bias"])
    v = F.linear(x, w[p + "self_attn.v_proj.weight"], w[p + "self_attn.v_proj.bias"])
    q = q.view(B, T, n_heads, hd).transpose(1, 2)
    k = k.view(B, T, n_heads, hd).transpose(1, 2)
    v = v.view(B, T, n_heads, hd).transpose(1, 2)
    a = F.softmax(q @ k.transpose(-1, -2), dim=-1) @ v
    a = a.transpose(1, 2).reshape(B, T, C)
    a = F.linear(a, w[p + "self_attn.out_proj.weight"], w[p + "self_attn.out_proj.bias"])
    x = F.layer_norm(x + a, (C,), w[p + "self_attn_layer_norm.weight"], w[p + "self_attn_layer_norm.bias"], 1e-5)
    h = F.linear(F.gelu(F.linear(x, w[p + "fc1.weight"], w[p + "fc1.bias"])), w[p + "fc2.weight"], w[p + "fc2.bias"])
    return F.layer_norm(x + h, (C,), w[p + "final_layer_norm.weight"], w[p + "final_layer_norm.bias"], 1e-5)


def extract_features(w: W, source: torch.Tensor, output_layer: int = 12,
                     taps: Optional[dict] = None) -> torch.Tensor:
    """fairseq HubertModel.extract_features(source, padding_mask=all-False, mask=False,
    output_layer=L)[0]: [B, T_h, 768] (no final_proj; v1 callers apply it,
    pipeline.py:110)."""
    f = conv_feature_extractor(w, source)
    x = f.transpose(1, 2)
    x = F.layer_norm(x, (512,), w["layer_norm.weight"], w["layer_norm.bias"], 1e-5)
    x = F.linear(x, w["post_extract_proj.weight"], w["post_extract_proj.bias"])
    if taps is not None:
        taps["proj"] = x
    pc = F.conv1d(x.transpose(1, 2), pos_conv_weight(w), w["encoder.pos_conv.0.bias"], padding=64, groups=16)
    pc = F.gelu(pc[:, :, :-1]).transpose(1, 2)
    x = x + pc
    x = F.layer_norm(x, (768,), w["encoder.layer_norm.weight"], w["encoder.layer_norm.bias"], 1e-5)
    if taps is not None:
        taps["enc_in"] = x
    # rvc/hubert.py:45-52 pads T to a multiple of 2 with a key-padding-masked zero frame;
    # a masked key never influences the real frames and the padded row is dropped
    # (rvc/hubert.py:79-80), so the computation over the real frames is unchanged.
    for i in range(output_layer):
        x = encoder_layer(w, f"encoder.layers.{i}.", x)
        if taps is not None:
            taps[f"layer{i}"] = x
    return x


def final_proj(w: W, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, w["final_proj.weight"], w["final_proj.bias"])


def n_frames(n_samples: int) -> int:
    n = n_samples
    for (_c, k, s) in HUBERT_CONV:
        n = (n - k) // s + 1
    return n

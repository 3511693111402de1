"""Seeded synthetic checkpoints / audio / index vectors with the reference's state_dict key names and shapes.

Pure data generation (torch CPU RNG, no arithmetic of the hot path): shared by the tests, the oracle and
bench.py because there are no model assets in
the build container nor on the GPU box, so every parity test
and the bench use weights generated here from a seed with torch's CPU
generator (deterministic for a fixed torch version).  Key names and shapes
follow the reference containers:

  * synthesizer: ``cpt["weight"]`` of rvc/synthesizer.py:10-28 after
    ``remove_weight_norm``; config = the 18-list of
    rvc/layers/synthesizers.py:19-38 (configs/v2/48k.json).
  * RMVPE: plain ``E2E(4,1,(2,2))`` state_dict (rvc/f0/models.py:9-11).
  * HuBERT: fairseq ``HubertModel`` parameter names.

Values are fp16-representable, like the tensors a real ``.pth`` stores
(infer/lib/train/process_ckpt.py:22 saves ``.half()``).
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch

V2_48K_CONFIG = [1025, 32, 192, 192, 768, 2, 6, 3, 0, "1", [3, 7, 11],
                 [[1, 3, 5], [1, 3, 5], [1, 3, 5]], [12, 10, 2, 2], 512,
                 [24, 20, 4, 4], 109, 256, 48000]
V1_40K_CONFIG = [1025, 32, 192, 192, 768, 2, 6, 3, 0, "1", [3, 7, 11],
                 [[1, 3, 5], [1, 3, 5], [1, 3, 5]], [10, 10, 2, 2], 512,
                 [16, 16, 4, 4], 109, 256, 40000]


# the other decoder schedules of configs/{v1,v2}/*.json: 5-stage decoders (last stage 16 channels), kernel 16 at stride 4 / 6
V1_48K_CONFIG = [1025, 32, 192, 192, 768, 2, 6, 3, 0, "1", [3, 7, 11],
                 [[1, 3, 5], [1, 3, 5], [1, 3, 5]], [10, 6, 2, 2, 2], 512,
                 [16, 16, 4, 4, 4], 109, 256, 48000]
V1_32K_CONFIG = [513, 32, 192, 192, 768, 2, 6, 3, 0, "1", [3, 7, 11],
                 [[1, 3, 5], [1, 3, 5], [1, 3, 5]], [10, 4, 2, 2, 2], 512,
                 [16, 16, 4, 4, 4], 109, 256, 32000]
V2_32K_CONFIG = [513, 32, 192, 192, 768, 2, 6, 3, 0, "1", [3, 7, 11],
                 [[1, 3, 5], [1, 3, 5], [1, 3, 5]], [10, 8, 2, 2], 512,
                 [20, 16, 4, 4], 109, 256, 32000]


def _r16(t: torch.Tensor) -> torch.Tensor:
    return t.half().float()


class _Gen:
    def __init__(self, seed: int):
        self.g = torch.Generator().manual_seed(seed)

    def n(self, shape, std=1.0, mean=0.0):
        return _r16(torch.randn(shape, generator=self.g) * std + mean)

    def u(self, shape, lo, hi):
        return _r16(torch.rand(shape, generator=self.g) * (hi - lo) + lo)


# --------------------------------------------------------------------------
# synthesizer  (SynthesizerTrnMs{256,768}NSFsid, weight-norm removed)
# --------------------------------------------------------------------------
def synth_weights(seed: int = 1234, config: List = V2_48K_CONFIG,
                  encoder_dim: int = 768, use_f0: bool = True) -> Dict[str, torch.Tensor]:
    (spec, seg, inter, hidden, filt, n_heads, n_layers, ksz, _pd, _rb, rb_k, rb_d,
     up_rates, up_init, up_k, n_spk, gin, _sr) = config
    G = _Gen(seed)
    w: Dict[str, torch.Tensor] = {}
    kc = hidden // n_heads
    # enc_p  (rvc/layers/encoders.py:84-159)
    w["enc_p.emb_phone.weight"] = G.n((hidden, encoder_dim), 1.0 / math.sqrt(encoder_dim * hidden) * 2.0)
    w["enc_p.emb_phone.bias"] = G.n((hidden,), 0.01)
    w["enc_p.emb_pitch.weight"] = G.n((256, hidden), 1.0 / math.sqrt(hidden))
    for i in range(n_layers):
        p = f"enc_p.encoder.attn_layers.{i}."
        for nm in ("conv_q", "conv_k", "conv_v", "conv_o"):
            w[p + nm + ".weight"] = G.n((hidden, hidden, 1), 1.0 / math.sqrt(hidden))
            w[p + nm + ".bias"] = G.n((hidden,), 0.05)
        w[p + "emb_rel_k"] = G.n((1, 21, kc), kc ** -0.5)
        w[p + "emb_rel_v"] = G.n((1, 21, kc), kc ** -0.5)
        for nl in ("norm_layers_1", "norm_layers_2"):
            w[f"enc_p.encoder.{nl}.{i}.gamma"] = G.n((hidden,), 0.1, 1.0)
            w[f"enc_p.encoder.{nl}.{i}.beta"] = G.n((hidden,), 0.1)
        p = f"enc_p.encoder.ffn_layers.{i}."
        w[p + "conv_1.weight"] = G.n((filt, hidden, ksz), 1.0 / math.sqrt(hidden * ksz))
        w[p + "conv_1.bias"] = G.n((filt,), 0.05)
        w[p + "conv_2.weight"] = G.n((hidden, filt, ksz), 1.0 / math.sqrt(filt * ksz))
        w[p + "conv_2.bias"] = G.n((hidden,), 0.05)
    pw = G.n((2 * inter, hidden, 1), 1.0 / math.sqrt(hidden))
    pw[inter:] = _r16(pw[inter:] * 0.3)       # keep exp(logs) tame
    w["enc_p.proj.weight"] = pw
    pb = G.n((2 * inter,), 0.05)
    pb[inter:] = _r16(pb[inter:] - 1.0)
    w["enc_p.proj.bias"] = pb
    # flow  (rvc/layers/residuals.py:145-330, norms.py:27-152); only even
    # indices hold parameters (odd entries are Flip modules)
    half = inter // 2
    for f in (0, 2, 4, 6):
        p = f"flow.flows.{f}."
        w[p + "pre.weight"] = G.n((hidden, half, 1), 1.0 / math.sqrt(half))
        w[p + "pre.bias"] = G.n((hidden,), 0.05)
        w[p + "enc.cond_layer.weight"] = G.n((2 * hidden * 3, gin, 1), 0.5 / math.sqrt(gin))
        w[p + "enc.cond_layer.bias"] = G.n((2 * hidden * 3,), 0.05)
        for l in range(3):
            w[p + f"enc.in_layers.{l}.weight"] = G.n((2 * hidden, hidden, 5), 1.0 / math.sqrt(hidden * 5))
            w[p + f"enc.in_layers.{l}.bias"] = G.n((2 * hidden,), 0.05)
            rs = 2 * hidden if l < 2 else hidden
            w[p + f"enc.res_skip_layers.{l}.weight"] = G.n((rs, hidden, 1), 1.0 / math.sqrt(hidden))
            w[p + f"enc.res_skip_layers.{l}.bias"] = G.n((rs,), 0.05)
        w[p + "post.weight"] = G.n((half, hidden, 1), 0.5 / math.sqrt(hidden))
        w[p + "post.bias"] = G.n((half,), 0.05)
    # dec  (rvc/layers/nsf.py:64-143)
    w["dec.m_source.l_linear.weight"] = _r16(torch.tensor([[1.25]]))
    w["dec.m_source.l_linear.bias"] = _r16(torch.tensor([0.01]))
    w["dec.conv_pre.weight"] = G.n((up_init, inter, 7), 1.0 / math.sqrt(inter * 7))
    w["dec.conv_pre.bias"] = G.n((up_init,), 0.05)
    w["dec.cond.weight"] = G.n((up_init, gin, 1), 0.5 / math.sqrt(gin))
    w["dec.cond.bias"] = G.n((up_init,), 0.05)
    ch = up_init
    for i, (u, k) in enumerate(zip(up_rates, up_k)):
        cin, cout = up_init // (2 ** i), up_init // (2 ** (i + 1))
        w[f"dec.ups.{i}.weight"] = G.n((cin, cout, k), 1.0 / math.sqrt(cin * k / u))
        w[f"dec.ups.{i}.bias"] = G.n((cout,), 0.05)
        if i + 1 < len(up_rates):
            s = math.prod(up_rates[i + 1:])
            w[f"dec.noise_convs.{i}.weight"] = G.n((cout, 1, 2 * s), 2.0 / math.sqrt(2 * s))
        else:
            w[f"dec.noise_convs.{i}.weight"] = G.n((cout, 1, 1), 2.0)
        w[f"dec.noise_convs.{i}.bias"] = G.n((cout,), 0.05)
        ch = cout
        for j, (k2, dil) in enumerate(zip(rb_k, rb_d)):
            r = i * len(rb_k) + j
            for c in range(len(dil)):
                for nm in ("convs1", "convs2"):
                    w[f"dec.resblocks.{r}.{nm}.{c}.weight"] = G.n((ch, ch, k2), 0.9 / math.sqrt(ch * k2))
                    w[f"dec.resblocks.{r}.{nm}.{c}.bias"] = G.n((ch,), 0.05)
    w["dec.conv_post.weight"] = G.n((1, ch, 7), 0.35 / math.sqrt(ch * 7))
    w["emb_g.weight"] = G.n((n_spk, gin), 0.5)
    if not use_f0:      # SynthesizerTrnMs*NSFsid_nono: TextEncoder without emb_pitch, plain Generator decoder (generators.py:14-113)
        for k in [k for k in w if k.startswith("enc_p.emb_pitch") or k.startswith("dec.m_source") or k.startswith("dec.noise_convs")]:
            del w[k]
    return w


def is_weight_normed(key: str) -> bool:
    """Layers wrapped in weight_norm by the reference (nsf.py:92-101, residuals.py:33-60,
    norms.py:55-84)."""
    return (key.startswith("dec.ups.") or key.startswith("dec.resblocks.")
            or ".enc.cond_layer." in key or ".enc.in_layers." in key or ".enc.res_skip_layers." in key)


def synth_cpt(seed: int = 1234, version: str = "v2", config: List = None, f0: int = 1) -> dict:
    """A dict shaped like the small inference ``.pth``
    (infer/lib/train/process_ckpt.py:15-54)."""
    if config is None:
        config = V2_48K_CONFIG if version == "v2" else V1_40K_CONFIG
    enc = 768 if version == "v2" else 256
    wt = synth_weights(seed, config, enc, use_f0=(f0 == 1))
    sd = {}
    for k, v in wt.items():
        if k.endswith(".weight") and is_weight_normed(k):
            # legacy weight-norm keys, as real checkpoints store them; g is kept in
            # fp32 so that g*v/||v|| reproduces the fp16-representable weight.
            base = k[: -len("weight")]
            sd[base + "weight_v"] = v.half()
            sd[base + "weight_g"] = v.flatten(1).norm(dim=1).view(-1, *([1] * (v.dim() - 1)))
        else:
            sd[k] = v.half()
    return {"weight": sd, "config": list(config),
            "info": "synthetic", "sr": {48000: "48k", 40000: "40k", 32000: "32k"}[config[-1]],
            "f0": f0, "version": version}


# --------------------------------------------------------------------------
# RMVPE  E2E(4, 1, (2,2))   (rvc/f0/e2e.py, rvc/f0/deepunet.py)
# --------------------------------------------------------------------------
def _bn(G: _Gen, w, p, c):
    w[p + "weight"] = G.u((c,), 0.6, 1.2)
    w[p + "bias"] = G.n((c,), 0.1)
    w[p + "running_mean"] = G.n((c,), 0.1)
    w[p + "running_var"] = G.u((c,), 0.6, 1.4)
    w[p + "num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def _conv_block_res(G: _Gen, w, p, cin, cout):
    w[p + "conv.0.weight"] = G.n((cout, cin, 3, 3), math.sqrt(2.0 / (cin * 9)) * 0.8)
    _bn(G, w, p + "conv.1.", cout)
    w[p + "conv.3.weight"] = G.n((cout, cout, 3, 3), math.sqrt(2.0 / (cout * 9)) * 0.3)
    _bn(G, w, p + "conv.4.", cout)
    if cin != cout:
        w[p + "shortcut.weight"] = G.n((cout, cin, 1, 1), math.sqrt(1.0 / cin) * 0.8)
        w[p + "shortcut.bias"] = G.n((cout,), 0.05)


def rmvpe_weights(seed: int = 4321, n_blocks: int = 4, en_de_layers: int = 5,
                  inter_layers: int = 4, en_out: int = 16) -> Dict[str, torch.Tensor]:
    G = _Gen(seed)
    w: Dict[str, torch.Tensor] = {}
    _bn(G, w, "unet.encoder.bn.", 1)
    cin, cout = 1, en_out
    for l in range(en_de_layers):
        for b in range(n_blocks):
            _conv_block_res(G, w, f"unet.encoder.layers.{l}.conv.{b}.", cin if b == 0 else cout, cout)
        cin, cout = cout, cout * 2
    # intermediate: in = enc_out_channel//2, out = enc_out_channel
    cin, cout = cout // 2, cout
    for l in range(inter_layers):
        for b in range(n_blocks):
            _conv_block_res(G, w, f"unet.intermediate.layers.{l}.conv.{b}.",
                            (cin if l == 0 else cout) if b == 0 else cout, cout)
    cin = cout
    for l in range(en_de_layers):
        cout = cin // 2
        p = f"unet.decoder.layers.{l}."
        w[p + "conv1.0.weight"] = G.n((cin, cout, 3, 3), math.sqrt(2.0 / (cin * 9 / 4)))
        _bn(G, w, p + "conv1.1.", cout)
        for b in range(n_blocks):
            _conv_block_res(G, w, p + f"conv2.{b}.", cout * 2 if b == 0 else cout, cout)
        cin = cout
    w["cnn.weight"] = G.n((3, en_out, 3, 3), math.sqrt(1.0 / (en_out * 9)))
    w["cnn.bias"] = G.n((3,), 0.05)
    H, I = 256, 384
    for sfx in ("", "_reverse"):
        w["fc.0.gru.weight_ih_l0" + sfx] = G.n((3 * H, I), 1.0 / math.sqrt(I))
        w["fc.0.gru.weight_hh_l0" + sfx] = G.n((3 * H, H), 1.0 / math.sqrt(H))
        w["fc.0.gru.bias_ih_l0" + sfx] = G.n((3 * H,), 0.05)
        w["fc.0.gru.bias_hh_l0" + sfx] = G.n((3 * H,), 0.05)
    w["fc.1.weight"] = G.n((360, 512), 2.0 / math.sqrt(512))
    w["fc.1.bias"] = G.n((360,), 0.05, -1.0)
    return w


# --------------------------------------------------------------------------
# HuBERT-base (fairseq parameter names)
# --------------------------------------------------------------------------
HUBERT_CONV = [(512, 10, 5)] + [(512, 3, 2)] * 4 + [(512, 2, 2)] * 2


def hubert_weights(seed: int = 777, n_layers: int = 12, with_final_proj: bool = True
                   ) -> Dict[str, torch.Tensor]:
    G = _Gen(seed)
    w: Dict[str, torch.Tensor] = {}
    cin = 1
    for i, (c, k, _s) in enumerate(HUBERT_CONV):
        w[f"feature_extractor.conv_layers.{i}.0.weight"] = G.n((c, cin, k), math.sqrt(2.0 / (cin * k)))
        cin = c
    w["feature_extractor.conv_layers.0.2.weight"] = G.n((512,), 0.1, 1.0)
    w["feature_extractor.conv_layers.0.2.bias"] = G.n((512,), 0.1)
    w["layer_norm.weight"] = G.n((512,), 0.1, 1.0)
    w["layer_norm.bias"] = G.n((512,), 0.1)
    w["post_extract_proj.weight"] = G.n((768, 512), 1.0 / math.sqrt(512))
    w["post_extract_proj.bias"] = G.n((768,), 0.05)
    w["encoder.pos_conv.0.weight_g"] = G.u((1, 1, 128), 1.5, 2.5)
    w["encoder.pos_conv.0.weight_v"] = G.n((768, 48, 128), 1.0)
    w["encoder.pos_conv.0.bias"] = G.n((768,), 0.05)
    w["encoder.layer_norm.weight"] = G.n((768,), 0.1, 1.0)
    w["encoder.layer_norm.bias"] = G.n((768,), 0.1)
    for i in range(n_layers):
        p = f"encoder.layers.{i}."
        for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
            w[p + f"self_attn.{nm}.weight"] = G.n((768, 768), 1.0 / math.sqrt(768))
            w[p + f"self_attn.{nm}.bias"] = G.n((768,), 0.05)
        w[p + "self_attn_layer_norm.weight"] = G.n((768,), 0.1, 1.0)
        w[p + "self_attn_layer_norm.bias"] = G.n((768,), 0.1)
        w[p + "fc1.weight"] = G.n((3072, 768), 1.0 / math.sqrt(768))
        w[p + "fc1.bias"] = G.n((3072,), 0.05)
        w[p + "fc2.weight"] = G.n((768, 3072), 1.0 / math.sqrt(3072))
        w[p + "fc2.bias"] = G.n((768,), 0.05)
        w[p + "final_layer_norm.weight"] = G.n((768,), 0.1, 1.0)
        w[p + "final_layer_norm.bias"] = G.n((768,), 0.1)
    if with_final_proj:
        w["final_proj.weight"] = G.n((256, 768), 1.0 / math.sqrt(768))
        w["final_proj.bias"] = G.n((256,), 0.05)
    return w


# --------------------------------------------------------------------------
# synthetic audio / index  (configs #1/#2)
# --------------------------------------------------------------------------
def synth_voice(seconds: float = 10.0, sr: int = 16000, seed: int = 0) -> torch.Tensor:
    """Sum of 8 harmonics of a 110-330 Hz glide + 1e-3 white noise, peak 0.95,
    with two short silences so the voiced/unvoiced logic is exercised."""
    g = torch.Generator().manual_seed(seed)
    n = int(seconds * sr)
    t = torch.arange(n, dtype=torch.float64) / sr
    f0 = 110.0 + 220.0 * (0.5 - 0.5 * torch.cos(2 * math.pi * t / max(seconds, 1e-9) * 1.5))
    ph = 2 * math.pi * torch.cumsum(f0, 0) / sr
    x = torch.zeros(n, dtype=torch.float64)
    for h in range(1, 9):
        x += torch.sin(h * ph) / h
    env = torch.ones(n, dtype=torch.float64)
    for a, b in ((0.30, 0.34), (0.71, 0.74)):
        env[int(a * n): int(b * n)] = 0.0
    x = x * env + 1e-3 * torch.randn(n, generator=g, dtype=torch.float64)
    x = x / x.abs().max() * 0.95
    return x.float()


def index_vectors(n: int = 100000, d: int = 768, seed: int = 0, scale: float = 0.22) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, d, generator=g) * scale

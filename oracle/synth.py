"""Oracle: SynthesizerTrnMs{256,768}NSFsid.infer, fp32 CPU restatement.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Functional PyTorch fp32,
weights in a dict keyed by the reference's state_dict names (weight-norm
already folded, rvc/synthesizer.py:27).  Noise is an explicit input
(precedent: rvc/onnx/synthesizer.py:66-80), because the reference draws it
with randn_like (rvc/layers/synthesizers.py:180,188; generators.py:192).

Reference sites restated:
  TextEncoder.forward          rvc/layers/encoders.py:135-159
  Encoder.forward              rvc/layers/encoders.py:64-81
  MultiHeadAttention           rvc/layers/attentions.py:71-214
  FFN                          rvc/layers/attentions.py:260-314
  LayerNorm (over channels)    rvc/layers/norms.py:12-24
  ResidualCouplingBlock (rev)  rvc/layers/residuals.py:210-235, 311-324
  WN                           rvc/layers/norms.py:93-124
  gate                         rvc/layers/utils.py:47-55
  SineGenerator / source       rvc/layers/generators.py:148-194, nsf.py:57-61
  NSFGenerator.forward         rvc/layers/nsf.py:145-191
  ResBlock1.forward            rvc/layers/residuals.py:68-85
  infer (skip_head variant)    rvc/layers/synthesizers.py:159-203
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

W = Dict[str, torch.Tensor]


def _ln_c(x, gamma, beta, eps=1e-5):
    """LayerNorm over the channel dim of [B,C,T] (norms.py:21-24)."""
    return F.layer_norm(x.transpose(1, -1), (x.shape[1],), gamma, beta, eps).transpose(1, -1)


def rel_attention(w: W, p: str, x: torch.Tensor, mask: torch.Tensor, n_heads: int, window: int = 10):
    """attentions.py:86-142 with the pad/reshape skew written as an explicit band:
    S[i,j] += q_i . E_k[j-i+w]  and  O[i] += sum_{|j-i|<=w} P[i,j] E_v[j-i+w]."""
    B, C, T = x.shape
    kc = C // n_heads
    q = F.conv1d(x, w[p + "conv_q.weight"], w[p + "conv_q.bias"])
    k = F.conv1d(x, w[p + "conv_k.weight"], w[p + "conv_k.bias"])
    v = F.conv1d(x, w[p + "conv_v.weight"], w[p + "conv_v.bias"])
    q = q.view(B, n_heads, kc, T).transpose(2, 3) / math.sqrt(kc)
    k = k.view(B, n_heads, kc, T).transpose(2, 3)
    v = v.view(B, n_heads, kc, T).transpose(2, 3)
    scores = q @ k.transpose(-2, -1)                       # [B,H,T,T]
    ek = w[p + "emb_rel_k"][0]                             # [2w+1, kc]
    ev = w[p + "emb_rel_v"][0]
    rel = q @ ek.t()                                       # [B,H,T,2w+1]
    idx = torch.arange(T, device=x.device)
    d = idx[None, :] - idx[:, None]                        # j - i
    band = d.abs() <= window
    di = (d + window).clamp(0, 2 * window)
    scores = scores + torch.where(band, rel.gather(-1, di.expand(B, n_heads, T, T)), torch.zeros((), device=x.device, dtype=rel.dtype))
    am = mask.unsqueeze(2) * mask.unsqueeze(-1)            # [B,1,T,T]
    scores = scores.masked_fill(am == 0, -1e4)
    pr = F.softmax(scores, dim=-1)
    out = pr @ v
    pband = torch.where(band, pr, torch.zeros((), device=x.device, dtype=pr.dtype))         # [B,H,T,T]
    relw = torch.zeros(B, n_heads, T, 2 * window + 1, device=x.device, dtype=pr.dtype)
    relw.scatter_add_(-1, di.expand(B, n_heads, T, T), pband)
    out = out + relw @ ev
    out = out.transpose(2, 3).contiguous().view(B, C, T)
    return F.conv1d(out, w[p + "conv_o.weight"], w[p + "conv_o.bias"])


def text_encoder(w: W, phone, pitch, lengths, n_heads=2, n_layers=6, ksz=3, skip_head: Optional[int] = None):
    hidden = w["enc_p.emb_phone.weight"].shape[0]
    x = F.linear(phone, w["enc_p.emb_phone.weight"], w["enc_p.emb_phone.bias"])
    if pitch is not None:
        x = x + F.embedding(pitch, w["enc_p.emb_pitch.weight"])
    x = x * math.sqrt(hidden)
    x = F.leaky_relu(x, 0.1)
    x = x.transpose(1, -1)
    T = x.shape[2]
    mask = (torch.arange(T, device=x.device)[None, :] < lengths.to(x.device)[:, None]).unsqueeze(1).to(x.dtype)
    x = x * mask
    x = x * mask   # Encoder.forward re-applies (encoders.py:66)
    for i in range(n_layers):
        y = rel_attention(w, f"enc_p.encoder.attn_layers.{i}.", x, mask, n_heads)
        x = _ln_c(x + y, w[f"enc_p.encoder.norm_layers_1.{i}.gamma"], w[f"enc_p.encoder.norm_layers_1.{i}.beta"])
        p = f"enc_p.encoder.ffn_layers.{i}."
        pl, pr_ = (ksz - 1) // 2, ksz // 2
        y = F.conv1d(F.pad(x * mask, (pl, pr_)), w[p + "conv_1.weight"], w[p + "conv_1.bias"])
        y = torch.relu(y)
        y = F.conv1d(F.pad(y * mask, (pl, pr_)), w[p + "conv_2.weight"], w[p + "conv_2.bias"]) * mask
        x = _ln_c(x + y, w[f"enc_p.encoder.norm_layers_2.{i}.gamma"], w[f"enc_p.encoder.norm_layers_2.{i}.beta"])
    x = x * mask
    if skip_head is not None:
        x = x[:, :, int(skip_head):]
        mask = mask[:, :, int(skip_head):]
    stats = F.conv1d(x, w["enc_p.proj.weight"], w["enc_p.proj.bias"]) * mask
    out = stats.shape[1] // 2
    return stats[:, :out], stats[:, out:], mask


def wn(w: W, p: str, x, mask, g, hidden: int, n_layers: int = 3, ksz: int = 5):
    out = torch.zeros_like(x)
    gc = F.conv1d(g, w[p + "cond_layer.weight"], w[p + "cond_layer.bias"])
    for i in range(n_layers):
        xin = F.conv1d(x, w[p + f"in_layers.{i}.weight"], w[p + f"in_layers.{i}.bias"], padding=(ksz - 1) // 2)
        a = xin + gc[:, i * 2 * hidden:(i + 1) * 2 * hidden]
        acts = torch.tanh(a[:, :hidden]) * torch.sigmoid(a[:, hidden:])
        rs = F.conv1d(acts, w[p + f"res_skip_layers.{i}.weight"], w[p + f"res_skip_layers.{i}.bias"])
        if i < n_layers - 1:
            x = (x + rs[:, :hidden]) * mask
            out = out + rs[:, hidden:]
        else:
            out = out + rs
    return out * mask


def flow_reverse(w: W, z, mask, g, hidden: int = 192, n_flows: int = 4):
    half = z.shape[1] // 2
    for f in reversed(range(n_flows)):
        z = torch.flip(z, [1])
        p = f"flow.flows.{2 * f}."
        x0, x1 = z[:, :half], z[:, half:]
        h = F.conv1d(x0, w[p + "pre.weight"], w[p + "pre.bias"]) * mask
        h = wn(w, p + "enc.", h, mask, g, hidden)
        m = F.conv1d(h, w[p + "post.weight"], w[p + "post.bias"]) * mask
        x1 = (x1 - m) * mask                      # mean_only: logs == 0
        z = torch.cat([x0, x1], 1)
    return z


def sine_source(w: W, f0: torch.Tensor, upp: int, sr: int, noise: torch.Tensor):
    """generators.py:148-194 + nsf.py:57-61.  f0 [B,T]; noise [B,T*upp,1] ~ N(0,1).
    Returns har_source [B,1,T*upp]."""
    f0 = f0.unsqueeze(-1)
    a = torch.arange(1, upp + 1, dtype=f0.dtype, device=f0.device)
    rad = f0 / sr * a                                         # [B,T,upp]
    rad2 = torch.fmod(rad[:, :-1, -1:].float() + 0.5, 1.0) - 0.5
    rad_acc = rad2.cumsum(dim=1).fmod(1.0).to(f0)
    rad = rad + F.pad(rad_acc, (0, 0, 1, 0))
    rad = rad.reshape(f0.shape[0], -1, 1)
    sines = torch.sin(2 * torch.pi * rad) * 0.1
    uv = (f0 > 0).to(f0.dtype)
    uv = F.interpolate(uv.transpose(2, 1), scale_factor=float(upp), mode="nearest").transpose(2, 1)
    noise_amp = uv * 0.003 + (1 - uv) * 0.1 / 3
    sw = sines * uv + noise_amp * noise
    lw = w["dec.m_source.l_linear.weight"]
    har = torch.tanh(F.linear(sw.to(lw.dtype), lw, w["dec.m_source.l_linear.bias"]))
    return har.transpose(1, 2)


def nsf_generator(w: W, x, f0, g, noise, up_rates: List[int], up_k: List[int], rb_k: List[int],
                  rb_d: List[List[int]], sr: int, n_res: Optional[int] = None, taps: Optional[dict] = None):
    upp = math.prod(up_rates)
    use_f0 = f0 is not None and "dec.noise_convs.0.weight" in w      # else: plain Generator (generators.py:84-113)
    har = sine_source(w, f0, upp, sr, noise) if use_f0 else None
    if n_res is not None:
        n_res = int(n_res)
        if use_f0 and n_res * upp != har.shape[-1]:
            har = F.interpolate(har, size=n_res * upp, mode="linear")
        if n_res != x.shape[-1]:
            x = F.interpolate(x, size=n_res, mode="linear")
    if taps is not None and use_f0:
        taps["har"] = har
    x = F.conv1d(x, w["dec.conv_pre.weight"], w["dec.conv_pre.bias"], padding=3)
    x = x + F.conv1d(g, w["dec.cond.weight"], w["dec.cond.bias"])
    nk = len(rb_k)
    for i, (u, k) in enumerate(zip(up_rates, up_k)):
        x = F.leaky_relu(x, 0.1)
        x = F.conv_transpose1d(x, w[f"dec.ups.{i}.weight"], w[f"dec.ups.{i}.bias"], stride=u, padding=(k - u) // 2)
        if use_f0:
            if i + 1 < len(up_rates):
                s = math.prod(up_rates[i + 1:])
                xs_ = F.conv1d(har, w[f"dec.noise_convs.{i}.weight"], w[f"dec.noise_convs.{i}.bias"], stride=s, padding=s // 2)
            else:
                xs_ = F.conv1d(har, w[f"dec.noise_convs.{i}.weight"], w[f"dec.noise_convs.{i}.bias"])
            x = x + xs_
        acc = None
        for j in range(nk):
            r = i * nk + j
            y = x
            for c, d in enumerate(rb_d[j]):
                t = F.leaky_relu(y, 0.1)
                t = F.conv1d(t, w[f"dec.resblocks.{r}.convs1.{c}.weight"], w[f"dec.resblocks.{r}.convs1.{c}.bias"],
                             dilation=d, padding=(rb_k[j] * d - d) // 2)
                t = F.leaky_relu(t, 0.1)
                t = F.conv1d(t, w[f"dec.resblocks.{r}.convs2.{c}.weight"], w[f"dec.resblocks.{r}.convs2.{c}.bias"],
                             padding=(rb_k[j] - 1) // 2)
                y = t + y
            acc = y if acc is None else acc + y
        x = acc / nk
        if taps is not None:
            taps[f"stage{i}"] = x
    x = F.leaky_relu(x)            # default slope 0.01 (nsf.py:187)
    x = F.conv1d(x, w["dec.conv_post.weight"], None, padding=3)
    return torch.tanh(x)


def synth_infer(w: W, config: List, phone, phone_lengths, sid, pitch, pitchf, noise_prior, noise_src,
                skip_head: Optional[int] = None, return_length: Optional[int] = None,
                return_length2: Optional[int] = None, taps: Optional[dict] = None):
    """synthesizers.py:159-203.  noise_prior: [B,inter,T'] (T' = T - flow_head for the
    realtime variant), noise_src: [B, T_dec*upp, 1]."""
    (_spec, _seg, inter, hidden, _filt, n_heads, n_layers, ksz, _pd, _rb, rb_k, rb_d,
     up_rates, _up_init, up_k, _n_spk, _gin, sr) = config
    if isinstance(sr, str):
        sr = {"32k": 32000, "40k": 40000, "48k": 48000}[sr]
    g = F.embedding(sid, w["emb_g.weight"]).unsqueeze(-1)
    if skip_head is not None and return_length is not None:
        head, length = int(skip_head), int(return_length)
        flow_head = max(head - 24, 0)
        dec_head = head - flow_head
        m_p, logs_p, mask = text_encoder(w, phone, pitch, phone_lengths, n_heads, n_layers, ksz, flow_head)
        z_p = (m_p + torch.exp(logs_p) * noise_prior * 0.66666) * mask
        z = flow_reverse(w, z_p, mask, g, hidden)
        z = z[:, :, dec_head:dec_head + length]
        mask = mask[:, :, dec_head:dec_head + length]
        if pitchf is not None:
            pitchf = pitchf[:, head:head + length]
    else:
        m_p, logs_p, mask = text_encoder(w, phone, pitch, phone_lengths, n_heads, n_layers, ksz)
        z_p = (m_p + torch.exp(logs_p) * noise_prior * 0.66666) * mask
        z = flow_reverse(w, z_p, mask, g, hidden)
    if taps is not None:
        taps.update(m_p=m_p, logs_p=logs_p, z_p=z_p, z=z)
    return nsf_generator(w, z * mask, pitchf, g, noise_src, up_rates, up_k, rb_k, rb_d, sr,
                         n_res=return_length2, taps=taps)

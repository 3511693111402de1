"""GPU parity: HuBERT content features (hand-written sm_100a path) vs the fp32 CPU oracle."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seconds,layer", [(1.0, 12), (2.53, 12), (1.0, 9)])
def test_hubert_features_match_oracle(seconds, layer):
    from oracle import hubert as OH, weights as OW
    from rvc_b200.engine import Hubert
    w = OW.hubert_weights(777)
    wav = OW.synth_voice(seconds, seed=3)
    with torch.no_grad():
        ref = OH.extract_features(w, wav[None], layer)[0]
    m = Hubert(w)
    out = m.extract(wav.cuda(), layer).cpu()
    assert out.shape == ref.shape
    err = (out - ref).abs().max().item()
    # fp16 tensor-core operands, fp32 accumulation and residual stream on O(1) LayerNorm outputs: measured 1.9e-3 max / 3e-4 mean
    # at 799 frames (profiles/r2a_parity_config2.json); the reference's own fp16 GPU path is 5.2e-3 / 5.8e-4 on the same input
    print(f"[parity] hubert {seconds}s layer {layer}: max {err:.3e} mean {(out - ref).abs().mean().item():.3e}")
    assert err < 5e-3, f"max abs err {err}"
    assert (out - ref).abs().mean().item() < 1e-3
    if layer == 9:
        with torch.no_grad():
            refp = OH.final_proj(w, ref[None])[0]
        outp = m.final_proj(out.cuda()).cpu()
        assert (outp - refp).abs().max().item() < 5e-3

"""CPU: the oracle against golden vectors produced by the REFERENCE's own modules (tests/golden/make_golden.py) and the
reference's own fixtures (infer/modules/vc/lgdsng.npz, logs/mute/2a_f0, 2b-f0nsf)."""
import os

import numpy as np
import torch

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_synth_oracle_matches_reference_golden():
    from oracle import synth as OS, weights as OW
    z = np.load(os.path.join(G, "synth_v2_48k_T24.npz"))
    w, cfg = OW.synth_weights(1234), OW.V2_48K_CONFIG
    phone = torch.from_numpy(z["phone"].astype(np.float32))
    pitch, pitchf = torch.from_numpy(z["pitch"]), torch.from_numpy(z["pitchf"])
    T = phone.shape[1]
    torch.manual_seed(int(z["seed"]))
    n1 = torch.randn(1, 192, T); torch.rand(1, 1, 1); n2 = torch.randn(1, T * 480, 1)
    with torch.no_grad():
        out = OS.synth_infer(w, cfg, phone, torch.tensor([T]), torch.tensor([int(z["sid"])]), pitch, pitchf, n1, n2)[0, 0]
    assert np.abs(out.numpy() - z["out"]).max() < 5e-6
    torch.manual_seed(int(z["seed_rt"]))
    n1 = torch.randn(1, 192, T); torch.rand(1, 1, 1); n2 = torch.randn(1, 6 * 480, 1)   # flow_head = max(16-24, 0) = 0
    with torch.no_grad():
        out = OS.synth_infer(w, cfg, phone, torch.tensor([T]), torch.tensor([int(z["sid"])]), pitch, pitchf, n1, n2, 16, 6, 6)[0, 0]
    assert np.abs(out.numpy() - z["out_rt"]).max() < 5e-6


def test_rmvpe_oracle_matches_reference_golden():
    from oracle import rmvpe as ORM, weights as OW
    z = np.load(os.path.join(G, "rmvpe_e2e_64.npz"))
    with torch.no_grad():
        hid = ORM.e2e_forward(OW.rmvpe_weights(4321), torch.from_numpy(z["mel"]))[0].numpy()
    assert np.abs(hid - z["hidden"]).max() < 1e-5
    assert np.abs(ORM.decode(z["hidden"].copy(), 0.03) - z["f0"]).max() == 0
    for x, r in zip(z["tracks"], z["resized"]):
        assert np.array_equal(ORM.interpolate_f0(ORM.resize_f0(x.copy(), 37)), r)


def test_post_process_against_reference_fixtures():
    """lgdsng.npz: the reference stores (pitch, pitchf) produced by its own post_process (hash.py:51-54);
    logs/mute: unvoiced input -> coarse 1, f0 0 (gen.py:34-40)."""
    from oracle import rmvpe as ORM
    from rvc_b200 import f0post
    z = np.load(os.path.join(G, "ref_fixtures_f0.npz"))
    for fn in (ORM.post_process, lambda f, k: f0post.post_process(f, k)):
        c, f = fn(z["lgdsng_pitchf"].copy(), 0)
        assert np.array_equal(c, z["lgdsng_pitch"])
        c, f = fn(z["mute_2b_f0nsf"].copy(), 0)
        assert np.array_equal(c, z["mute_2a_f0"]) and (f == 0).all()


def test_hubert_oracle_matches_transformers_golden():
    from oracle import hubert as OH, weights as OW
    z = np.load(os.path.join(G, "hubert_hf_0p5s.npz"))
    w = OW.hubert_weights(777)
    wav = OW.synth_voice(0.5, seed=6)[None]
    with torch.no_grad():
        assert np.abs(OH.extract_features(w, wav, 9)[0].numpy() - z["layer9"]).max() < 5e-5
        assert np.abs(OH.extract_features(w, wav, 12)[0].numpy() - z["layer12"]).max() < 5e-5

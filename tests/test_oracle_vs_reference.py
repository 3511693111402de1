"""CPU, build container only: the oracle against the reference's own modules imported from /root/reference
(skipped on the GPU box, where /root/reference does not exist)."""
import math
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/rvc"), reason="/root/reference not present")

# the product package mirrors the reference's module names (rvc.*, infer.*), so the comparison runs in a clean
# subprocess whose sys.path puts /root/reference first.
SCRIPT = r'''
import sys, math, torch, numpy as np
sys.path.insert(0, "/root/reference"); sys.path.insert(1, ROOT)
from rvc.synthesizer import get_synthesizer
from rvc.f0.e2e import E2E
assert "/root/reference" in get_synthesizer.__code__.co_filename
from oracle import weights as OW, synth as OS, rmvpe as ORM
torch.set_grad_enabled(False)
cpt = OW.synth_cpt(77, "v2")
net_g, _ = get_synthesizer({**cpt, "weight": dict(cpt["weight"]), "config": list(cpt["config"])}, "cpu")
w = OW.synth_weights(77)
assert set(net_g.state_dict()) == set(w)
T = 60
g = torch.Generator().manual_seed(1)
phone = torch.randn(1, T, 768, generator=g) * 0.5
pitchf = torch.zeros(1, T); pitchf[:, 5:50] = 150 + 50 * torch.rand(45, generator=g)
fm = 1127 * torch.log(1 + pitchf / 700); mn, mx = 1127 * math.log(1 + 50 / 700), 1127 * math.log(1 + 1100 / 700)
pitch = torch.round(torch.where(fm > 0, (fm - mn) * 254 / (mx - mn) + 1, fm).clamp(1, 255)).long()
torch.manual_seed(5); ref = net_g.infer(phone, torch.tensor([T]), torch.tensor([2]), pitch, pitchf)
torch.manual_seed(5); n1 = torch.randn(1, 192, T); torch.rand(1, 1, 1); n2 = torch.randn(1, T * 480, 1)
out = OS.synth_infer(w, cpt["config"], phone, torch.tensor([T]), torch.tensor([2]), pitch, pitchf, n1, n2)
assert (out - ref).abs().max().item() < 5e-6, (out - ref).abs().max().item()
torch.manual_seed(6); ref = net_g.infer(phone, torch.tensor([T]), torch.tensor([2]), pitch, pitchf, skip_head=40, return_length=12, return_length2=14)
torch.manual_seed(6); n1 = torch.randn(1, 192, T - 16); torch.rand(1, 1, 1); n2 = torch.randn(1, 12 * 480, 1)
out = OS.synth_infer(w, cpt["config"], phone, torch.tensor([T]), torch.tensor([2]), pitch, pitchf, n1, n2, 40, 12, 14)
assert out.shape == ref.shape and (out - ref).abs().max().item() < 5e-6
# no-f0 family (SynthesizerTrnMs768NSFsid_nono: no emb_pitch, plain Generator decoder)
cpt0 = OW.synth_cpt(5, "v2", f0=0)
net0, _ = get_synthesizer({**cpt0, "weight": dict(cpt0["weight"]), "config": list(cpt0["config"])}, "cpu")
w0 = OW.synth_weights(5, use_f0=False)
assert set(net0.state_dict()) == set(w0)
torch.manual_seed(3); ref = net0.infer(phone, torch.tensor([T]), torch.tensor([1]))
torch.manual_seed(3); n1 = torch.randn(1, 192, T)
out = OS.synth_infer(w0, cpt0["config"], phone, torch.tensor([T]), torch.tensor([1]), None, None, n1, None)
assert out.shape == ref.shape and (out - ref).abs().max().item() < 5e-6
# the other decoder schedules of configs/{v1,v2}/*.json (5-stage decoders, kernel 16 at stride 4 / 6)
for cfgx, ver in ((OW.V1_48K_CONFIG, "v1"), (OW.V1_32K_CONFIG, "v1"), (OW.V2_32K_CONFIG, "v2")):
    cptx = OW.synth_cpt(5, ver, config=cfgx)
    netx, _ = get_synthesizer({**cptx, "weight": dict(cptx["weight"]), "config": list(cptx["config"])}, "cpu")
    encx = 256 if ver == "v1" else 768
    wx = OW.synth_weights(5, cfgx, encx)
    assert set(netx.state_dict()) == set(wx)
    Tx = 24; phx = torch.randn(1, Tx, encx, generator=g) * 0.5
    torch.manual_seed(3); ref = netx.infer(phx, torch.tensor([Tx]), torch.tensor([1]), pitch[:, :Tx], pitchf[:, :Tx])
    torch.manual_seed(3); n1 = torch.randn(1, 192, Tx); torch.rand(1, 1, 1); n2 = torch.randn(1, Tx * (cfgx[-1] // 100), 1)
    out = OS.synth_infer(wx, cptx["config"], phx, torch.tensor([Tx]), torch.tensor([1]), pitch[:, :Tx], pitchf[:, :Tx], n1, n2)
    assert out.shape == ref.shape and (out - ref).abs().max().item() < 5e-6, (cfgx[-1], (out - ref).abs().max().item())
m = E2E(4, 1, (2, 2)).eval(); rw = OW.rmvpe_weights(9)
assert set(m.state_dict()) == set(rw)
m.load_state_dict(rw)
mel = torch.randn(1, 128, 96, generator=g) * 3 - 4
assert (m(mel) - ORM.e2e_forward(rw, mel)).abs().max().item() < 1e-5
print("OK")
'''


def test_oracle_matches_reference_modules():
    r = subprocess.run([sys.executable, "-c", SCRIPT.replace("ROOT", repr(ROOT))], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]

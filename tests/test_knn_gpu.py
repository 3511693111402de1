"""GPU parity: IVF-Flat search (bit-exact distances and indices), blend, brute-force top-1."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(n=4000, d=768, nlist=None, seed=0):
    from oracle import ivf as OI, weights as OW
    vec = OW.index_vectors(n, d, seed).numpy()
    return OI.build_ivf(vec, nlist, seed=seed, exact_assign=True), vec


def test_ivf_search_bit_exact_and_blend():
    from oracle import ivf as OI
    from rvc_b200.engine import Index
    idx, vec = _mk()
    rng = np.random.RandomState(1)
    q = (vec[rng.choice(len(vec), 150)] + 0.05 * rng.randn(150, 768)).astype(np.float32)
    q[:5] = vec[:5]                       # exact hits: distance 0 -> inf weights, NaN blend (IEEE, like numpy)
    Dr, Ir = idx.search(q, 8)
    g = Index.from_oracle_layout(idx)
    Dg, Ig = g.search_device(torch.from_numpy(q).cuda(), 8)
    assert np.array_equal(Ig.cpu().numpy(), Ir), "retrieval indices must be bit-exact"
    assert np.array_equal(Dg.cpu().numpy().view(np.uint32), Dr.view(np.uint32)), "distances must be bit-exact"
    feats = rng.randn(150, 768).astype(np.float32)
    ref = OI.blend(feats, Dr, Ir, idx.vectors, 0.75)
    out = g.blend_device(torch.from_numpy(feats).cuda(), Dg, Ig, 0.75).cpu().numpy()
    ok = ~np.isnan(ref)
    assert np.array_equal(np.isnan(out), np.isnan(ref))
    assert np.abs(out[ok] - ref[ok]).max() <= 1e-6


def test_ivf_short_lists_pad_with_minus_one():
    from rvc_b200.engine import Index
    idx, vec = _mk(n=300, nlist=100, seed=2)      # ~3 vectors per list < k
    q = vec[:40] + 0.01
    Dr, Ir = idx.search(q, 8)
    assert (Ir < 0).any()
    g = Index.from_oracle_layout(idx)
    Dg, Ig = g.search_device(torch.from_numpy(q.astype(np.float32)).cuda(), 8)
    assert np.array_equal(Ig.cpu().numpy(), Ir)
    assert np.array_equal(Dg.cpu().numpy().view(np.uint32), Dr.view(np.uint32))


@pytest.mark.parametrize("n,nq,d", [(3000, 1, 768), (5000, 70, 768), (2048, 33, 256)])
def test_bruteforce_top1(n, nq, d):
    from oracle import ivf as OI, weights as OW
    from rvc_b200.engine import knn_bruteforce_top1
    db = OW.index_vectors(n, d, 5).numpy()
    rng = np.random.RandomState(7)
    q = (db[rng.choice(n, nq)] + 0.1 * rng.randn(nq, d)).astype(np.float32)
    Dr, Ir = OI.brute_force_top1(q, db)
    Dg, Ig = knn_bruteforce_top1(torch.from_numpy(db).cuda(), torch.from_numpy(q).cuda())
    assert np.array_equal(Ig.cpu().numpy(), Ir)
    assert np.array_equal(Dg.cpu().numpy().view(np.uint32), Dr.view(np.uint32))


def test_bruteforce_large_property():
    """BASELINE config #5 size class (1e6 x 768): the nearest neighbour of a stored row is itself at distance 0."""
    from rvc_b200.engine import knn_bruteforce_top1
    g = torch.Generator(device="cuda").manual_seed(0)
    db = torch.randn(1_000_000, 768, device="cuda", generator=g)
    rows = torch.tensor([0, 17, 999_999, 123_456, 500_000], device="cuda")
    D, I = knn_bruteforce_top1(db, db[rows].clone())
    assert torch.equal(I, rows) and (D == 0).all()


@pytest.mark.parametrize("N,nq", [(5000, 40), (70000, 700), (33000, 2100)])
def test_flat_tensor_core_short_list_is_bit_identical_to_the_exact_scan(N, nq):
    """rvcb_flat_search_top1 (fp16 GEMM scores -> 32 candidates -> exact lane-order re-rank + rounding certificate) vs the
    exact SIMT scan: D and I bit for bit, including exact duplicates in the database (ties -> lowest row) and near-ties."""
    import torch
    from rvc_b200 import engine
    g = torch.Generator(device="cuda").manual_seed(N + nq)
    db = torch.randn(N, 768, device="cuda", generator=g) * 0.3
    db[N // 2: N // 2 + 50] = db[7]                                   # 51 identical rows
    db[100:140] = db[99] + 1e-4 * torch.randn(40, 768, device="cuda", generator=g)      # near-duplicates inside fp16 resolution
    q = db[torch.randint(0, N, (nq,), device="cuda", generator=g)] + 0.03 * torch.randn(nq, 768, device="cuda", generator=g)
    q[0] = db[7]
    q[1] = db[99]
    q[2] = db[120] + 1e-5
    D0, I0 = engine.knn_bruteforce_top1(db, q)
    flat = engine.FlatIndex(db)
    D1, I1 = flat.search(q)
    assert torch.equal(I0, I1) and torch.equal(D0, D1)
    assert int(I1[0]) == 7                                            # lowest of the 51 identical rows
    # small batches take the scan itself
    D2, I2 = flat.search(q[:5])
    assert torch.equal(I2, I0[:5]) and torch.equal(D2, D0[:5])

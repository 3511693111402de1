"""CPU: host-side logic, the C-ABI library surface, the numpy IVF oracle's own invariants, and the N>1 plumbing (gloo)."""
import os
import re
import socket
import sys
import tempfile

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    from rvc_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "rvcb200.h")).read()
    declared = set(re.findall(r"\b(rvcb_[a-z0-9_]+)\s*\(", hdr))
    l = _lib.lib()
    for name in sorted(declared):
        assert hasattr(l, name), f"librvcb200.so does not export {name}"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert b"sm_100a" in l.rvcb_version()


def test_product_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from rvc_b200 import synthetic as SY
    from rvc.synthesizer import get_synthesizer
    with pytest.raises(RuntimeError):
        get_synthesizer(SY.synth_cpt(1, "v2"), "cuda:0")      # no CPU / PyTorch fallback on the product path


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "retrieval-based-voice-conversion-webui_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), os.path.join(dp, f)


def test_f0post_matches_oracle_loops():
    from oracle import rmvpe as ORM
    from rvc_b200 import f0post
    rng = np.random.RandomState(0)
    for _ in range(400):
        n = rng.randint(1, 60)
        x = np.abs(rng.randn(n)) * 100 * (rng.rand(n) > rng.rand())
        assert np.array_equal(ORM.interpolate_f0(x.copy()), f0post.interpolate_f0(x.copy()))
        tl = rng.randint(1, 80)
        assert np.array_equal(ORM.resize_f0(x.copy(), tl), f0post.resize_f0(x.copy(), tl))
        a, b = ORM.post_process(x.copy(), 3), f0post.post_process(x.copy(), 3)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_ivf_oracle_invariants_and_index_file_roundtrip():
    from oracle import ivf as OI, weights as OW
    from rvc_b200 import faiss_io
    vec = OW.index_vectors(700, 768, 3).numpy()
    idx = OI.build_ivf(vec, 14, seed=1, exact_assign=True)
    D, I = idx.search(vec[:50], 8)
    assert np.array_equal(I[:, 0], np.arange(50)) and (D[:, 0] == 0).all()        # a stored vector finds itself
    assert (np.diff(D, axis=1) >= 0).all()                                          # ascending
    bf_D, bf_I = OI.brute_force_top1(vec[100:120] + 0.01, vec)
    assert np.array_equal(bf_I, np.arange(100, 120))
    ref = ((vec[100:120] + np.float32(0.01) - vec[100:120]).astype(np.float64) ** 2).sum(1)
    assert np.allclose(bf_D, ref, rtol=1e-5)
    # empty / short lists pad with (FLT_MAX, -1) like faiss
    tiny = OI.build_ivf(vec[:60], 30, seed=0, exact_assign=True)
    D, I = tiny.search(vec[:10] + 0.5, 8)
    assert (I == -1).any() and (D[I == -1] == OI.FLT_MAX).all()
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "added_IVF14_Flat_nprobe_1_x_v2.index")
        faiss_io.write_index(p, idx)
        lay = faiss_io.read_index(p)
        for f in ("centroids", "vectors", "list_off", "list_ids"):
            assert np.array_equal(getattr(lay, f), getattr(idx, f))
        faiss_io.save_layout(os.path.join(td, "x.npz"), idx)
        assert np.array_equal(faiss_io.read_index(os.path.join(td, "x.npz")).vectors, idx.vectors)
    # blend: numpy semantics incl. -1 ids and exact hits
    feats = np.random.RandomState(0).randn(10, 768).astype(np.float32)
    out = OI.blend(feats, D, I, tiny.vectors, 0.75)
    assert out.shape == feats.shape and np.isfinite(out).all()


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.join(ROOT, "retrieval-based-voice-conversion-webui_b200"))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from rvc_b200 import dist_utils, synthetic as SY
    from rvc_b200.index_build import build_ivf_layout
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lay = build_ivf_layout(SY.index_vectors(400, 768, 0).numpy(), 8, device="cpu") if rank == 0 else None
    got = dist_utils.broadcast_layout(lay, 0, "cpu")
    mine = dist_utils.shard(list(range(11)), rank, world)
    t = torch.tensor([float(len(mine))])
    dist.all_reduce(t)
    q.put((rank, float(got.vectors.sum()), int(got.list_off[-1]), mine, float(t.item())))
    dist.destroy_process_group()


def test_world_size_2_sharding_and_index_broadcast_gloo():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in ps]
    (r0, s0, n0, m0, t0), (r1, s1, n1, m1, t1) = res
    assert s0 == s1 and n0 == n1 == 400                     # identical index on both ranks
    assert sorted(m0 + m1) == list(range(11)) and not set(m0) & set(m1)   # every utterance exactly once
    assert t0 == t1 == 11.0


def test_host_filtfilt_is_scipy_bit_for_bit():
    """rvcb_host_filtfilt (C, host) == scipy.signal.filtfilt(bh, ah, x) of pipeline.py:23,221, including the error on inputs
    no longer than padlen."""
    from scipy import signal
    from rvc_b200 import engine
    bh, ah = signal.butter(N=5, Wn=48, btype="high", fs=16000)
    zi = signal.lfilter_zi(bh, ah)
    rng = np.random.default_rng(3)
    for n in (19, 20, 1000, 160000):
        x = (rng.standard_normal(n) * 0.3).astype(np.float32)
        assert np.array_equal(signal.filtfilt(bh, ah, x), engine.host_filtfilt(bh, ah, zi, x)), n      # float32 audio (the reference's)
        x64 = x.astype(np.float64) * 1.0000001
        assert np.array_equal(signal.filtfilt(bh, ah, x64), engine.host_filtfilt(bh, ah, zi, x64)), n
    with pytest.raises(RuntimeError, match="padlen"):
        engine.host_filtfilt(bh, ah, zi, np.zeros(18))


def test_highpass_sections_reproduce_the_direct_form_filter():
    """The device filter's second-order sections (engine.highpass_sos_from_ba) realise the same transfer function as the
    reference's (bh, ah): denominator reconstructs to 1e-11, and scipy's own sosfiltfilt over them agrees with
    filtfilt(bh, ah) to 1e-7 on broadband and sub-corner signals."""
    from scipy import signal
    from rvc_b200 import engine
    bh, ah = signal.butter(N=5, Wn=48, btype="high", fs=16000)
    sos, zi = engine.highpass_sos_from_ba(bh, ah)
    assert sos.shape == (3, 6) and zi.shape == (3, 2) and np.all(sos[:, 3] == 1.0)
    a_rec = np.polymul(np.polymul(sos[0, 3:], sos[1, 3:]), sos[2, 3:5])
    assert np.abs(a_rec - ah).max() < 1e-11
    rng = np.random.default_rng(2)
    t = np.arange(48000) / 16000
    for x in (rng.standard_normal(48000) * 0.3 + 0.05, 0.5 * np.sin(2 * np.pi * 110 * t) + 0.3 * np.sin(2 * np.pi * 30 * t) + 0.1):
        y = signal.sosfiltfilt(sos, x, padtype="odd", padlen=18)
        assert np.abs(y - signal.filtfilt(bh, ah, x)).max() < 1e-7


def test_vc_facade_host_logic_with_stub_models(monkeypatch, tmp_path):
    """The VC facade (infer/modules/vc/modules.py:18-266) without a GPU: Gradio-shaped returns of get_vc, index-path
    resolution, peak normalisation, the info strings and the exception -> info-string convention of vc_single, and the
    folder loop of vc_multi (rank striding, per-file log).  Models and pipeline are stubs; nothing here computes audio."""
    import types
    from infer.modules.vc import modules as M

    seen = {}

    class FakeNet:
        def half(self): return self
        def float(self): return self

    class FakePipeline:
        def __init__(self, tgt_sr, config):
            self.tgt_sr = tgt_sr

        def pipeline(self, model, net_g, sid, audio, times, f0_up_key, f0_method, file_index, index_rate, if_f0, *rest):
            seen.update(audio_peak=float(np.abs(audio).max()), file_index=file_index, want_i16=getattr(self, "_want_int16", None),
                        f0_up_key=f0_up_key, if_f0=if_f0)
            if f0_method == "harvest":
                raise ValueError("f0 method harvest has not yet been supported")
            times[1] += 0.25
            return np.linspace(-2000.7, 2000.7, 480).astype(np.float32)

    def fake_get(cpt, device):
        return FakeNet(), {**cpt, "config": list(cpt["config"])}

    monkeypatch.setattr(M, "Pipeline", FakePipeline)
    monkeypatch.setattr(M, "get_synthesizer", fake_get)
    monkeypatch.setattr(M, "load_hubert", lambda device, is_half: "hubert")
    monkeypatch.setattr(M, "get_index_path_from_model", lambda name: f"logs/{name}/added.index")
    monkeypatch.setattr(M.torch.cuda, "empty_cache", lambda: None)
    cfg = types.SimpleNamespace(device="cuda:0", is_half=True)
    cpt = {"config": [1025, 32, 192, 192, 768, 2, 6, 3, 0, "1", [3, 7, 11], [[1, 3, 5]] * 3, [12, 10, 2, 2], 512, [24, 20, 4, 4], 109, 256, 48000],
           "weight": {"emb_g.weight": torch.zeros(7, 256)}, "f0": 1, "version": "v2", "info": "200 epochs", "name": "alice.pth"}
    vc = M.VC(cfg)
    # no model selected yet: both return shapes of get_vc("")
    assert vc.get_vc("") == {"visible": True, "maximum": 0, "__type__": "update"}
    out = vc.get_vc("", 0.4, 0.2, "a.index", "b.index")
    assert out[0] == {"visible": False, "__type__": "update"} and out[1]["value"] == 0.4 and out[2]["value"] == 0.2
    assert out[3]["value"] == "a.index" and out[4]["value"] == "b.index" and out[5]["value"] == ""
    # select a model: speaker count comes from the embedding table, index path from the model name
    assert vc.get_vc(cpt) == {"visible": True, "maximum": 7, "__type__": "update"}
    assert (vc.tgt_sr, vc.if_f0, vc.version, vc.n_spk) == (48000, 1, "v2", 7)
    full = vc.get_vc(cpt, 0.4, 0.2)
    assert full[0]["maximum"] == 7 and full[1] == {"visible": True, "value": 0.4, "__type__": "update"}
    assert full[3] == full[4] == {"value": "logs/alice.pth/added.index", "__type__": "update"} and full[5]["value"] == "200 epochs"
    # one utterance: loud input is scaled to a 0.95 peak, the text-box index wins over the dropdown and is cleaned up
    audio = np.full(32000, 3.0, dtype=np.float32)
    info, (sr, wav) = vc.vc_single(0, audio, "2", None, "rmvpe", ' "logs/x/trained_IVF1_Flat.index"\n', "logs/y/added.index", 0.75, 3, 0, 0.25, 0.33)
    assert info == "Success.\nIndex not used.\nTime: npy: 0.00s, f0: 0.25s, infer: 0.00s." and sr == 48000
    assert wav.dtype == np.int16 and wav[0] == -2000 and wav[-1] == 2000          # C truncation, like .astype(np.int16)
    assert abs(seen["audio_peak"] - 0.95) < 1e-6 and seen["file_index"] == "logs/x/added_IVF1_Flat.index"
    assert seen["want_i16"] is True and vc.pipeline._want_int16 is False and seen["f0_up_key"] == 2 and vc.hubert_model == "hubert"
    ipath = tmp_path / "added.index"
    ipath.write_bytes(b"x")
    info, _ = vc.vc_single(0, audio.copy(), 0, None, "rmvpe", "", str(ipath), 0.75, 3, 0, 0.25, 0.33)
    assert f"Index: {ipath}." in info and seen["file_index"] == str(ipath)
    info, _ = vc.vc_single(0, audio.copy(), 0, None, "rmvpe", None, "", 0.75, 3, 44100, 0.25, 0.33)
    assert seen["file_index"] == "" and _[0] == 44100                                # resample_sr >= 16000 and != tgt_sr is reported
    # errors become the info string; a missing input is its own message
    assert vc.vc_single(0, audio.copy(), 0, None, "harvest", "", "", 0.75, 3, 0, 0.25, 0.33) == ("f0 method harvest has not yet been supported", None)
    assert vc.vc_single(0, None, 0, None, "rmvpe", "", "", 0.75, 3, 0, 0.25, 0.33) == ("You need to upload an audio", None)
    # a folder: every file of this rank, cumulative log, audio written next to it
    indir, outdir = tmp_path / "in", tmp_path / "out"
    indir.mkdir()
    for n in ("a.wav", "b.wav", "c.wav"):
        (indir / n).write_bytes(b"")
    saved = []
    monkeypatch.setattr(M, "load_audio", lambda path, sr: np.zeros(16000, np.float32) + 0.1)
    monkeypatch.setattr(M, "save_audio", lambda path, wav, sr, f32=False: saved.append((os.path.basename(path), sr, wav.dtype)))
    monkeypatch.setenv("RANK", "1")
    monkeypatch.setenv("WORLD_SIZE", "2")
    msgs = list(vc.vc_multi(0, f' "{indir}" ', str(outdir), [], 0, "rmvpe", "", "", 0.75, 3, 0, 0.25, 0.33, "flac"))
    assert outdir.is_dir() and len(saved) == len(os.listdir(indir)[1::2]) and all(s[0].endswith(".flac") and s[1] == 48000 for s in saved)
    assert msgs[-1].count("->Success.") == len(saved) and msgs[-1] == msgs[-2]


def test_training_side_extraction_scripts_host_logic(tmp_path):
    """Drop-ins for infer/modules/train/extract_{feature,f0}_print.py: file discovery, [i::n] striding,
    skip-if-present, output names / shapes / dtypes and the log lines, with stub models (the arithmetic behind them is the
    HuBERT / RMVPE path of the inference tests)."""
    from infer.modules.train import extract_f0_print as XF0, extract_feature_print as XFE
    exp = tmp_path / "exp"
    wavs = exp / "1_16k_wavs"
    wavs.mkdir(parents=True)
    for n in ("0_0.wav", "0_1.wav", "0_2.wav", "0_3.wav", "notes.txt", "0_4.spec.wav"):
        (wavs / n).write_bytes(b"")

    class FakeHubert:
        def __init__(self): self.calls = []
        def extract_features(self, source, padding_mask, output_layer):
            assert source.shape[0] == 1 and padding_mask.shape == source.shape and not padding_mask.any()
            self.calls.append(output_layer)
            T = source.shape[1] // 320
            x = torch.full((1, T, 768), float(len(self.calls)))
            if len(self.calls) == 2:
                x[0, 0, 0] = float("nan")
            return x, None
        def final_proj(self, x): return x[..., :256]

    lines = []
    load = lambda p: np.zeros(3200, np.float32)
    m = FakeHubert()
    n = XFE.run(m, str(exp), "v2", 2, 0, load, lines.append)            # rank 0 of 2: files 0, 2, 4 of the sorted listing
    mine = sorted(os.listdir(wavs))[0::2]
    assert m.calls == [12] * sum(f.endswith(".wav") for f in mine)
    assert lines[0] == "all-feature-%d" % len(mine) and lines[-1] == "all-feature-done" and any("contains nan" in l for l in lines)
    out = sorted(os.listdir(exp / "3_feature768"))
    assert n == len(out) and all(f.endswith(".npy") for f in out) and "0_0.npy" in out
    assert np.load(exp / "3_feature768" / "0_0.npy").shape == (10, 768)
    lines.clear()
    assert XFE.run(FakeHubert(), str(exp), "v2", 2, 0, load, lines.append) <= 1     # existing outputs are skipped (only the NaN file is retried)
    m1 = FakeHubert()
    XFE.run(m1, str(exp), "v1", 1, 0, load, lines.append)
    assert m1.calls and set(m1.calls) == {9} and np.load(exp / "3_feature256" / "0_0.npy").shape == (10, 256)
    assert XFE.run(FakeHubert(), str(exp), "v2", 64, 63, load, lines.append) == 0 and lines[-1] == "no-feature-todo"

    class FakeF0:
        def calculate(self, x, p_len, key, method, radius, manual=None):
            assert (p_len, key, method) == (x.shape[0] // 160, 0, "rmvpe")
            return np.full(p_len, 42, np.int32), np.full(p_len, 220.0)

    jobs = XF0.list_jobs(str(exp))
    assert [os.path.basename(j[0]) for j in jobs] == ["0_0.wav", "0_1.wav", "0_2.wav", "0_3.wav", "notes.txt"]     # "spec" files are skipped
    lines.clear()
    assert XF0.run(FakeF0(), jobs[:4], "rmvpe", load, lines.append) == 4
    assert lines[0] == "todo-f0-4" and lines[1].startswith("f0ing,now-0,all-4,-")
    c, f = np.load(exp / "2a_f0" / "0_1.wav.npy"), np.load(exp / "2b-f0nsf" / "0_1.wav.npy")
    assert c.dtype == np.int32 and c.shape == f.shape == (20,) and c[0] == 42 and f[0] == 220.0
    assert XF0.run(FakeF0(), jobs[:4], "rmvpe", load, lines.append) == 0          # both outputs present -> skipped
    assert XF0.run(FakeF0(), [], "rmvpe", load, lines.append) == 0 and lines[-1] == "no-f0-todo"
    assert XFE.main(["x"]) == 0                                                    # wrong arity: silently exit 0 like the reference


def test_faiss_container_reader_variants(tmp_path):
    """The IwFl reader on the variants faiss itself writes besides our own writer's: sparse list-size table ("sprs", used when
    most lists are empty), an array direct map, and the refusals (other index types, inner-product metric)."""
    import struct
    from rvc_b200 import faiss_io
    d, nlist = 8, 6
    rng = np.random.default_rng(0)
    cent = rng.standard_normal((nlist, d)).astype("<f4")
    vec = rng.standard_normal((5, d)).astype("<f4")
    lists = {1: [3, 0], 4: [2, 4, 1]}                                  # list -> ids (4 of 6 lists empty)

    def hdr(dd, nt, metric=1): return struct.pack("<iqqqBi", dd, nt, 1 << 20, 1 << 20, 1, metric)

    def build(fourcc=b"IwFl", metric=1):
        b = fourcc + hdr(d, 5, metric) + struct.pack("<QQ", nlist, 1)
        b += b"IxF2" + hdr(d, nlist) + struct.pack("<Q", cent.size) + cent.tobytes()
        b += struct.pack("<BQ", 1, 5) + np.arange(5, dtype="<i8").tobytes()           # direct map type 1 (array) with 5 entries
        b += b"ilar" + struct.pack("<QQ", nlist, 4 * d) + b"sprs" + struct.pack("<Q", 2 * len(lists))
        b += b"".join(struct.pack("<QQ", l, len(ids)) for l, ids in lists.items())
        for l, ids in lists.items():
            b += vec[ids].tobytes() + np.asarray(ids, dtype="<i8").tobytes()
        return b
    p = tmp_path / "added_IVF6_Flat_nprobe_1_x_v2.index"
    p.write_bytes(build())
    lay = faiss_io.read_index(str(p))
    assert np.array_equal(lay.centroids, cent) and np.array_equal(lay.vectors, vec)
    assert lay.list_off.tolist() == [0, 0, 2, 2, 2, 5, 5] and lay.list_ids.tolist() == [3, 0, 2, 4, 1]
    p.write_bytes(build(fourcc=b"IxF2"))
    with pytest.raises(ValueError, match="unsupported faiss index type"):
        faiss_io.read_index(str(p))
    p.write_bytes(build(metric=0))
    with pytest.raises(ValueError, match="METRIC_L2"):
        faiss_io.read_index(str(p))


def test_train_index_file_protocol_and_minibatch_kmeans_branch(tmp_path):
    """rvc_b200.index_build.train_index vs web.py:499-596: feature files -> shuffled total_fea.npy, n_ivf formula, trained_ / added_
    index files in the IwFl container (read back by the faiss-free reader), the MiniBatchKMeans branch above the row threshold."""
    from rvc_b200 import faiss_io, index_build
    logs = tmp_path / "logs"
    fdir = logs / "exp1" / "3_feature768"
    fdir.mkdir(parents=True)
    rng = np.random.default_rng(0)
    parts = [rng.standard_normal((n, 768)).astype(np.float32) * 0.2 for n in (700, 900, 400)]
    for i, p in enumerate(parts):
        np.save(fdir / f"u{i}.npy", p)
    out_root = tmp_path / "outside"
    out_root.mkdir()
    np.random.seed(3)
    msgs = list(index_build.train_index("exp1", "v2", logs_root=str(logs), outside_index_root=str(out_root), device="cpu"))
    n, n_ivf = 2000, min(int(16 * np.sqrt(2000)), 2000 // 39)
    assert f"({n}, 768),{n_ivf}" in msgs[-1] and "training" in msgs[-1] and "adding" in msgs[-1] and "Successfully built index into" in msgs[-1]
    exp = logs / "exp1"
    total = np.load(exp / "total_fea.npy")
    np.random.seed(3)
    idx = np.arange(n); np.random.shuffle(idx)
    assert np.array_equal(total, np.concatenate(parts, 0)[idx])                     # web.py:517-520, 543
    added = exp / f"added_IVF{n_ivf}_Flat_nprobe_1_exp1_v2.index"
    trained = exp / f"trained_IVF{n_ivf}_Flat_nprobe_1_exp1_v2.index"
    assert added.exists() and trained.exists()
    lay = faiss_io.read_index(str(added))
    assert lay.centroids.shape == (n_ivf, 768) and np.array_equal(lay.vectors, total)      # reconstruct_n(0, ntotal) == total_fea (add order)
    assert sorted(lay.list_ids.tolist()) == list(range(n)) and lay.list_off[-1] == n
    # every vector sits in the list of its nearest centroid
    d = ((total[:, None, :] - lay.centroids[None, :, :]) ** 2).sum(-1)
    assign = np.empty(n, np.int64)
    for l in range(n_ivf):
        assign[lay.list_ids[lay.list_off[l]:lay.list_off[l + 1]]] = l
    assert (d[np.arange(n), assign] <= d.min(1) * (1 + 1e-5) + 1e-6).all()
    empty = faiss_io.read_index(str(trained))
    assert empty.ntotal == 0 and np.array_equal(empty.centroids, lay.centroids)
    assert (out_root / f"exp1_IVF{n_ivf}_Flat_nprobe_1_exp1_v2.index").exists()
    # the > 2e5-rows branch (threshold lowered): features are replaced by MiniBatchKMeans centres
    msgs = list(index_build.train_index("exp1", "v2", logs_root=str(logs), device="cpu", kmeans_threshold=1000, kmeans_clusters=300))
    assert "Trying doing kmeans 2000 shape to 10k centers." in msgs[0]
    assert np.load(exp / "total_fea.npy").shape == (300, 768)
    assert list(index_build.train_index("missing", "v2", logs_root=str(logs)))[0].startswith("请先进行特征提取")


def test_index_reader_rejects_inconsistent_files(tmp_path):
    from oracle import ivf as OI, weights as OW
    from rvc_b200 import faiss_io
    idx = OI.build_ivf(OW.index_vectors(300, 768, 1).numpy(), 8, seed=0, exact_assign=True)
    path = str(tmp_path / "a.index")
    faiss_io.write_index(path, idx)
    good = open(path, "rb").read()
    assert faiss_io.read_index(path).ntotal == 300
    import struct
    bad = bytearray(good)
    struct.pack_into("<q", bad, 4 + 4, 301)                          # ntotal in the header no longer matches the list sizes
    open(path, "wb").write(bytes(bad))
    with pytest.raises(ValueError):
        faiss_io.read_index(path)
    bad = bytearray(good)
    struct.pack_into("<Q", bad, 4 + 33 + 8, 4)                       # nprobe != 1
    open(path, "wb").write(bytes(bad))
    with pytest.raises(ValueError):
        faiss_io.read_index(path)
    bad = bytearray(good)
    struct.pack_into("<q", bad, len(bad) - 8, 10 ** 6)               # last vector id out of range
    open(path, "wb").write(bytes(bad))
    with pytest.raises(ValueError):
        faiss_io.read_index(path)


def test_checkpoint_unpickler_is_an_allowlist(tmp_path):
    """infer/modules/vc/utils.py: a crafted hubert_base.pt must not be able to import callables (os.system, eval, ...)."""
    import argparse
    import collections
    from infer.modules.vc import utils as U

    class Evil:
        def __reduce__(self):
            return (os.system, ("touch %s" % (tmp_path / "pwned"),))
    sd = collections.OrderedDict(a=torch.randn(3, 4).half(), b=torch.nn.Parameter(torch.randn(2)))
    torch.save({"model": sd, "cfg": argparse.Namespace(x=1), "evil": Evil(), "np": np.arange(3)}, tmp_path / "ck.pt")
    out = U._load_fairseq_state_dict(str(tmp_path / "ck.pt"))
    assert set(out) == {"a", "b"} and out["a"].dtype == torch.float16 and not (tmp_path / "pwned").exists()


def test_mel_filterbank_matches_torchaudio_htk_slaney():
    """oracle/rmvpe.py mel_filterbank restates librosa.filters.mel(sr=16000, n_fft=1024, n_mels=128, fmin=30, fmax=8000, htk=True)
    (rvc/f0/mel.py:27-34; librosa is not installed): independent check against torchaudio's HTK / slaney-normalised bank."""
    torchaudio = pytest.importorskip("torchaudio")
    from oracle import rmvpe as ORM
    fb = torchaudio.functional.melscale_fbanks(513, 30.0, 8000.0, 128, 16000, norm="slaney", mel_scale="htk").t().numpy()
    assert np.abs(ORM.mel_filterbank() - fb).max() < 1e-6


def test_realtime_host_side_pieces(tmp_path):
    """CPU-testable parts of the realtime / batch front doors: the response-threshold gate (gui.py:951-966) against the oracle
    callback's restatement over consecutive blocks, the TorchGate mask-smoothing filter (torchgate.py:72-126), save_audio's two modes
    (infer/lib/audio.py:29-56)."""
    from scipy.io import wavfile
    from infer.lib.audio import save_audio
    from infer.modules.gui.realtime_block import response_gate
    from infer.modules.gui.torchgate import TorchGate
    from oracle import rtrvc as ORT, torchgate as OT
    zc = 480
    rng = np.random.default_rng(5)

    class _Stub:                      # OracleCallback.block's gate needs only these fields
        tgt_sr = 48000
    orc = ORT.OracleCallback.__new__(ORT.OracleCallback)
    orc.zc, orc.threhold, orc.rms_buffer = zc, -40.0, np.zeros(4 * zc, dtype="float32")
    buf = np.zeros(4 * zc, dtype="float32")
    for b in range(4):
        x = (rng.standard_normal(7680) * (0.2 if b % 2 else 0.002)).astype(np.float32)
        x[: 7680 // 3] *= 0.01
        got = response_gate(x.copy(), buf, zc, -40.0)
        # the oracle's literal restatement of the same lines
        ind = np.append(orc.rms_buffer, x)
        rms = ORT.rms_frames(ind, 4 * zc, zc)[:, 2:]
        orc.rms_buffer[:] = ind[-4 * zc:]
        ind = ind[2 * zc - zc // 2:]
        db = 20.0 * np.log10(np.maximum(1e-5, rms)); db = np.maximum(db, db.max() - 80.0)
        for i in np.nonzero(db[0] < -40.0)[0]:
            ind[i * zc: (i + 1) * zc] = 0
        ref = ind[zc // 2:]
        assert got.shape == ref.shape == (7680 + 2 * zc,) and np.array_equal(got, ref) and np.array_equal(buf, orc.rms_buffer)
        assert (got == 0).any() and (b % 2 == 0 or (got != 0).any())      # the quiet third of a loud block is gated, the rest kept
    for sr, n_fft in ((48000, 1920), (40000, 1600), (16000, 1024)):
        f = TorchGate(sr=sr, n_fft=n_fft).smoothing_filter[0, 0]
        assert torch.equal(f, OT.smoothing_filter(sr, n_fft, n_fft // 4)) and abs(float(f.sum()) - 1.0) < 1e-6
    i16 = (rng.standard_normal(1000) * 8000).astype(np.int16)
    save_audio(str(tmp_path / "a.wav"), i16, 48000, f32=True)
    sr, y = wavfile.read(str(tmp_path / "a.wav"))
    assert sr == 48000 and y.dtype == np.float32 and np.array_equal(y, i16.astype(np.float32))        # int16-range floats, like the reference
    fl = np.linspace(-0.5, 0.5, 100).astype(np.float32)
    save_audio(str(tmp_path / "b.wav"), fl, 16000)
    sr, y = wavfile.read(str(tmp_path / "b.wav"))
    assert y.dtype == np.int16 and np.array_equal(y, np.multiply(fl, 32767).astype(np.int16))

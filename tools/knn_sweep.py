"""BASELINE config #5: brute-force L2 top-1 sweep, d = 768, fp32 database; reports GB/s = N*768*4 / t per query tile pass
and the HBM roofline fraction (MEASURED_PEAKS.json).  Also times the RVC-realistic IVF2564 k=8 search."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "retrieval-based-voice-conversion-webui_b200"))
from rvc_b200 import _lib, engine, synthetic as SY  # noqa: E402
from rvc_b200.index_build import build_ivf_layout  # noqa: E402

_lib.init(0)
try:
    _pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    peak, tpeak, src = _pk["hbm_gbs"], _pk["bf16_tflops"], "measured"
except Exception:
    peak, tpeak, src = 6650.0, 1590.0, "fallback"


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


print(f"# brute-force L2 top-1, d=768 fp32; HBM peak {peak} GB/s ({src}); bytes = N*768*4 per 32-query tile")
g = torch.Generator(device="cuda").manual_seed(0)
for N in (10_000, 100_000, 1_000_000, 10_000_000):
    db = torch.randn(N, 768, device="cuda", generator=g)
    flat = engine.FlatIndex(db)
    for nq in (1, 8, 32, 64, 512, 4096):
        if N * nq > 1.1e10:
            continue
        q = db[torch.randint(0, N, (nq,), device="cuda", generator=g)] + 0.1 * torch.randn(nq, 768, device="cuda", generator=g)
        ms = timeit(lambda: engine.knn_bruteforce_top1(db, q), reps=5 if N >= 1_000_000 else 20)
        tiles = (nq + 31) // 32
        gbs = N * 768 * 4 * tiles / (ms * 1e-3) / 1e9
        D, I = engine.knn_bruteforce_top1(db, q)
        line = (f"N={N:>9} nq={nq:>5}  exact scan {ms:9.3f} ms  {gbs:8.1f} GB/s (db bytes x query tiles)  frac_of_hbm_peak={gbs / peak:5.2f}  "
                f"single-pass GB/s={N * 768 * 4 / (ms * 1e-3) / 1e9:8.1f}")
        if nq >= 32:
            # tensor-core short list + exact re-rank (rvcb_flat_*): same D / I; tensor roofline = 2*nq*N*768 flops
            ms2 = timeit(lambda: flat.search(q), reps=5 if N >= 1_000_000 else 20)
            D2, I2 = flat.search(q)
            tf = 2.0 * nq * N * 768 / (ms2 * 1e-3) / 1e12
            line += f"  | tensor short-list {ms2:9.3f} ms  {tf:7.1f} TFLOP/s  frac_of_bf16_peak={tf / tpeak:5.3f}  identical={bool(torch.equal(I, I2) and torch.equal(D, D2))}"
        print(line)
    del flat, db
lay = build_ivf_layout(SY.index_vectors(100000, 768, 0).numpy(), None, seed=0, device="cuda")
ix = engine.Index.from_oracle_layout(lay)
for nq in (1, 135, 799, 4096):
    q = torch.from_numpy(lay.vectors[:nq] + 0.05).cuda().contiguous()
    ms = timeit(lambda: ix.search_device(q, 8))
    print(f"IVF2564,Flat N=100000 nprobe=1 k=8 nq={nq:>5}: {ms:8.3f} ms   coarse bytes {2564 * 768 * 4 * ((nq + 31) // 32) / 1e6:.1f} MB")

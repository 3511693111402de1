"""Reader / writer for the on-disk index the WebUI trains (``added_IVF{n}_Flat_nprobe_1_*.index``,
web.py:544-571: ``faiss.index_factory(d, "IVF{n},Flat")`` written with ``faiss.write_index``), without faiss.

Layout (faiss/impl/index_write.cpp, IndexIVFFlat = fourcc "IwFl"; recalled from upstream, the file format
is not part of /root/reference -- SURVEY §8f-1 / Appendix F.2):
  "IwFl" | index header (d:i32, ntotal:i64, dummy:i64 x2, is_trained:u8, metric:i32)
  nlist:u64, nprobe:u64 | quantizer: "IxF2" + header + (n:u64, f32[n]) centroids
  direct map: type:u8, n:u64 (+ i64[n])
  inverted lists: "ilar" | nlist:u64 | code_size:u64 | list type "full" (u64 sizes[nlist]) or "sprs"
  (u64 n, then (id,size) pairs) | per non-empty list: codes u8[n*code_size], ids i64[n]
Also accepts the faiss-free ``.npz`` written by ``save_layout`` (keys centroids, vectors, list_off, list_ids).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass

import numpy as np


@dataclass
class IVFLayout:
    centroids: np.ndarray
    vectors: np.ndarray        # id order (what reconstruct_n returns)
    list_off: np.ndarray
    list_ids: np.ndarray
    nprobe: int = 1

    @property
    def ntotal(self): return self.vectors.shape[0]


def save_layout(path: str, lay) -> None:
    np.savez(path, centroids=lay.centroids, vectors=lay.vectors, list_off=lay.list_off, list_ids=lay.list_ids)


class _R:
    def __init__(self, b): self.b, self.o = b, 0
    def take(self, n):
        v = self.b[self.o:self.o + n]; self.o += n
        if len(v) != n: raise ValueError("truncated index file")
        return v
    def u8(self): return self.take(1)[0]
    def i32(self): return struct.unpack("<i", self.take(4))[0]
    def i64(self): return struct.unpack("<q", self.take(8))[0]
    def u64(self): return struct.unpack("<Q", self.take(8))[0]
    def fourcc(self): return bytes(self.take(4)).decode("ascii", "replace")
    def header(self):
        d = self.i32(); nt = self.i64(); self.i64(); self.i64(); tr = self.u8(); metric = self.i32()
        if metric > 1: self.take(4)        # metric_arg
        return d, nt, tr, metric


def read_index(path: str) -> IVFLayout:
    if str(path).endswith(".npz"):
        z = np.load(path)
        return IVFLayout(z["centroids"], z["vectors"], z["list_off"], z["list_ids"])
    r = _R(memoryview(open(path, "rb").read()))
    cc = r.fourcc()
    if cc != "IwFl":
        raise ValueError(f"unsupported faiss index type {cc!r} (only IVF-Flat, web.py:547)")
    d, ntotal, _, metric = r.header()
    if metric != 1:
        raise ValueError("only METRIC_L2 indexes are supported")
    nlist, nprobe = r.u64(), r.u64()
    qc = r.fourcc()
    if qc not in ("IxF2", "IxFl", "IxFI"):
        raise ValueError(f"unsupported coarse quantizer {qc!r}")
    qd, qn, _, _ = r.header()
    n = r.u64()
    centroids = np.frombuffer(r.take(4 * n), dtype="<f4").reshape(qn, qd).copy()
    dm_type = r.u8()
    dm_n = r.u64()
    r.take(8 * dm_n)
    if dm_type == 2:                       # hashtable: vector of pairs
        npairs = r.u64(); r.take(16 * npairs)
    il = r.fourcc()
    if il != "ilar":
        raise ValueError(f"unsupported inverted list type {il!r}")
    nl, code_size = r.u64(), r.u64()
    lt = r.fourcc()
    sizes = np.zeros(nl, dtype=np.int64)
    if lt == "full":
        m = r.u64(); sizes[:] = np.frombuffer(r.take(8 * m), dtype="<u8")
    elif lt == "sprs":
        m = r.u64(); pairs = np.frombuffer(r.take(8 * m), dtype="<u8").reshape(-1, 2)
        sizes[pairs[:, 0].astype(np.int64)] = pairs[:, 1]
    else:
        raise ValueError(f"unsupported list encoding {lt!r}")
    # structural validation: the device kernels index vectors[id * d] and list_ids[list_off[l]..] unchecked
    if nl != nlist or qn != nlist or qd != d:
        raise ValueError(f"inconsistent index file: nlist {nlist} / quantizer {qn}x{qd} / lists {nl} / d {d}")
    if code_size != 4 * d:
        raise ValueError(f"code_size {code_size} != 4*d: not a Flat (float32) inverted list")
    if int(sizes.sum()) != ntotal:
        raise ValueError(f"list sizes sum to {int(sizes.sum())}, header says ntotal = {ntotal}")
    if nprobe != 1:
        raise ValueError(f"index stores nprobe = {nprobe}; the search here is nprobe = 1 (the WebUI writes nprobe 1, web.py:564-571)")
    vectors = np.zeros((ntotal, d), dtype=np.float32)
    list_ids = np.full(ntotal, -1, dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    seen = np.zeros(ntotal, dtype=bool)
    for l in range(nl):
        k = int(sizes[l])
        if k == 0: continue
        codes = np.frombuffer(r.take(k * code_size), dtype="<f4").reshape(k, d)
        ids = np.frombuffer(r.take(8 * k), dtype="<i8")
        if ids.min() < 0 or ids.max() >= ntotal:
            raise ValueError(f"list {l}: vector id outside [0, {ntotal})")
        list_ids[off[l]:off[l + 1]] = ids
        vectors[ids] = codes
        seen[ids] = True
    if not seen.all():
        raise ValueError("index file does not hold every id in [0, ntotal) exactly once (reconstruct_n would be undefined)")
    return IVFLayout(centroids, vectors, off, list_ids, int(nprobe))


def write_index(path: str, lay) -> None:
    """Writes the same "IwFl" container (so a file produced here loads in faiss and vice versa)."""
    cent = np.ascontiguousarray(lay.centroids, dtype="<f4"); vec = np.ascontiguousarray(lay.vectors, dtype="<f4")
    off = np.asarray(lay.list_off, dtype=np.int64); ids = np.asarray(lay.list_ids, dtype=np.int64)
    nlist, d = cent.shape
    def hdr(dd, nt): return struct.pack("<iqqqBi", dd, nt, 1 << 20, 1 << 20, 1, 1)
    with open(path, "wb") as f:
        f.write(b"IwFl" + hdr(d, vec.shape[0]) + struct.pack("<QQ", nlist, 1))
        f.write(b"IxF2" + hdr(d, nlist) + struct.pack("<Q", cent.size) + cent.tobytes())
        f.write(struct.pack("<BQ", 0, 0))
        f.write(b"ilar" + struct.pack("<QQ", nlist, 4 * d) + b"full" + struct.pack("<Q", nlist) + np.diff(off).astype("<u8").tobytes())
        for l in range(nlist):
            li = ids[off[l]:off[l + 1]]
            if len(li):
                f.write(vec[li].tobytes()); f.write(li.astype("<i8").tobytes())

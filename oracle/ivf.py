"""Oracle: faiss ``IndexIVFFlat`` (L2, nprobe=1) search + the retrieval blend, numpy.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The search arithmetic lives in a third-party dependency absent from
/root/reference: ``faiss-cpu``, no version pin (requirements/main.txt:8).  This
file restates the published IndexIVFFlat algorithm and anchors on the
reference's call sites:

  index factory / build   web.py:544-571  ("IVF{n_ivf},Flat", n_ivf = min(int(16*sqrt(N)), N//39),
                          nprobe = 1, add in 8192-row batches, ids = add order)
  read + reconstruct_n    infer/modules/vc/pipeline.py:213-215, infer/lib/rtrvc.py:56-57
  search(npy, k=8)        infer/modules/vc/pipeline.py:126, infer/lib/rtrvc.py:172
  blend                   infer/modules/vc/pipeline.py:129-138, infer/lib/rtrvc.py:174-185

Algorithm: coarse = argmin over centroids of squared L2; fine = exact squared L2
over the vectors of that one list, ascending top-k, ties -> lower position in the
list, missing results -> (3.4028235e38, -1) like faiss.

PARITY UNPINNED: the reference holds no test, fixture or golden vector that
touches faiss, and faiss-cpu's SIMD summation order is build dependent, so this
restatement DEFINES the expected distances bit-for-bit:

    lane l in 0..31 accumulates, in fp32 with separate round-to-nearest multiply
    and add (no FMA), the squared differences of elements 128*c + 4*l + e for
    c = 0..d/128-1 (outer), e = 0..3 (inner); the 32 lane sums are combined by the
    xor-butterfly 16, 8, 4, 2, 1.

The CUDA kernel uses exactly this order (coalesced float4 loads + warp
shuffles), so distances AND indices are bit-exact against this oracle.  Against
real faiss they agree to fp32 rounding; indices can differ only on near-ties
below fp32 resolution.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

FLT_MAX = np.float32(3.4028235e38)


try:  # the same arithmetic, compiled (serial: numba's parallel runtime fights torch's OpenMP pool); the numpy version below is the definition
    import numba

    @numba.njit(cache=True, fastmath=False)
    def _l2sqr_lane_order_nb(q, v, out):
        nq, d = q.shape
        nv = v.shape[0]
        ch = d // 128
        acc = np.zeros(32, np.float32)
        tmp = np.zeros(32, np.float32)
        for i in range(nq):
            for j in range(nv):
                for l in range(32):
                    acc[l] = np.float32(0.0)
                for c in range(ch):
                    for e in range(4):
                        for l in range(32):
                            diff = q[i, 128 * c + 4 * l + e] - v[j, 128 * c + 4 * l + e]
                            acc[l] = acc[l] + diff * diff
                for s in (16, 8, 4, 2, 1):
                    for l in range(32):
                        tmp[l] = acc[l] + acc[l ^ s]
                    for l in range(32):
                        acc[l] = tmp[l]
                out[i, j] = acc[0]
except Exception:  # pragma: no cover
    numba = None


def l2sqr_lane_order(q: np.ndarray, v: np.ndarray, use_numba: bool = True) -> np.ndarray:
    """q [nq,d], v [nv,d] f32 -> [nq,nv] f32 squared L2 in the defined summation order."""
    q = np.ascontiguousarray(q, dtype=np.float32)
    v = np.ascontiguousarray(v, dtype=np.float32)
    nq, d = q.shape
    nv = v.shape[0]
    assert d % 128 == 0
    if use_numba and numba is not None:
        out = np.empty((nq, nv), dtype=np.float32)
        _l2sqr_lane_order_nb(q, v, out)
        return out
    qc = q.reshape(nq, 1, d // 128, 32, 4)
    vc = v.reshape(1, nv, d // 128, 32, 4)
    acc = np.zeros((nq, nv, 32), dtype=np.float32)
    for c in range(d // 128):
        for e in range(4):
            diff = qc[:, :, c, :, e] - vc[:, :, c, :, e]
            acc = acc + diff * diff                    # fp32 mul (rn), fp32 add (rn)
    for s in (16, 8, 4, 2, 1):
        acc = acc + acc[:, :, np.arange(32) ^ s]
    return acc[:, :, 0]


def n_ivf_for(n: int) -> int:
    return min(int(16 * np.sqrt(n)), n // 39)


class IVFFlat:
    """Duck-type of the faiss index object the pipeline uses: ``.search``, ``.reconstruct_n``,
    ``.ntotal``."""

    def __init__(self, centroids: np.ndarray, vectors: np.ndarray, assign: np.ndarray):
        self.centroids = np.ascontiguousarray(centroids, dtype=np.float32)
        self.vectors = np.ascontiguousarray(vectors, dtype=np.float32)     # id order
        self.ntotal = self.vectors.shape[0]
        self.d = self.vectors.shape[1]
        self.nlist = self.centroids.shape[0]
        self.nprobe = 1
        order = np.argsort(assign, kind="stable")                             # ids ascending inside a list
        self.list_ids = order.astype(np.int64)
        counts = np.bincount(assign, minlength=self.nlist)
        self.list_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)

    def reconstruct_n(self, i0: int, n: int) -> np.ndarray:
        return self.vectors[i0:i0 + n]

    def coarse(self, x: np.ndarray, chunk: int = 64) -> np.ndarray:
        out = np.empty(x.shape[0], dtype=np.int64)
        for s in range(0, x.shape[0], chunk):
            d = l2sqr_lane_order(x[s:s + chunk], self.centroids)
            out[s:s + chunk] = np.argmin(d, axis=1)       # first minimum
        return out

    def search(self, x: np.ndarray, k: int = 8) -> Tuple[np.ndarray, np.ndarray]:
        x = np.ascontiguousarray(x, dtype=np.float32)
        nq = x.shape[0]
        D = np.full((nq, k), FLT_MAX, dtype=np.float32)
        I = np.full((nq, k), -1, dtype=np.int64)
        lists = self.coarse(x)
        for qi in range(nq):
            a, b = self.list_off[lists[qi]], self.list_off[lists[qi] + 1]
            if b == a:
                continue
            ids = self.list_ids[a:b]
            dist = l2sqr_lane_order(x[qi:qi + 1], self.vectors[ids])[0]
            o = np.argsort(dist, kind="stable")[:k]
            D[qi, :len(o)] = dist[o]
            I[qi, :len(o)] = ids[o]
        return D, I


def brute_force_top1(x: np.ndarray, vectors: np.ndarray, chunk: int = 16) -> Tuple[np.ndarray, np.ndarray]:
    """BASELINE config #5 oracle: exact L2 top-1 over the whole database."""
    nq = x.shape[0]
    D = np.empty(nq, np.float32)
    I = np.empty(nq, np.int64)
    for s in range(0, nq, chunk):
        d = l2sqr_lane_order(x[s:s + chunk], vectors)
        I[s:s + chunk] = np.argmin(d, axis=1)
        D[s:s + chunk] = d[np.arange(d.shape[0]), I[s:s + chunk]]
    return D, I


def kmeans(vectors: np.ndarray, nlist: int, iters: int = 2, seed: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """Plain Lloyd k-means (index *construction* is a "next" row; any
    centroid set is a valid IVF index).  Returns (centroids, assignment)."""
    import torch
    rng = np.random.RandomState(seed)
    x = torch.from_numpy(np.ascontiguousarray(vectors, dtype=np.float32))
    c = x[torch.from_numpy(rng.choice(x.shape[0], nlist, replace=False))].clone()
    assign = None
    for it in range(iters + 1):
        d = (c * c).sum(1)[None, :] - 2.0 * (x @ c.t())
        assign = d.argmin(1)
        if it == iters:
            break
        cnt = torch.bincount(assign, minlength=nlist).clamp(min=1).float()
        cs = torch.zeros_like(c).index_add_(0, assign, x)
        nz = torch.bincount(assign, minlength=nlist) > 0
        c = torch.where(nz[:, None], cs / cnt[:, None], c)
    return c.numpy(), assign.numpy().astype(np.int64)


def build_ivf(vectors: np.ndarray, nlist: Optional[int] = None, seed: int = 0, exact_assign: bool = True) -> IVFFlat:
    n = vectors.shape[0]
    if nlist is None:
        nlist = n_ivf_for(n)
    cent, assign = kmeans(vectors, nlist, seed=seed)
    idx = IVFFlat(cent, vectors, assign)
    if exact_assign:
        # a vector must live in the list of its nearest centroid *under the defined
        # arithmetic* so that querying a stored vector probes its own list.
        assign = idx.coarse(np.ascontiguousarray(vectors, dtype=np.float32))
        idx = IVFFlat(cent, vectors, assign)
    return idx


def blend(feats: np.ndarray, score: np.ndarray, ix: np.ndarray, big_npy: np.ndarray, index_rate: float) -> np.ndarray:
    """pipeline.py:129-138 (fp32 path).  feats [nq,d] f32."""
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        weight = np.square(1 / score)
        weight /= weight.sum(axis=1, keepdims=True)
        npy = np.sum(big_npy[ix] * np.expand_dims(weight, axis=2), axis=1)
    return (npy * np.float32(index_rate) + np.float32(1 - index_rate) * feats).astype(np.float32)

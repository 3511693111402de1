"""IVF-Flat index construction utility (the WebUI's train_index, web.py:499-596: n_ivf = min(int(16*sqrt(N)), N//39),
"IVF{n},Flat", nprobe 1).  Construction is NOT on the inference hot path (SURVEY §8f-1 "next"); a plain Lloyd k-means on
whatever torch device is handy is enough to produce a valid (centroids, lists) layout that the device index consumes."""
from __future__ import annotations

import numpy as np
import torch

from .faiss_io import IVFLayout


def n_ivf_for(n: int) -> int:
    return min(int(16 * np.sqrt(n)), n // 39)


def build_ivf_layout(vectors: np.ndarray, nlist: int = None, iters: int = 3, seed: int = 0, device: str = None) -> IVFLayout:
    x = torch.from_numpy(np.ascontiguousarray(vectors, dtype=np.float32))
    n = x.shape[0]
    if nlist is None:
        nlist = n_ivf_for(n)
    dev = torch.device(device) if device else torch.device("cuda" if torch.cuda.is_available() else "cpu")
    xd = x.to(dev)
    g = torch.Generator().manual_seed(seed)
    c = xd[torch.randperm(n, generator=g)[:nlist].to(dev)].clone()
    assign = None
    for it in range(iters + 1):
        d = (c * c).sum(1)[None, :] - 2.0 * (xd @ c.t())
        assign = d.argmin(1)
        if it == iters:
            break
        cnt = torch.bincount(assign, minlength=nlist)
        cs = torch.zeros_like(c).index_add_(0, assign, xd)
        c = torch.where((cnt > 0)[:, None], cs / cnt.clamp(min=1)[:, None].float(), c)
    assign = assign.cpu().numpy().astype(np.int64)
    order = np.argsort(assign, kind="stable").astype(np.int64)
    off = np.concatenate([[0], np.cumsum(np.bincount(assign, minlength=nlist))]).astype(np.int64)
    return IVFLayout(c.cpu().numpy(), x.numpy(), off, order)

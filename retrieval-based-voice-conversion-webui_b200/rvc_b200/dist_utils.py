"""Multi-GPU plumbing of the batch path: utterances are independent units strided over ranks exactly like
the reference's own idiom (infer/modules/train/extract_feature_print.py:110 ``todo[i_part::n_part]``); the only collective
is the optional one-time broadcast of the parsed index from rank 0 (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch
import torch.distributed as dist

from .faiss_io import IVFLayout


def shard(items: Sequence, rank: int, world: int) -> List:
    return list(items[rank::world])


def broadcast_layout(lay, src: int = 0, device: str = "cpu") -> IVFLayout:
    """Rank ``src`` holds the layout; everyone returns an identical IVFLayout."""
    rank = dist.get_rank()
    meta = [None]
    fields = ("centroids", "vectors", "list_off", "list_ids")
    if rank == src:
        meta = [[(tuple(getattr(lay, f).shape), str(getattr(lay, f).dtype)) for f in fields]]
    dist.broadcast_object_list(meta, src=src)
    out = []
    for f, (shape, dt) in zip(fields, meta[0]):
        if rank == src:
            t = torch.from_numpy(np.ascontiguousarray(getattr(lay, f))).to(device)
        else:
            t = torch.empty(shape, dtype=getattr(torch, dt), device=device)
        dist.broadcast(t, src=src)
        out.append(t.cpu().numpy())
    return IVFLayout(*out)

"""IVF-Flat index construction (the WebUI's train_index, web.py:499-596: shuffle, MiniBatchKMeans above 2e5 rows,
n_ivf = min(int(16*sqrt(N)), N//39), "IVF{n},Flat", nprobe 1, trained_/added_ index files).  Construction is NOT on the inference
hot path (SURVEY §8f-1 "next"): ``build_ivf_layout`` is a Lloyd k-means + exact nearest-centroid assignment on the torch device
(matmul distances; library GEMM, not a hot-path kernel) producing the (centroids, lists) layout the device index consumes;
``train_index`` wraps it in the reference's file protocol."""
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


def train_index(exp_dir1: str, version19: str, logs_root: str = "logs", outside_index_root: str = None, n_cpu: int = 8,
                device: str = None, kmeans_threshold: float = 2e5, kmeans_clusters: int = 10000, train_iters: int = 10):
    """The WebUI's ``train_index`` (web.py:499-596) without faiss: same inputs (``logs/<exp>/3_feature{256,768}/*.npy``), same
    steps and progress strings (a generator of the accumulated info text), same outputs:

      * shuffle of the concatenated features (``np.random.shuffle`` on the row indices, :518-520),
      * above 2e5 rows: ``MiniBatchKMeans(n_clusters=10000, batch_size=256*n_cpu, compute_labels=False, init="random")``
        cluster centres replace the features (:521-540; scikit-learn, as in the reference),
      * ``total_fea.npy`` (:543), ``n_ivf = min(int(16*sqrt(N)), N // 39)`` (:544),
      * ``trained_IVF{n}_Flat_nprobe_1_{exp}_{ver}.index`` (coarse quantiser only, empty lists, :551-557) and
        ``added_IVF{n}_Flat_nprobe_1_{exp}_{ver}.index`` (:559-571) in the faiss ``IwFl`` container (rvc_b200/faiss_io.py),
        vectors added in order so ids are sequential, and the link into ``outside_index_root`` (:573-594).

    faiss trains the coarse quantiser with its own k-means (10 iterations, random initial centroids); here it is a Lloyd
    k-means on the torch device with the same iteration count -- a different but equally valid set of centroids: the index
    semantics (nearest centroid, exact search inside the probed list) do not depend on which k-means produced them."""
    import os
    import traceback
    exp_dir = "%s/%s" % (logs_root, exp_dir1)
    os.makedirs(exp_dir, exist_ok=True)
    feature_dir = "%s/3_feature256" % exp_dir if version19 == "v1" else "%s/3_feature768" % exp_dir
    if not os.path.exists(feature_dir):
        yield "请先进行特征提取!"
        return
    listdir_res = list(os.listdir(feature_dir))
    if len(listdir_res) == 0:
        yield "请先进行特征提取！"
        return
    infos = []
    npys = [np.load("%s/%s" % (feature_dir, name)) for name in sorted(listdir_res)]
    big_npy = np.concatenate(npys, 0)
    big_npy_idx = np.arange(big_npy.shape[0])
    np.random.shuffle(big_npy_idx)
    big_npy = big_npy[big_npy_idx]
    if big_npy.shape[0] > kmeans_threshold:
        infos.append("Trying doing kmeans %s shape to 10k centers." % big_npy.shape[0])
        yield "\n".join(infos)
        try:
            from sklearn.cluster import MiniBatchKMeans
            big_npy = MiniBatchKMeans(n_clusters=kmeans_clusters, verbose=False, batch_size=256 * n_cpu, compute_labels=False,
                                      init="random").fit(big_npy).cluster_centers_.astype(np.float32)
        except Exception:
            infos.append(traceback.format_exc())
            yield "\n".join(infos)
    np.save("%s/total_fea.npy" % exp_dir, big_npy)
    n_ivf = n_ivf_for(big_npy.shape[0])
    infos.append("%s,%s" % (big_npy.shape, n_ivf))
    yield "\n".join(infos)
    infos.append("training")
    yield "\n".join(infos)
    from .faiss_io import write_index
    lay = build_ivf_layout(big_npy, n_ivf, iters=train_iters, seed=int(np.random.randint(1 << 30)), device=device)
    empty = IVFLayout(lay.centroids, np.zeros((0, big_npy.shape[1]), np.float32), np.zeros(n_ivf + 1, np.int64), np.zeros(0, np.int64))
    write_index("%s/trained_IVF%s_Flat_nprobe_%s_%s_%s.index" % (exp_dir, n_ivf, 1, exp_dir1, version19), empty)
    infos.append("adding")
    yield "\n".join(infos)
    index_save_path = "%s/added_IVF%s_Flat_nprobe_%s_%s_%s.index" % (exp_dir, n_ivf, 1, exp_dir1, version19)
    write_index(index_save_path, lay)
    infos.append("Successfully built index into " + index_save_path)
    if outside_index_root:
        link_target = "%s/%s_IVF%s_Flat_nprobe_%s_%s_%s.index" % (outside_index_root, exp_dir1, n_ivf, 1, exp_dir1, version19)
        try:
            os.symlink(os.path.abspath(index_save_path), link_target)
            infos.append("Link index to outside folder " + link_target)
        except Exception:
            infos.append("Link index to outside folder " + link_target + " Fail")
    yield "\n".join(infos)

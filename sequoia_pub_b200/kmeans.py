"""Drop-in for the k-means step of `pre_processing/kmean_features.py:96-105`:

    kmeans = KMeans(n_clusters=num_clusters, random_state=0).fit(features)
    cluster_features[pos] = np.mean(features[np.where(kmeans.labels_ == pos)], axis=0)

`KMeans` mirrors the part of `sklearn.cluster.KMeans` the reference touches (constructor kwargs `n_clusters`,
`random_state`; `.fit(X)`; `.labels_`, `.n_iter_`) and adds `.cluster_features_` (the per-label means the script computes
next).  All arithmetic runs in `sq_kmeans_fit` (csrc/kmeans.cu); only the MT19937 stream — which does not depend on the
data — is drawn on the host with numpy's `RandomState`, exactly as `sklearn.utils.check_random_state(0)` would.
Empty clusters are relocated like sklearn does; any `n_clusters` < 2981 works (2 + int(ln k) local seeding trials <= 10);
a slide may hold up to ~40 000 tiles (the reference's `max_patch_number` default is 4000).  The C call only enqueues work
(the Lloyd loop is a device-controlled CUDA-graph while node); this wrapper synchronises when it copies the results to the host.
There is no CPU fallback.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib


class KMeans:
    def __init__(self, n_clusters=8, *, random_state=None, max_iter=300, tol=1e-4, n_init="auto", init="k-means++",
                 algorithm="lloyd", device="cuda"):
        if init != "k-means++" or algorithm != "lloyd" or n_init not in ("auto", 1):
            raise NotImplementedError("sequoia_b200 KMeans implements the reference's configuration only: "
                                      "init='k-means++', n_init='auto' (=1), algorithm='lloyd'")
        if not isinstance(random_state, (int, np.integer)):
            raise NotImplementedError("random_state must be an int (the reference passes 0)")
        self.n_clusters, self.random_state, self.max_iter, self.tol = n_clusters, int(random_state), max_iter, tol
        self.device = torch.device(device)
        self._ws = None

    def fit(self, X, y=None, sample_weight=None, _init_rows=None):
        """`_init_rows` (tests): int array [n_clusters] of rows of X used as initial centres instead of k-means++ — sklearn's
        `init=X[rows], n_init=1`; duplicate rows produce empty clusters and exercise the relocation step."""
        if sample_weight is not None:
            raise NotImplementedError("sample_weight is not used by the reference")
        _lib.require_device()
        if isinstance(X, np.ndarray):
            X = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32))
        X = X.to(device=self.device, dtype=torch.float32).contiguous()
        if X.dim() != 2:
            raise ValueError("X must be [n_samples, n_features]")
        n, d = X.shape
        k = self.n_clusters
        if n < k:
            raise ValueError(f"n_samples={n} should be >= n_clusters={k}.")          # sklearn's message
        trials = 2 + int(math.log(k))
        if trials > 10:
            raise NotImplementedError("n_clusters >= 2981 (more than 10 local seeding trials) is not supported")
        # sklearn: random_state.choice(n, p=sample_weight / sample_weight.sum()), then uniform(size=trials) per new centre
        rs = np.random.RandomState(self.random_state)
        w = np.ones(n, dtype=np.float32)
        first = int(rs.choice(n, p=w / w.sum()))
        uniforms = torch.from_numpy(rs.uniform(size=max(k - 1, 1) * trials)).to(self.device)
        L = _lib.lib()
        need = L.sq_kmeans_workspace_bytes(n, d, k)
        if self._ws is None or self._ws.numel() < need or self._ws.device != X.device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=X.device)
        labels = torch.empty(n, dtype=torch.int32, device=X.device)
        means = torch.empty(k, d, dtype=torch.float32, device=X.device)
        chosen = torch.empty(k, dtype=torch.int32, device=X.device)
        n_iter = torch.zeros(1, dtype=torch.int32, device=X.device)
        init = None
        if _init_rows is not None:
            init = torch.as_tensor(np.asarray(_init_rows), dtype=torch.int32).to(X.device)
            if init.numel() != k or int(init.min()) < 0 or int(init.max()) >= n:
                raise ValueError("_init_rows must hold n_clusters valid row indices")
        with _lib.on_device(X):
            _lib.check(L.sq_kmeans_fit(_lib.ptr(X), n, d, k, trials, first, _lib.ptr(uniforms), _lib.ptr(init), self.max_iter, self.tol,
                                       _lib.ptr(labels), _lib.ptr(means), _lib.ptr(chosen), _lib.ptr(n_iter), _lib.ptr(self._ws),
                                       self._ws.numel(), _lib.stream_ptr(X)))
        self.labels_device_, self.cluster_features_device_ = labels, means
        self.labels_ = labels.cpu().numpy()
        self.cluster_features_ = means.cpu().numpy()
        self.seed_rows_ = chosen.cpu().numpy()
        self.n_iter_ = int(n_iter.item())
        return self


def cluster_features(features, num_clusters=100, random_state=0):
    """features [n_tiles, D] -> float32 [num_clusters, D], the array kmean_features.py stores as 'cluster_features'."""
    return KMeans(n_clusters=num_clusters, random_state=random_state).fit(features).cluster_features_

"""ORACLE (test infrastructure, not product code) — CPU restatement of the per-slide k-means reduction.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

Reference call site: pre_processing/kmean_features.py:96-105
    kmeans = KMeans(n_clusters=100, random_state=0).fit(features)          # :96
    cluster_features[pos] = mean(features[where(labels_ == pos)], axis=0)   # :99-105
The arithmetic lives in the third-party dependency scikit-learn (pinned ==1.4.2 in the reference's requirements.txt:69;
1.9.0 is what this image has; same algorithm and defaults: init='k-means++', n_init='auto' -> 1, max_iter=300, tol=1e-4,
algorithm='lloyd').  This file restates that algorithm in numpy, following sklearn/cluster/_kmeans.py
(fit L1436-1563, _kmeans_plusplus L180-279, _kmeans_single_lloyd L630-758) and _k_means_lloyd.pyx / _k_means_common.pyx,
and is pinned by construction: tests/test_oracle_cpu.py asserts label equality with the installed sklearn.KMeans on
synthetic slides, and tests/golden/kmeans_golden.npz holds sklearn's own labels for the GPU tests.

Two places in sklearn's seeding are BLAS-summation-order dependent: `current_pot = closest_dist_sq @ sample_weight`
(cblas sdot) and `candidates_pot = distance_to_candidates @ sample_weight.reshape(-1, 1)` (cblas sgemv).  The candidate
draw is `searchsorted(cumsum(closest), uniform * pot)`, so a 1-ulp change of `pot` can move a draw across a bin edge;
reproducing sklearn's labels bit-exactly therefore needs the SAME fp32 summation order.  `blas_order_*` below restate
the orders of the OpenBLAS 0.3.30 x86-64 kernels numpy links here (Haswell/SkylakeX sdot and sgemv_t micro-kernels),
recovered empirically and asserted against numpy in tests/test_oracle_cpu.py; the CUDA path uses the same orders.
"""
import numpy as np

f32 = np.float32
N_LOCAL_TRIALS = lambda k: 2 + int(np.log(k))      # noqa: E731  (_kmeans.py L224-228)
GEMV_NB = 4096                                     # OpenBLAS sgemv_t processes the dot in blocks of 4096 elements


# ------------------------------------------------------------------ BLAS summation orders (w = ones)
def _fold_hadd4(v4):
    return f32(f32(v4[0] + v4[1]) + f32(v4[2] + v4[3]))


def blas_order_sdot(a):
    """cblas_sdot(a, ones): n & -32 elements in the AVX-512 kernel (4 x 16-lane accumulators over 64-element blocks,
    folded to 4 x 8 lanes, one optional 32-element block, accumulators summed in order, 8 -> 4 lanes, two hadds);
    the tail is accumulated in double and rounded once."""
    a = np.asarray(a, dtype=f32)
    n = a.shape[0]
    n32 = n & ~31
    n64 = n32 & ~63
    s = f32(0)
    if n32:
        acc16 = np.zeros((4, 16), f32)
        for r in a[:n64].reshape(-1, 4, 16):
            acc16 = (acc16 + r).astype(f32)
        acc = (acc16[:, :8] + acc16[:, 8:]).astype(f32)
        for r in a[n64:n32].reshape(-1, 4, 8):
            acc = (acc + r).astype(f32)
        t = acc[0]
        for u in range(1, 4):
            t = (t + acc[u]).astype(f32)
        s = _fold_hadd4((t[:4] + t[4:]).astype(f32))
    d = float(s)
    for x in a[n32:]:
        d += float(x)
    return f32(d)


def _gemv_block(a, kind):
    """One <=4096-element block of sgemv_t (length a multiple of 4). kind 0: the 4-column AVX2 kernel (an optional
    leading group of 4 into lanes 0-3, then 8-lane blocks; fold 8 -> 4; two hadds). kind 1: the 2-column kernel
    (4 lanes; two hadds)."""
    if kind == 0:
        acc = np.zeros(8, f32)
        i = 0
        if a.shape[0] & 4:
            acc[:4] = a[:4]
            i = 4
        for r in a[i:].reshape(-1, 8):
            acc = (acc + r).astype(f32)
        return _fold_hadd4((acc[:4] + acc[4:]).astype(f32))
    acc = np.zeros(4, f32)
    for r in a.reshape(-1, 4):
        acc = (acc + r).astype(f32)
    return _fold_hadd4(acc)


def gemv_row_kind(row, nrows):
    """Which sgemv_t micro-kernel sums row `row` of an (nrows, n) matrix: rows are taken four at a time by the 4-column
    kernel (kind 0), then a remaining pair by the 2-column kernel (kind 1), then a remaining single row by the 1-column
    kernel, whose summation order equals kind 0 (two 4-lane accumulators over 8-element groups, lo + hi, two hadds)."""
    rem = nrows % 4
    if row < nrows - rem:
        return 0
    r = row - (nrows - rem)
    return 1 if (rem & 2) and r < 2 else 0


def blas_order_gemv_row(a, row, nrows=6):
    """Row `row` of (nrows, n) @ ones(n, 1) through cblas_sgemv (nrows >= 2; nrows = 6 -> rows 0-3 kind 0, rows 4-5
    kind 1; nrows = 7 -> rows 0-3 kind 0, 4-5 kind 1, 6 kind 0); the n % 4 tail is summed left to right and added last."""
    a = np.asarray(a, dtype=f32)
    n = a.shape[0]
    kind = gemv_row_kind(row, nrows)
    m1 = n - (n & 3)
    y = f32(0)
    for b0 in range(0, m1, GEMV_NB):
        y = f32(y + _gemv_block(a[b0:min(b0 + GEMV_NB, m1)], kind))
    if n & 3:
        t = a[m1]
        for x in a[m1 + 1:]:
            t = f32(t + x)
        y = f32(y + t)
    return y


# ------------------------------------------------------------------ sklearn pipeline
def _dist_sq_upcast(C, X, xx64=None):
    """_euclidean_distances(C, X, squared=True) for float32 inputs: computed in float64 (pairwise.py
    _euclidean_distances_upcast: d = -2 C.X^T; d += ||C||^2; d += ||X||^2), cast to float32, clamped at 0."""
    C64, X64 = C.astype(np.float64), X.astype(np.float64)
    d = -2.0 * (C64 @ X64.T)
    d += np.einsum("ij,ij->i", C64, C64)[:, None]
    d += (np.einsum("ij,ij->i", X64, X64) if xx64 is None else xx64)[None, :]
    d = d.astype(f32)
    np.maximum(d, 0, out=d)
    return d


def kmeans_plusplus(X, k, seed=0, pot_mode="blas"):
    """_kmeans_plusplus (_kmeans.py L180-279) with sample_weight = ones. Returns the chosen row indices.
    pot_mode 'blas': potentials through numpy's BLAS exactly as sklearn does; 'emulated': the restated orders."""
    n = X.shape[0]
    rs = np.random.RandomState(seed)
    w = np.ones(n, dtype=X.dtype)
    trials = N_LOCAL_TRIALS(k)
    xx64 = np.einsum("ij,ij->i", X.astype(np.float64), X.astype(np.float64))
    idx = np.full(k, -1, dtype=np.int64)
    idx[0] = rs.choice(n, p=w / w.sum())
    closest = _dist_sq_upcast(X[idx[0], None], X, xx64)                       # (1, n) float32
    pot = (closest @ w)[0] if pot_mode == "blas" else blas_order_sdot(closest[0])
    for c in range(1, k):
        rand_vals = rs.uniform(size=trials) * pot                             # float64
        cand = np.searchsorted(np.cumsum(w * closest), rand_vals)             # sequential float32 cumsum
        np.clip(cand, None, n - 1, out=cand)
        d = _dist_sq_upcast(X[cand], X, xx64)
        np.minimum(closest, d, out=d)
        if pot_mode == "blas":
            pots = (d @ w.reshape(-1, 1))[:, 0]
        else:
            pots = np.array([blas_order_gemv_row(d[t], t, trials) for t in range(trials)], dtype=f32)
        best = int(np.argmin(pots))
        pot, closest, idx[c] = pots[best], d[best:best + 1], cand[best]
    return idx


def lloyd(X, centers, tol, max_iter=300):
    """_kmeans_single_lloyd (L630-758) + lloyd_iter_chunked_dense (_k_means_lloyd.pyx): per 256-row chunk
    D = ||c||^2 - 2 X C^T in float32, label = first minimum, centers = per-label means. Returns labels, n_iter."""
    n, k = X.shape[0], centers.shape[0]
    labels = np.full(n, -1, dtype=np.int32)
    labels_old = labels.copy()
    strict = False
    it = 0
    for it in range(max_iter):
        csq = np.einsum("ij,ij->i", centers, centers)
        new = np.zeros_like(centers)
        weight = np.zeros(k, dtype=X.dtype)
        for s in range(0, n, 256):
            xc = X[s:s + 256]
            d = np.tile(csq, (xc.shape[0], 1)).astype(f32)
            d += f32(-2.0) * (xc @ centers.T)
            lab = np.argmin(d, axis=1).astype(np.int32)
            labels[s:s + 256] = lab
            for i, l in enumerate(lab):
                new[l] += xc[i]
                weight[l] += 1
        if (weight == 0).any():
            relocate_empty_clusters(X, centers, new, weight, labels)
        nz = weight > 0                                        # _average_centers: empty (un-relocated) clusters keep 0
        new[nz] *= (f32(1.0) / weight[nz])[:, None]
        shift = np.sqrt(((new - centers) ** 2).sum(axis=1))
        centers = new
        if np.array_equal(labels, labels_old):
            strict = True
            break
        if (shift ** 2).sum() <= tol:
            break
        labels_old[:] = labels
    if not strict:
        csq = np.einsum("ij,ij->i", centers, centers)
        d = csq[None, :] - f32(2.0) * (X @ centers.T)
        labels = np.argmin(d, axis=1).astype(np.int32)
    return labels, it + 1


def pairwise_sum_f32(a):
    """numpy's float32 pairwise summation of a contiguous 1-D array (loops_utils.h.src, PW_BLOCKSIZE 128, 8 accumulators):
    the order `ndarray.sum(axis=1)` uses for every row of a C-contiguous matrix.  Restated for the CUDA kernel that has to
    reproduce the relocation distances; asserted equal to numpy in tests/test_oracle_cpu.py."""
    a = np.asarray(a, dtype=f32)
    n = a.shape[0]
    if n < 8:
        res = f32(0)
        for x in a:
            res = f32(res + x)
        return res
    if n <= 128:
        r = a[:8].copy()
        m = n - (n % 8)
        for i in range(8, m, 8):
            r = (r + a[i:i + 8]).astype(f32)
        res = f32(f32(f32(r[0] + r[1]) + f32(r[2] + r[3])) + f32(f32(r[4] + r[5]) + f32(r[6] + r[7])))
        for x in a[m:]:
            res = f32(res + x)
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return f32(pairwise_sum_f32(a[:n2]) + pairwise_sum_f32(a[n2:]))


def relocate_empty_clusters(X, centers_old, centers_new, weight, labels):
    """_relocate_empty_clusters_dense (_k_means_common.pyx L167-211, sample_weight = 1): every empty cluster takes the
    sample farthest from its own (old) centre, which is removed from the cluster it was assigned to.  `centers_new` holds
    the per-cluster SUMS here (the averaging follows), labels are left as they are.
    The candidates are `np.argpartition(distances, -n_empty)[:-n_empty-1:-1]`; numpy's introselect leaves the order inside
    the partition unspecified, numpy 2.x on x86-64 returns them in descending distance (checked here against numpy on 720
    random cases) and that is the order restated: empty cluster i (ascending id) takes the i-th farthest sample."""
    empty = np.where(weight == 0)[0]
    n_empty = empty.shape[0]
    if n_empty == 0:
        return
    distances = ((X - centers_old[labels]) ** 2).sum(axis=1)
    if np.max(distances) == 0:
        return
    far = np.argsort(-distances, kind="stable")[:n_empty]
    for new_id, far_idx in zip(empty, far):
        old_id = labels[far_idx]
        centers_new[old_id] -= X[far_idx]
        centers_new[new_id] = X[far_idx]
        weight[new_id] = 1
        weight[old_id] -= 1


def fit_labels(features, k=100, seed=0, pot_mode="blas"):
    """KMeans(n_clusters=k, random_state=seed).fit(features).labels_ restated (fit L1436-1563)."""
    X = np.array(features, dtype=f32, order="C", copy=True)
    tol = f32(np.mean(np.var(X, axis=0)) * 1e-4)            # _tolerance
    X -= X.mean(axis=0)
    idx = kmeans_plusplus(X, k, seed, pot_mode)
    labels, n_iter = lloyd(X, X[idx].copy(), tol)
    return labels, idx, n_iter


def cluster_means(features, labels, k=100):
    """kmean_features.py:99-105: per-label mean of the RAW features, rows in ascending order."""
    features = np.asarray(features)
    out = []
    for pos in range(k):
        out.append(np.mean(features[np.where(labels == pos)], axis=0))
    return np.asarray(out)


def make_slide_features(slide_id, n=4096, d=2048, modes=150):
    """Synthetic feature matrix with cluster structure: X = relu(modes[asg]*0.6 + 0.25*randn)."""
    rs = np.random.RandomState(1000 + slide_id)
    centers = rs.randn(modes, d).astype(f32)
    asg = rs.randint(0, modes, size=n)
    x = centers[asg] * f32(0.6) + f32(0.25) * rs.randn(n, d).astype(f32)
    return np.maximum(x, 0).astype(f32)

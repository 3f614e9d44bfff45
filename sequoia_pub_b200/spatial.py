"""Sliding-window spatial inference (second half): host-side mirror of `sliding_window_method`
(spatial_vis/visualize.py:35-102) on top of the B200 aggregators.

What the reference does: for every window origin (x, y) on a `stride` grid it selects the tiles whose grid coordinates
fall inside the 10 x 10 window (:46-49), skips windows with <= 50 tiles (:51), RE-EXTRACTS the features of every tile of
the window one tile at a time (:54-68; at stride 1 every tile is pushed through the backbone up to 100 times), zero-pads to
100 tiles (:71-74), runs the aggregator (:77-81) and attributes the window's prediction to every tile of the window
(:86-94); with stride < 10 a tile's value is the mean over the windows that cover it (:96-99).

What changes here is only the schedule: tile features are computed ONCE (the caller passes the [n_tiles, D] matrix — e.g.
from `SlideExtractor` — or a `featurize(indices)` callable), all accepted windows are pushed through the aggregator in
batches, and the per-tile means are taken with numpy's own 1-D summation order, so the returned dictionaries are identical
to the reference's.

Reference semantics kept on purpose: visualize.py hands the aggregator an UNBATCHED [100, D] tensor (:80).  With
`rearrange(x, 'b ... d -> b (...) d') + pos_emb1D` (tformer_lin.py:100, vit.py:109) that is 100 "slides" of one token each,
broadcast against the 100 positional rows, and `[0]` (:83) keeps the first — i.e. the value the reference attributes to a
window is the aggregator's output for the window's FIRST tile repeated over the 100 positions.  `reference_semantics=True`
(default) reproduces exactly that (one [1, D] row per window, broadcast inside `ViS.forward` / `ViT.forward`, 100x less
work than the reference spends on it); `reference_semantics=False` feeds the zero-padded [100, D] window as one slide,
which is what the surrounding code suggests was intended.
"""
import numpy as np
import torch

WINDOW = 10          # tiles per window side (hard-coded in the reference, :48-51,72)


def window_index(xs, ys, stride, window=WINDOW):
    """Accepted windows of visualize.py:43-51 in the reference's enumeration order (x outer, y inner).
    xs, ys: integer grid coordinates per tile (`xcoord_tf`, `ycoord_tf`).  Returns a list of int arrays: the positional
    indices of the tiles of every accepted window, ascending (the order `df[mask]` yields)."""
    xs = np.asarray(xs, dtype=np.int64)
    ys = np.asarray(ys, dtype=np.int64)
    out = []
    if xs.size == 0:
        return out
    max_x, max_y = int(xs.max()), int(ys.max())
    order = np.argsort(xs, kind="stable")
    xs_sorted = xs[order]
    for x in range(0, max_x, stride):
        lo, hi = np.searchsorted(xs_sorted, x, "left"), np.searchsorted(xs_sorted, x + window, "left")
        col = np.sort(order[lo:hi])                    # tiles with x <= xcoord_tf < x + 10, ascending index
        if col.size <= window * window // 2:
            continue                                   # no window of this column can reach 51 tiles
        cy = ys[col]
        for y in range(0, max_y, stride):
            sel = col[(cy >= y) & (cy < y + window)]
            if sel.size > (window * window) / 2:       # :51
                out.append(sel)
    return out


def sliding_window_method(df, patch_size_resized, feat_model, model, inds_gene_of_interest, stride, feat_model_type, feat_dim,
                          model_type='vis', device='cuda', *, tile_features=None, featurize=None, reference_semantics=True,
                          windows_per_batch=256):
    """Same positional arguments and return value as the reference: {gene index: {tile key: prediction}}.
    `df` needs `xcoord_tf` / `ycoord_tf` (and a default RangeIndex, which the reference's `df.iloc[ind]` also assumes).
    Exactly one of `tile_features` (tensor/array [len(df), feat_dim]) or `featurize(list of tile positions) -> [n, feat_dim]`
    supplies the backbone features; `feat_model`, `patch_size_resized`, `feat_model_type` are accepted for signature
    compatibility and passed to `featurize` when it takes keyword arguments."""
    if model_type not in ("vis", "vit"):
        raise NotImplementedError("sequoia_b200 implements the 'vis' and 'vit' aggregators (HE2RNA is out of scope)")
    if (tile_features is None) == (featurize is None):
        raise ValueError("pass exactly one of tile_features / featurize")
    keys = np.asarray(df.index)
    windows = window_index(df['xcoord_tf'].to_numpy(), df['ycoord_tf'].to_numpy(), stride)
    genes = list(inds_gene_of_interest)
    preds = {g: {} for g in genes}
    if not windows or not genes:
        return preds
    dev = torch.device(device)
    if tile_features is None:
        need = np.unique(np.concatenate(windows))
        feats = torch.zeros(len(keys), feat_dim, dtype=torch.float32, device=dev)
        feats[torch.as_tensor(need, device=dev)] = torch.as_tensor(featurize(need.tolist()), dtype=torch.float32).to(dev)
    else:
        feats = torch.as_tensor(tile_features, dtype=torch.float32).to(dev)
        if feats.shape != (len(keys), feat_dim):
            raise ValueError(f"tile_features must be [{len(keys)}, {feat_dim}]")
    gsel = torch.as_tensor(genes, dtype=torch.long, device=dev)
    out = np.empty((len(windows), len(genes)), dtype=np.float32)
    with torch.no_grad():
        for lo in range(0, len(windows), windows_per_batch):
            chunk = windows[lo:lo + windows_per_batch]
            if reference_semantics:
                x = feats[torch.as_tensor([int(w[0]) for w in chunk], device=dev)].unsqueeze(1)          # [nw, 1, D]
            else:
                x = torch.zeros(len(chunk), WINDOW * WINDOW, feat_dim, dtype=torch.float32, device=dev)    # :71-74
                for i, w in enumerate(chunk):
                    x[i, :len(w)] = feats[torch.as_tensor(w, device=dev)]
            out[lo:lo + len(chunk)] = model(x)[:, gsel].float().cpu().numpy()
    # attribute every window's prediction to its tiles (:86-94)
    if stride == WINDOW:
        for wi, w in enumerate(windows):
            for gi, g in enumerate(genes):
                v = out[wi, gi]
                d = preds[g]
                for k in keys[w]:
                    d[k] = v
        return preds
    cover = {}
    for wi, w in enumerate(windows):
        for t in w:
            cover.setdefault(int(t), []).append(wi)
    if stride < WINDOW:
        # np.mean over each tile's list (:96-99): same values, same 1-D pairwise summation order, batched by list length
        by_len = {}
        for t, ws in cover.items():
            by_len.setdefault(len(ws), []).append(t)
        for n, tiles in by_len.items():
            idx = np.array([cover[t] for t in tiles])                          # [tiles, n] window ids in enumeration order
            vals = np.ascontiguousarray(out[idx].transpose(0, 2, 1))           # [tiles, genes, n]: reduce the contiguous axis
            means = vals.mean(axis=2)
            for ti, t in enumerate(tiles):
                for gi, g in enumerate(genes):
                    preds[g][keys[t]] = means[ti, gi]
        for g in genes:                                                        # the reference's dict order: first appearance
            preds[g] = {keys[t]: preds[g][keys[t]] for t in cover}
    else:
        for t, ws in cover.items():                                            # stride > 10: lists are left as they are
            for gi, g in enumerate(genes):
                preds[g][keys[t]] = [out[wi, gi] for wi in ws]
    return preds

"""Per-slide driver mirroring the reference's two pre-processing scripts on top of the B200 kernels:

  * `extract_slide`  = the body of the slide loop of pre_processing/compute_features_hdf5.py:110-136
                        (read the tile datasets of `<slide>.hdf5` in key order, optional random.sample sub-sampling,
                        feature matrix [n_tiles, D], dataset "{feat_type}_features");
  * `reduce_slide`   = the body of pre_processing/kmean_features.py:75-108 (skip slides with fewer tiles than clusters or an
                        existing 'cluster_features' dataset, KMeans(100, random_state=0), per-label means, dataset
                        'cluster_features' [100, D] float32);
  * `run_slides`     = the outer loops with the reference's per-slide `try/except -> print -> continue` isolation
                        (compute_features_hdf5.py:141-144, kmean_features.py:107-113), sharded over ranks like --start/--end.

HDF5 access goes through `hdf5.open_file`: h5py when it is importable (the reference's own dependency), else the built-in
codec of `sequoia_pub_b200/hdf5.py` (version-0 superblock, symbol-table root group, contiguous datasets: what h5py emits
with defaults).  With the built-in codec the tiles of a slide are read straight into one pinned host buffer
(`File.read_many`, one preadv per tile) instead of 4096 dataset objects + `np.stack`.  File layout kept: patch file = one uint8 [256,256,3] dataset per tile named "{x}_{y}"
(patch_gen_hdf5.py:119-120); feature file `<feature_path>/<project>/<WSI>/<WSI>.h5` with "resnet_features"/"uni_features"
[n, D] float32 and "cluster_features" [100, D] float32 (read back by src/read_data.py:47-49).
"""
import os
import random as _random

import numpy as np

from . import hdf5
from .dist import shard_slides
from .extract import SlideExtractor, select_keys
from .kmeans import KMeans


def read_tiles(f, keys, pinned=True):
    """Tiles `keys` of an open patch file as one uint8 [n, H, W, 3] host array (compute_features_hdf5.py:117 `f_read[key][:]`
    for every key).  Built-in codec: read in place into a (pinned) torch buffer; h5py: per-dataset reads."""
    if not keys:
        return np.zeros((0, 256, 256, 3), dtype=np.uint8)
    shape = tuple(f[keys[0]].shape)
    if isinstance(f, hdf5.File):
        import torch
        buf = torch.empty((len(keys),) + shape, dtype=torch.uint8, pin_memory=bool(pinned and torch.cuda.is_available()))
        f.read_many(keys, buf)
        return buf
    out = np.empty((len(keys),) + shape, dtype=np.uint8)
    for i, k in enumerate(keys):
        f[k].read_direct(out[i])
    return out


def extract_tiles(model, tiles, batch_size=64, extractor=None):
    """uint8 [n, H, W, 3] tiles (already in key order) -> float32 [n, D]."""
    ex = extractor or SlideExtractor(model, batch_size, (tiles.shape[1], tiles.shape[2]))
    return ex(tiles)


_km_objects = {}


def reduce_features(features, num_clusters=100):
    """float32 [n, D] -> float32 [num_clusters, D] or None when the slide has fewer tiles than clusters (kmean_features.py:86-89).
    One KMeans object per cluster count is kept, so consecutive slides reuse its workspace and its cached Lloyd-loop graph."""
    if features.shape[0] < num_clusters:
        return None
    km = _km_objects.get(num_clusters)
    if km is None:
        km = _km_objects[num_clusters] = KMeans(n_clusters=num_clusters, random_state=0)
    return km.fit(features).cluster_features_


def extract_slide(model, patch_file, feature_file, feat_type="resnet", max_patch_number=4000, rng=_random, batch_size=64,
                  extractor=None, prefer_h5py=True):
    """`feat_type` is the script's --feat_type ('resnet' | 'uni', :31) and must name the model that is passed: the dataset is
    written as "{feat_type}_features" and read back under that name by kmean_features.py / read_data.py."""
    if feat_type not in ("resnet", "uni"):
        raise ValueError(f"feat_type must be 'resnet' or 'uni', got {feat_type!r}")
    if getattr(model, "feat_type", None) != feat_type:
        raise ValueError(f"feat_type={feat_type!r} but the model is a {getattr(model, 'feat_type', type(model).__name__)!r} extractor")
    with hdf5.open_file(patch_file, "r", prefer_h5py) as f:
        keys = select_keys(list(f.keys()), max_patch_number, rng)          # :111-113
        tiles = read_tiles(f, keys)                                        # :117
    feats = extract_tiles(model, tiles, batch_size, extractor)
    os.makedirs(os.path.dirname(feature_file) or ".", exist_ok=True)
    with hdf5.open_file(feature_file, "w", prefer_h5py) as f:               # :134-136
        f.create_dataset(f"{feat_type}_features", data=feats)
    return feats


def reduce_slide(feature_file, num_clusters=100, feat_name="resnet_features", prefer_h5py=True):
    with hdf5.open_file(feature_file, "r+", prefer_h5py) as f:              # :75
        if "cluster_features" in f.keys():                                  # :91-94
            return None
        feats = f[feat_name][:]
        out = reduce_features(np.asarray(feats, dtype=np.float32), num_clusters)
        if out is not None:
            f.create_dataset("cluster_features", data=out)                  # :108
    return out


def run_slides(slide_ids, fn, rank=0, world=1):
    """Applies fn(slide_id) to this rank's shard; one failing slide must not stop the others."""
    done = []
    for i in shard_slides(len(slide_ids), rank, world):
        sid = slide_ids[i]
        try:
            fn(sid)
            done.append(sid)
        except Exception as e:          # compute_features_hdf5.py:141-144
            print(e)
            print(sid)
            continue
    return done

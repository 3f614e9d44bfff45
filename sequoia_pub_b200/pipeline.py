"""Per-slide driver mirroring the reference's two pre-processing scripts on top of the B200 kernels:

  * `extract_slide`  = the body of the slide loop of pre_processing/compute_features_hdf5.py:110-136
                        (read the tile datasets of `<slide>.hdf5` in key order, optional random.sample sub-sampling,
                        feature matrix [n_tiles, D], dataset "{feat_type}_features");
  * `reduce_slide`   = the body of pre_processing/kmean_features.py:75-108 (skip slides with fewer tiles than clusters or an
                        existing 'cluster_features' dataset, KMeans(100, random_state=0), per-label means, dataset
                        'cluster_features' [100, D] float32);
  * `run_slides`     = the outer loops with the reference's per-slide `try/except -> print -> continue` isolation
                        (compute_features_hdf5.py:141-144, kmean_features.py:107-113), sharded over ranks like --start/--end.

HDF5 access uses h5py exactly like the reference (it is the reference's dependency, not ours); h5py is not installed in the
build image, so the in-memory functions (`extract_tiles`, `reduce_features`) carry the tests and the file functions raise a
clear ImportError without it.  File layout kept: patch file = one uint8 [256,256,3] dataset per tile named "{x}_{y}"
(patch_gen_hdf5.py:119-120); feature file `<feature_path>/<project>/<WSI>/<WSI>.h5` with "resnet_features"/"uni_features"
[n, D] float32 and "cluster_features" [100, D] float32 (read back by src/read_data.py:47-49).
"""
import os
import random as _random

import numpy as np

from .dist import shard_slides
from .extract import SlideExtractor, select_keys
from .kmeans import KMeans


def _h5py():
    try:
        import h5py
        return h5py
    except ImportError as e:  # pragma: no cover - depends on the deployment image
        raise ImportError("the HDF5 file functions need h5py (the reference's own dependency); the in-memory functions "
                          "extract_tiles / reduce_features do not") from e


def extract_tiles(model, tiles, batch_size=64, extractor=None):
    """uint8 [n, H, W, 3] tiles (already in key order) -> float32 [n, D]."""
    ex = extractor or SlideExtractor(model, batch_size, (tiles.shape[1], tiles.shape[2]))
    return ex(tiles)


def reduce_features(features, num_clusters=100):
    """float32 [n, D] -> float32 [num_clusters, D] or None when the slide has fewer tiles than clusters (kmean_features.py:86-89)."""
    if features.shape[0] < num_clusters:
        return None
    return KMeans(n_clusters=num_clusters, random_state=0).fit(features).cluster_features_


def extract_slide(model, patch_file, feature_file, feat_type="resnet", max_patch_number=4000, rng=_random, batch_size=64):
    h5py = _h5py()
    with h5py.File(patch_file, "r") as f:
        keys = select_keys(list(f.keys()), max_patch_number, rng)          # :111-113
        tiles = np.stack([f[k][:] for k in keys])                          # :117
    feats = extract_tiles(model, tiles, batch_size)
    os.makedirs(os.path.dirname(feature_file), exist_ok=True)
    with h5py.File(feature_file, "w") as f:                                 # :134-136
        f.create_dataset(f"{feat_type}_features", data=feats)
    return feats


def reduce_slide(feature_file, num_clusters=100, feat_name="resnet_features"):
    h5py = _h5py()
    with h5py.File(feature_file, "r+") as f:                                # :75
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

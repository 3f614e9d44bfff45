"""Device-resident replacement for the training data path: `SuperTileRNADataset.__getitem__`
(src/read_data.py:38-56) opens one HDF5 file and gathers ~20 k `rna_*` columns from a pandas row PER SAMPLE, and
`DataLoader(..., collate_fn=custom_collate_fn)` (src/main.py:120-135, src/utils.py:10-18) stacks them on the host every
step; once the train step takes milliseconds that is the bottleneck.  Here every slide's `cluster_features` [100, D] and
RNA vector [G] are loaded ONCE (any loader: the reference Dataset itself works), kept on the device, and batches are
index gathers on the GPU.

Semantics kept: samples whose features are `None` (unreadable file) are dropped like `custom_collate_fn` does; a batch is
`(features [B,100,D], rna [B,G], wsi_file_names, tcga_projects)` in the reference's order; `shuffle=True` draws a fresh
permutation per epoch like `DataLoader(shuffle=True)`; the last, smaller batch is kept (drop_last=False).
"""
import os

import numpy as np
import torch

from . import hdf5


def read_samples(df, features_path, feature_use="cluster_features", prefer_h5py=True):
    """The samples `SuperTileRNADataset(df, features_path, feature_use)[i]` would return (src/read_data.py:38-56), read in
    one pass: the `rna_*` columns are gathered ONCE for the whole frame instead of a 20k-column `row[[...]]` selection per
    item, and every `<features_path>/<tcga_project>/<wsi>/<wsi>.h5` is opened once.  Unreadable files print the exception and
    the path and yield `features=None` like the reference (:52-55)."""
    rna_cols = [c for c in df.columns if "rna_" in c]                           # :42
    rna = np.array(df[rna_cols].to_numpy(dtype=np.float32))
    for i, (name, proj) in enumerate(zip(df["wsi_file_name"], df["tcga_project"])):
        path = os.path.join(features_path, proj, name, name + ".h5")            # :40-41
        try:
            if "GTEX" not in path:
                path = path.replace(".svs", "")                                 # :45-46
            with hdf5.open_file(path, "r", prefer_h5py) as f:
                feats = torch.as_tensor(f[feature_use][:], dtype=torch.float32)
        except Exception as e:
            print(e)
            print(path)
            feats = None
        yield feats, torch.from_numpy(rna[i]), name, proj


def filter_no_features(df, feature_path, feature_name, prefer_h5py=True):
    """Rows of `df` whose feature file exists and holds a `feature_name` dataset — src/utils.py:20-40 restated over the HDF5
    shim (a directory without a readable `<wsi>.h5`, or a file without the dataset, removes the slide)."""
    remove, present = [], []
    for proj in np.unique(df["tcga_project"]):
        wsis = os.listdir(os.path.join(feature_path, proj))
        for wsi in wsis:
            try:
                with hdf5.open_file(os.path.join(feature_path, proj, wsi, wsi + ".h5"), "r", prefer_h5py) as f:
                    if feature_name not in list(f.keys()):
                        remove.append(wsi)
            except Exception:
                remove.append(wsi)
        present += wsis
    remove += df[~df["wsi_file_name"].isin(present)]["wsi_file_name"].values.tolist()
    return df[~df["wsi_file_name"].isin(remove)].reset_index(drop=True)


class DeviceSlideDataset:
    def __init__(self, samples, device="cuda"):
        """samples: iterable of (features [100,D] or None, rna [G], wsi_file_name, tcga_project) — e.g. a
        `SuperTileRNADataset` — read exactly once."""
        feats, rna, self.names, self.projects = [], [], [], []
        self.dropped = []
        for f, r, name, proj in samples:
            if f is None:                      # custom_collate_fn: remove bad entries
                self.dropped.append(name)
                continue
            feats.append(torch.as_tensor(f, dtype=torch.float32))
            rna.append(torch.as_tensor(r, dtype=torch.float32))
            self.names.append(name)
            self.projects.append(proj)
        if not feats:
            raise ValueError("no readable samples")
        self.device = torch.device(device)
        self.features = torch.stack(feats).to(self.device)          # [n, 100, D]   (0.8 MB per slide at D = 2048)
        self.rna = torch.stack(rna).to(self.device)                 # [n, G]
        self.num_genes = self.rna.shape[1]
        self.feature_dim = self.features.shape[2]

    @classmethod
    def from_files(cls, df, features_path, feature_use="cluster_features", device="cuda", prefer_h5py=True):
        return cls(read_samples(df, features_path, feature_use, prefer_h5py), device)

    def __len__(self):
        return self.features.shape[0]

    def batches(self, batch_size, shuffle=False, generator=None, rank=0, world=1):
        """Yields (features, rna, names, projects); with world > 1 every rank gets an equal slice of each global batch
        (global batches that cannot be split evenly are truncated to a multiple of `world`)."""
        n = len(self)
        order = torch.randperm(n, generator=generator) if shuffle else torch.arange(n)
        for lo in range(0, n, batch_size):
            idx = order[lo:lo + batch_size]
            if world > 1:
                per = idx.numel() // world
                if per == 0:
                    continue
                idx = idx[rank * per:(rank + 1) * per]
            di = idx.to(self.device)
            yield (self.features.index_select(0, di), self.rna.index_select(0, di), [self.names[i] for i in idx.tolist()],
                   [self.projects[i] for i in idx.tolist()])

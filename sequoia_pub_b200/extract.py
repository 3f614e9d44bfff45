"""Host-side mirror of the per-slide feature-extraction loop of
pre_processing/compute_features_hdf5.py:110-131, re-expressed as a batched, double-buffered pipeline.

Reference behaviour kept:
  * tile order: features are produced in key order (`list(f.keys())`, :111) — row order matters downstream
    because k-means++ picks seeds by row index;
  * sub-sampling: more than `max_patch_number` tiles -> `random.sample(keys, max_patch_number)` with the module
    level `random` generator the script seeded with `random.seed(args.seed)` (:41,112-113);
  * output: float32 [n_tiles, D] (`np.asarray(features_tiles)`, :131).

What changes is only HOW the tiles reach the model: the reference moves one fp32 CHW tile per forward and
synchronises on every tile (:119-123); here uint8 HWC tiles go to the device in batches from pinned memory on a copy
stream while the previous batch is in the extractor, and one D2H copy returns the [n, D] matrix.
"""
import random as _random

import numpy as np
import torch


def select_keys(keys, max_patch_number=4000, rng=_random):
    """compute_features_hdf5.py:111-113."""
    keys = list(keys)
    if len(keys) > max_patch_number:
        keys = rng.sample(keys, max_patch_number)
    return keys


class SlideExtractor:
    """Runs the model's uint8 extractor over all tiles of a slide; the model is either extractor of the reference script
    (`--feat_type resnet`: `resnet.ResNet`, 2048 features; `--feat_type uni`: `uni.VisionTransformer`, 1024 features, with the
    script's `Resize(224)` of the 256-px tiles done on the GPU) - anything exposing `feature_dim`, `new_lane_workspace` and
    `extract_tiles_into`.  H2D copies (pinned staging, copy stream) overlap compute, and
    consecutive batches alternate between two compute lanes (stream + extractor workspace each), so the persistent
    convolution kernels of one batch fill the SMs the other leaves idle in partial waves and launch gaps.  There are twice as
    many device tile buffers as lanes: the tiles of a lane's NEXT batch are already on the device when its current batch
    finishes (with one buffer per lane every batch would wait for its own H2D copy, ~8 % of the lane's time)."""

    LANES = 2
    NBUF = 4

    def __init__(self, model, batch_size=64, tile_hw=(256, 256), device=None):
        self.model = model
        self.bs = batch_size
        self.device = torch.device(device if device is not None else "cuda")
        h, w = tile_hw
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.lanes = [torch.cuda.Stream(device=self.device) for _ in range(self.LANES)]
        n = self.NBUF
        self.dev_buf = [torch.empty(batch_size, h, w, 3, dtype=torch.uint8, device=self.device) for _ in range(n)]
        self.pin_buf = [torch.empty(batch_size, h, w, 3, dtype=torch.uint8, pin_memory=True) for _ in range(n)]
        self.copied = [torch.cuda.Event() for _ in range(n)]
        self.consumed = [torch.cuda.Event() for _ in range(n)]
        self.workspaces = [None] * self.LANES
        self._ws_hw = None
        self.feature_dim = int(model.feature_dim)
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _workspace(self, slot, h, w):
        # the model sizes its own lane state (ResNet: conv workspace; UNI: ViT workspace + the Resize(224) buffers)
        if self.workspaces[slot] is None or self._ws_hw != (h, w):
            self.workspaces = [None] * self.LANES
            self._ws_hw = (h, w)
        if self.workspaces[slot] is None:
            self.workspaces[slot] = self.model.new_lane_workspace(self.bs, h, w, self.device)
        return self.workspaces[slot]

    def __call__(self, tiles):
        """tiles: host uint8 [n, H, W, 3] (numpy array or CPU tensor, pinned or pageable) -> np.float32 [n, D]."""
        if isinstance(tiles, np.ndarray):
            tiles = torch.from_numpy(tiles)
        if tiles.dtype != torch.uint8 or tiles.dim() != 4 or tiles.shape[3] != 3:
            raise ValueError("tiles must be uint8 [n, H, W, 3]")
        n = tiles.shape[0]
        if n == 0:
            return np.zeros((0, self.feature_dim), dtype=np.float32)
        pinned = tiles.is_pinned()
        main = torch.cuda.current_stream(self.device)
        out = torch.empty(n, self.feature_dim, dtype=torch.float32, device=self.device)
        self.model._prepack()
        for s in self.lanes:
            s.wait_stream(main)
        nb = (n + self.bs - 1) // self.bs
        L, NB = self.LANES, self.NBUF
        for b in range(nb):
            lo, hi = b * self.bs, min(n, (b + 1) * self.bs)
            slot, li = b % NB, b % L
            with torch.cuda.stream(self.copy_stream):
                if b >= NB:
                    self.copy_stream.wait_event(self.consumed[slot])    # device buffer free again (batch b - NBUF is done)
                src = tiles[lo:hi]
                if not pinned:
                    if b >= NB:
                        self.copied[slot].synchronize()                 # staging buffer free again
                    self.pin_buf[slot][: hi - lo].copy_(src)
                    src = self.pin_buf[slot][: hi - lo]
                self.dev_buf[slot][: hi - lo].copy_(src, non_blocking=True)
                self.copied[slot].record(self.copy_stream)
            lane = self.lanes[li]
            with torch.cuda.stream(lane):
                lane.wait_event(self.copied[slot])
                buf = self.dev_buf[slot][: hi - lo]
                self.model.extract_tiles_into(buf, out[lo:hi], self._workspace(li, buf.shape[1], buf.shape[2]))
                self.consumed[slot].record(lane)
            self.h2d_bytes += (hi - lo) * tiles[0].numel()
        for s in self.lanes:
            main.wait_stream(s)
        host = torch.empty(out.shape, dtype=torch.float32, pin_memory=True)   # straight from the pinned allocator: no pageable copy
        host.copy_(out, non_blocking=True)
        main.synchronize()
        self.d2h_bytes += host.numel() * 4
        return host.numpy()


def extract_features(model, tiles, batch_size=64):
    """One-shot convenience wrapper (allocates the staging buffers each call)."""
    t = tiles if not isinstance(tiles, np.ndarray) else torch.from_numpy(tiles)
    return SlideExtractor(model, batch_size, (t.shape[1], t.shape[2]))(t)

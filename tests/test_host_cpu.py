"""CPU: host-side logic of the extraction driver that mirrors pre_processing/compute_features_hdf5.py:110-113."""
import random


def test_select_keys_matches_the_reference_subsampling():
    from sequoia_pub_b200.extract import select_keys
    keys = [f"{x}_{y}" for x in range(0, 9000, 256) for y in range(0, 9000, 256)]      # tile datasets are named "{x}_{y}"
    assert select_keys(keys, max_patch_number=len(keys)) == keys                        # all tiles, key order kept
    random.seed(99)                                                                     # the script seeds `random` (:41)
    want = random.sample(keys, 400)                                                     # :112-113
    random.seed(99)
    assert select_keys(keys, max_patch_number=400) == want
    rng = random.Random(5)
    assert select_keys(keys, 10, rng) == random.Random(5).sample(keys, 10)
    assert select_keys([], 10) == []


def test_product_modules_refuse_to_run_without_a_gpu():
    """No CPU fallback anywhere: every product entry point raises instead of computing on the host."""
    import numpy as np
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    from sequoia_pub_b200 import metrics
    from sequoia_pub_b200.kmeans import KMeans
    from sequoia_pub_b200.resnet import resnet50
    with pytest.raises(RuntimeError):
        resnet50().eval().forward_extract(torch.zeros(1, 3, 256, 256))
    with pytest.raises((RuntimeError, AssertionError)):
        KMeans(n_clusters=100, random_state=0).fit(np.zeros((200, 64), np.float32))
    with pytest.raises(RuntimeError):
        metrics.step_metrics(torch.zeros(4, 8), torch.zeros(4, 8))

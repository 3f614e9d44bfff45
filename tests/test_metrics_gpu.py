"""Device-side step metrics vs the reference's compute_correlations (golden from src/he2rna.py) and sklearn's MAE."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_metrics_oracle_matches_reference_golden():
    from oracle import metrics_oracle as MO
    g = np.load(os.path.join(GOLD, "metrics_golden.npz"))
    for tag, seed, b, n in (("small", 1, 5, 301), ("b2", 2, 2, 64)):
        y, p = MO.make_batch(seed, b, n)
        assert abs(MO.compute_correlations(y, p) - float(g[f"{tag}_corr"])) < 1e-12
        assert abs(MO.mean_absolute_error(y, p) - float(g[f"{tag}_mae"])) < 1e-6
        assert abs(MO.smape(y, p) - float(g[f"{tag}_smape"])) <= 1e-9 * float(g[f"{tag}_smape"])


@pytest.mark.gpu
@pytest.mark.parametrize("tag,seed,b,n", [("cfg3", 0, 32, 20530), ("small", 1, 5, 301), ("b2", 2, 2, 64)])
def test_step_metrics_match_reference(tag, seed, b, n):
    import torch
    from oracle import metrics_oracle as MO
    from sequoia_pub_b200 import metrics
    g = np.load(os.path.join(GOLD, "metrics_golden.npz"))
    y, p = MO.make_batch(seed, b, n)
    out = metrics.step_metrics(torch.from_numpy(y).cuda(), torch.from_numpy(p).cuda()).cpu().numpy()
    assert abs(out[1] - float(g[f"{tag}_corr"])) < 1e-6          # fp64 accumulation, fp32 result
    assert abs(out[0] - float(g[f"{tag}_mae"])) < 1e-6 * max(1.0, float(g[f"{tag}_mae"]))
    assert abs(metrics.compute_correlations(y, p) - float(g[f"{tag}_corr"])) < 1e-6      # numpy in, float out (reference signature)
    assert out[2] > 0
    assert abs(out[3] - float(g[f"{tag}_smape"])) < 2e-5 * float(g[f"{tag}_smape"])      # evaluate()'s smape (src/vit.py:32-33,269)
    assert abs(metrics.smape(y, p) - float(g[f"{tag}_smape"])) < 2e-5 * float(g[f"{tag}_smape"])

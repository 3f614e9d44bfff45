"""ORACLE (test infrastructure, not product code) — CPU restatement of the per-step training metrics.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

* compute_correlations: src/he2rna.py:140-149 (called every step at src/vit.py:168,268): mean over genes of the Pearson
  correlation between labels[:, g] and preds[:, g] (np.corrcoef, float64), skipping genes whose labels are constant in the
  batch and dropping NaNs (constant predictions).
* mean_absolute_error: sklearn.metrics.mean_absolute_error(labels, preds) at src/vit.py:167 (uniform average over genes).
Pinned: tests/golden/gen_golden.py executes the reference's own compute_correlations (extracted from src/he2rna.py with
ast, because importing the module needs tkinter) and stores its outputs in tests/golden/metrics_golden.npz.
"""
import numpy as np


def compute_correlations(labels, preds):
    metrics = []
    for i in range(labels.shape[1]):
        y_true = labels[:, i]
        if len(np.unique(y_true)) > 1:
            metrics.append(np.corrcoef(y_true, preds[:, i])[0, 1])
    metrics = np.asarray(metrics)
    metrics = metrics[~np.isnan(metrics)]
    return np.mean(metrics)


def mean_absolute_error(labels, preds):
    return float(np.mean(np.abs(np.asarray(labels, dtype=np.float64) - np.asarray(preds, dtype=np.float64))))


def smape(A, F):
    """src/vit.py:32-33, verbatim semantics (float32 arrays in evaluate(), :269)."""
    return 100 / len(A) * np.sum(2 * np.abs(F - A) / (np.abs(A) + np.abs(F)))


def make_batch(seed, batch=32, genes=20530):
    """Synthetic (labels, preds): log-FPKM-like labels with some constant genes, noisy predictions, a few constant columns."""
    rs = np.random.RandomState(seed)
    y = (rs.rand(batch, genes) * 10).astype(np.float32)
    y[:, ::17] = y[0, ::17]                     # genes that are constant in this batch -> skipped
    p = (y * 0.6 + rs.randn(batch, genes) * 2 + 1).astype(np.float32)
    p[:, 5::23] = 3.25                          # constant predictions -> NaN correlation -> dropped
    return y, p

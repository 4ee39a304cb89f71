"""Shared helpers for the parity tests."""
import numpy as np


def canon(boxes, scores):
    """Canonical row order for comparing detection *sets*: score desc, then box coordinates."""
    boxes = np.asarray(boxes, dtype=np.float32).reshape(-1, 4)
    scores = np.asarray(scores, dtype=np.float32).reshape(-1)
    order = np.lexsort((boxes[:, 3], boxes[:, 2], boxes[:, 1], boxes[:, 0], -scores.astype(np.float64)))
    return boxes[order], scores[order]


def random_boxes(rng, n, extent=400.0, size_lo=4.0, size_hi=120.0, distinct_scores=True):
    """Overlapping xyxy fp32 boxes with optional pairwise-distinct scores."""
    cx = rng.uniform(0, extent, n)
    cy = rng.uniform(0, extent, n)
    w = rng.uniform(size_lo, size_hi, n)
    h = rng.uniform(size_lo, size_hi, n)
    boxes = np.stack((cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2), axis=1).astype(np.float32)
    if distinct_scores:
        scores = rng.permutation(n).astype(np.float32)
        scores = ((scores + 1.0) / (n + 1.0)).astype(np.float32)
        assert np.unique(scores).shape[0] == n
    else:
        scores = (np.round(rng.uniform(0, 1, n) * 50) / 50).astype(np.float32)
    return boxes, scores


def clustered_boxes(rng, n, clusters=40, extent=1000.0, jitter=6.0, size=80.0):
    """Heavily overlapping boxes: long suppression chains."""
    centers = rng.uniform(0, extent, (clusters, 2))
    which = rng.randint(0, clusters, n)
    c = centers[which] + rng.randn(n, 2) * jitter
    wh = size * np.exp(rng.randn(n, 2) * 0.15)
    boxes = np.concatenate((c - wh / 2, c + wh / 2), axis=1).astype(np.float32)
    scores = ((rng.permutation(n) + 1.0) / (n + 1.0)).astype(np.float32)
    return boxes, scores

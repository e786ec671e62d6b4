"""Oracle for Improved Precision & Recall (metrics/ipr.py) -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

numpy restatement of the metric part of the reference (the feature extractor is torchvision's VGG-16, used directly):

  pairwise_distances   metrics/ipr.py:184-217   ||x||^2 - 2 x.y + ||y||^2 in float64, negatives clamped to 0, sqrt
  kth_value            metrics/ipr.py:228-233   (k+1)-th smallest entry of a row (the closest one is the row itself)
  distances2radii      metrics/ipr.py:220-225
  compute_metric       metrics/ipr.py:236-242   fraction of subjects inside at least one reference ball
  realism              metrics/ipr.py:253-263
  vgg16_fc2            metrics/ipr.py:139-141   vgg16.features -> view(-1, 7*7*512) -> classifier[:4]
Pinned by tests/golden/ipr_seed0.npz, produced by the reference's own functions (tests/golden/make_golden.py).
"""
from __future__ import annotations

import numpy as np


def pairwise_distances(X: np.ndarray, Y: np.ndarray | None = None) -> np.ndarray:
    X = X.astype(np.float64)                                   # :197 "to prevent underflow"
    xs = np.sum(X ** 2, axis=1, keepdims=True)                 # :198
    if Y is None:
        Y, ys = X, xs
    else:
        ys = np.sum(Y ** 2, axis=1, keepdims=True)             # :202 (Y keeps its own dtype, as in the reference)
    d2 = xs - 2 * np.dot(X, Y.T) + ys.T                        # :203-208
    d2[d2 < 0] = 0                                             # :211-214
    return np.sqrt(d2)                                         # :216


def kth_value(row: np.ndarray, k: int) -> float:
    kprime = k + 1                                             # :229
    idx = np.argpartition(row, kprime)                         # :230
    return row[idx[:kprime]].max()                             # :231-232


def distances2radii(distances: np.ndarray, k: int = 3) -> np.ndarray:
    return np.array([kth_value(distances[i], k) for i in range(distances.shape[0])])    # :220-225


def compute_metric(ref_features: np.ndarray, ref_radii: np.ndarray, subject_features: np.ndarray) -> float:
    dist = pairwise_distances(ref_features, subject_features)  # :238
    count = 0
    for i in range(subject_features.shape[0]):
        count += (dist[:, i] < ref_radii).any()                # :240
    return count / subject_features.shape[0]                   # :241


def realism(ref_features: np.ndarray, ref_radii: np.ndarray, feat: np.ndarray) -> float:
    dists = np.linalg.norm(ref_features - feat, axis=1)        # :256-258
    return float((ref_radii / (dists + 1e-6)).max())           # :259-262


def vgg16_fc2(state_dict, x):
    """x: float32 [N,3,224,224] (already normalised) -> [N,4096] with torchvision's own VGG-16 module."""
    import torch
    import torchvision
    m = torchvision.models.vgg16(weights=None)
    m.load_state_dict(state_dict)
    m.eval()
    with torch.no_grad():
        before_fc = m.features(x).reshape(-1, 7 * 7 * 512)     # :139-140 (`view` there; same values)
        return m.classifier[:4](before_fc)                     # :141


def make_vgg16_state_dict(seed: int = 0):
    """Seeded He-style weights with torchvision VGG-16's names (the real checkpoint cannot be downloaded here)."""
    import torch
    import torchvision
    g = torch.Generator().manual_seed(seed)
    sd = torchvision.models.vgg16(weights=None).state_dict()
    for k, v in sd.items():
        if v.dim() == 4:
            v.copy_(torch.randn(v.shape, generator=g) * (2.0 / (v.shape[1] * 9)) ** 0.5)
        elif v.dim() == 2:
            v.copy_(torch.randn(v.shape, generator=g) * (2.0 / v.shape[1]) ** 0.5)
        else:
            v.copy_(0.05 * torch.randn(v.shape, generator=g))
    return sd

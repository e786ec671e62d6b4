"""Oracle for the FID statistics (test infrastructure only; see oracle/__init__.py).

  activation_statistics   metrics/fid.py:199-221   mu = np.mean(act, axis=0); sigma = np.cov(act, rowvar=False)  (float64)
  frechet_distance        metrics/fid.py:142-196   ||mu1-mu2||^2 + Tr(s1) + Tr(s2) - 2 Tr(sqrtm(s1.s2)), eps retry,
                                                   imaginary-part check.  The reference calls
                                                   `linalg.sqrtm(..., disp=False)`, which scipy >= 1.18 rejects
                                                   (SURVEY.md 8c); `linalg.sqrtm(a)` is the same computation.
"""
from __future__ import annotations

import numpy as np
from scipy import linalg


def activation_statistics(act: np.ndarray):
    act = np.asarray(act, dtype=np.float64)          # pred_arr is float64 (metrics/fid.py:118)
    return np.mean(act, axis=0), np.cov(act, rowvar=False)


def frechet_distance(mu1, sigma1, mu2, sigma2, eps: float = 1e-6) -> float:
    mu1, mu2 = np.atleast_1d(mu1), np.atleast_1d(mu2)
    sigma1, sigma2 = np.atleast_2d(sigma1), np.atleast_2d(sigma2)
    assert mu1.shape == mu2.shape and sigma1.shape == sigma2.shape
    diff = mu1 - mu2
    covmean = linalg.sqrtm(sigma1.dot(sigma2))
    if not np.isfinite(covmean).all():
        offset = np.eye(sigma1.shape[0]) * eps
        covmean = linalg.sqrtm((sigma1 + offset).dot(sigma2 + offset))
    if np.iscomplexobj(covmean):
        if not np.allclose(np.diagonal(covmean).imag, 0, atol=1e-3):
            raise ValueError("Imaginary component {}".format(np.max(np.abs(covmean.imag))))
        covmean = covmean.real
    return float(diff.dot(diff) + np.trace(sigma1) + np.trace(sigma2) - 2 * np.trace(covmean))

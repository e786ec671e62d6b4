"""CPU: oracle Inception vs the golden vector produced by the reference module; Frechet distance restatements agree."""
import os

import numpy as np
import torch

from oracle import fid as ofid
from oracle import models as om

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_inception_oracle_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "fid_inception_seed0.npz"))
    img = torch.randint(0, 256, (2, 299, 299, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(8000))
    y = om.fid_inception_forward(om.make_inception_state_dict(0), om.fid_preprocess(img)).numpy()
    np.testing.assert_allclose(y, g["out"], rtol=0, atol=2e-5)


def test_frechet_eigh_formulation_matches_sqrtm_oracle():
    from dcr_b200 import fid as dfid
    rng = np.random.default_rng(0)
    d = 96
    a = rng.standard_normal((400, d)) @ rng.standard_normal((d, d)) * 0.3
    b = rng.standard_normal((500, d)) @ rng.standard_normal((d, d)) * 0.3 + 0.2
    m1, s1 = ofid.activation_statistics(a)
    m2, s2 = ofid.activation_statistics(b)
    ref = ofid.frechet_distance(m1, s1, m2, s2)
    got = dfid.frechet_distance(m1, s1, m2, s2, device=torch.device("cpu"))
    assert abs(got - ref) < 1e-6 * max(1.0, abs(ref)), (got, ref)
    assert abs(dfid.frechet_distance(m1, s1, m1, s1, device=torch.device("cpu"))) < 1e-6


def test_statistics_are_numpy_mean_cov():
    x = np.random.default_rng(1).standard_normal((50, 7)).astype(np.float32)
    mu, sig = ofid.activation_statistics(x)
    np.testing.assert_allclose(mu, x.astype(np.float64).mean(0))
    np.testing.assert_allclose(sig, np.cov(x.astype(np.float64), rowvar=False))

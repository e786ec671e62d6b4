"""Improved precision & recall (metrics/ipr.py): oracle vs the reference's own functions (golden), CUDA path vs oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import ipr as oipr

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _sets(seed=0):
    rng = np.random.default_rng(900 + seed)
    ref = (rng.standard_normal((300, 64)) * 3 + 1).astype(np.float32)
    sub = (rng.standard_normal((200, 64)) * 3.3 + 1.2).astype(np.float32)
    return ref, sub


def test_ipr_oracle_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "ipr_seed0.npz"))
    ref, sub = _sets()
    r_ref = oipr.distances2radii(oipr.pairwise_distances(ref), 3)
    r_sub = oipr.distances2radii(oipr.pairwise_distances(sub), 3)
    np.testing.assert_allclose(r_ref, g["radii_ref"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(r_sub, g["radii_sub"], rtol=0, atol=1e-12)
    assert oipr.compute_metric(ref, r_ref, sub) == float(g["precision"])
    assert oipr.compute_metric(sub, r_sub, ref) == float(g["recall"])
    real = [oipr.realism(ref, r_ref, sub[i:i + 1]) for i in range(8)]
    np.testing.assert_allclose(real, g["realism"], rtol=1e-12)


@pytest.mark.gpu
def test_ipr_cuda_metric_matches_reference_golden():
    from dcr_b200 import ipr
    g = np.load(os.path.join(GOLD, "ipr_seed0.npz"))
    ref, sub = _sets()
    r_ref, r_sub = ipr.kth_nn_radii(ref, 3), ipr.kth_nn_radii(sub, 3)
    np.testing.assert_allclose(r_ref, g["radii_ref"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(r_sub, g["radii_sub"], rtol=0, atol=1e-9)
    assert ipr.compute_metric(ipr.Manifold(ref, r_ref), sub) == float(g["precision"])
    assert ipr.compute_metric(ipr.Manifold(sub, r_sub), ref) == float(g["recall"])
    real = [ipr.realism(ipr.Manifold(ref, r_ref), sub[i:i + 1]) for i in range(8)]
    # the reference takes these norms in float32 (numpy float32 inputs, ipr.py:256-258); here they are float64
    np.testing.assert_allclose(real, g["realism"], rtol=1e-6)


@pytest.mark.gpu
def test_ipr_cuda_metric_at_vgg_dim():
    """4096-d features (the real fc2 width), 3000 x 2000 rows, against the numpy oracle."""
    from dcr_b200 import ipr
    rng = np.random.default_rng(5)
    base = np.abs(rng.standard_normal((1, 4096))).astype(np.float32) * 2          # ReLU-network-like common offset
    ref = (base + np.abs(rng.standard_normal((3000, 4096))) * 1.5).astype(np.float32)
    sub = (base + np.abs(rng.standard_normal((2000, 4096))) * 1.6).astype(np.float32)
    r_ref = ipr.kth_nn_radii(ref, 3)
    o_ref = oipr.distances2radii(oipr.pairwise_distances(ref), 3)
    np.testing.assert_allclose(r_ref, o_ref, rtol=1e-10, atol=1e-9)
    got = ipr.compute_metric(ipr.Manifold(ref, r_ref), sub)
    want = oipr.compute_metric(ref, o_ref, sub)
    assert abs(got - want) < 1e-12, (got, want)


@pytest.mark.gpu
def test_vgg16_fc2_features_match_torchvision_module():
    from dcr_b200 import ipr, nets, synthetic
    sd = oipr.make_vgg16_state_dict(0)
    img = synthetic.images(3, seed=61, size=224)
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    x = (img.permute(0, 3, 1, 2).float().div(255.0) - torch.tensor(mean).view(1, 3, 1, 1)) / torch.tensor(std).view(1, 3, 1, 1)
    ref = oipr.vgg16_fc2(sd, x)
    net = nets.build_vgg16_fc2(sd, max_batch=2, precision="exact")
    got = net(img.cuda()).cpu()
    assert (got - ref).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item())
    assert torch.equal(net(x.cuda()).cpu(), got)                     # the float32 NCHW entry, same kernels
    fast = nets.build_vgg16_fc2(sd, max_batch=4, precision="fast")
    gq = fast(img.cuda()).cpu()
    cos = torch.nn.functional.cosine_similarity(gq, ref, dim=1).min().item()
    assert cos > 0.999, cos
    # the IPR object end to end on tensors (uint8 path), against oracle features of the same network
    obj = ipr.IPR(batch_size=2, k=1, model=net)
    man = obj.compute_manifold(img)
    np.testing.assert_allclose(man.radii, oipr.distances2radii(oipr.pairwise_distances(man.features), 1), rtol=1e-9)

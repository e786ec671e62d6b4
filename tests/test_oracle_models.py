"""Oracle networks vs golden vectors produced by the reference's own modules / vs torchvision (CPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import models as om

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _golden_inputs(seed, n=2, size=224):
    return torch.randn(n, 3, size, size, generator=torch.Generator().manual_seed(7000 + seed))


@pytest.mark.parametrize("seed", [0, 1])
def test_vit_oracle_matches_reference_golden(seed):
    g = np.load(os.path.join(GOLD, f"dino_vits16_seed{seed}.npz"))
    x = _golden_inputs(seed)
    assert abs(float(x.double().sum()) - float(g["in_checksum"])) < 1e-6
    y = om.vit_forward(om.make_vit_state_dict(seed), x).numpy()
    np.testing.assert_allclose(y, g["out"], rtol=0, atol=2e-5)


def test_sscd_oracle_matches_torchvision_module():
    """The functional restatement equals a torchvision ResNet-50 module with the same weights + GeM/Linear/L2 head."""
    import torchvision
    sd = om.make_sscd_state_dict(3)
    m = torchvision.models.resnet50(weights=None)
    tv = {k[len("backbone."):]: v for k, v in sd.items() if k.startswith("backbone.")}
    tv["fc.weight"], tv["fc.bias"] = m.fc.weight.detach(), m.fc.bias.detach()
    m.load_state_dict(tv)
    m.eval()
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        f = m.layer4(m.layer3(m.layer2(m.layer1(m.maxpool(m.relu(m.bn1(m.conv1(x))))))))
        f = f.clamp(min=1e-6).pow(3).mean(dim=(2, 3)).pow(1.0 / 3)
        f = torch.nn.functional.linear(f, sd["embeddings.1.weight"], sd["embeddings.1.bias"])
        ref = torch.nn.functional.normalize(f, dim=1)
    got = om.sscd_forward(sd, x)
    np.testing.assert_allclose(got.numpy(), ref.numpy(), atol=2e-5)
    np.testing.assert_allclose(got.norm(dim=1).numpy(), 1.0, atol=1e-6)


def test_preprocess_matches_torchvision_transform():
    from torchvision import transforms
    from PIL import Image
    img = torch.randint(0, 256, (256, 256, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(2))
    tf = transforms.Compose([transforms.Resize(256), transforms.CenterCrop(224), transforms.ToTensor(),
                             transforms.Normalize([0.5, 0.5, 0.5], [0.5, 0.5, 0.5])])   # diff_retrieval.py:325-330
    ref = tf(Image.fromarray(img.numpy()))
    got = om.preprocess(img[None])[0]
    assert torch.equal(got, ref)


def test_sscd_oracle_grouped_trunk_matches_torchvision_module():
    """ResNeXt-style trunk (grouped 3x3 convs; the family upstream documents for sscd_disc_large) through the same
    functional restatement."""
    from torchvision.models.resnet import Bottleneck, ResNet
    sd = om.make_sscd_state_dict(5, dims=1024, arch="resnext_tiny")
    m = ResNet(Bottleneck, [2, 2, 2, 2], groups=8, width_per_group=8)
    tv = {k[len("backbone."):]: v for k, v in sd.items() if k.startswith("backbone.")}
    tv["fc.weight"], tv["fc.bias"] = m.fc.weight.detach(), m.fc.bias.detach()
    m.load_state_dict(tv)
    m.eval()
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        f = m.layer4(m.layer3(m.layer2(m.layer1(m.maxpool(m.relu(m.bn1(m.conv1(x))))))))
        f = f.clamp(min=1e-6).pow(3).mean(dim=(2, 3)).pow(1.0 / 3)
        f = torch.nn.functional.linear(f, sd["embeddings.1.weight"], sd["embeddings.1.bias"])
        ref = torch.nn.functional.normalize(f, dim=1)
    got = om.sscd_forward(sd, x)
    assert got.shape == (2, 1024)
    np.testing.assert_allclose(got.numpy(), ref.numpy(), atol=2e-5)


def test_vit_other_input_size_matches_reference_golden():
    """160 x 160 input: position embeddings resampled as dino_vits.py:213-233 does -- oracle forward, oracle helper and
    the host-side parameter preparation of the product (dcr_b200.nets.interpolate_pos_embed) against the reference's own
    outputs (tests/golden/make_golden.py: make_dino_other_size)."""
    from dcr_b200 import nets
    g = np.load(os.path.join(GOLD, "dino_vits16_seed0_160.npz"))
    sd = om.make_vit_state_dict(0)
    x = _golden_inputs(0, size=160)
    assert abs(float(x.double().sum()) - float(g["in_checksum"])) < 1e-6
    y = om.vit_forward(sd, x).numpy()
    np.testing.assert_allclose(y, g["out"], rtol=0, atol=2e-5)
    pos_o = om.interpolate_pos_encoding(sd["pos_embed"], 100, 160, 160, 16).numpy()
    pos_p = nets.interpolate_pos_embed(sd["pos_embed"], 10, 10).numpy()
    assert pos_o.shape == g["pos_embed"].shape == pos_p.shape == (1, 101, 384)
    np.testing.assert_array_equal(pos_o, g["pos_embed"])
    np.testing.assert_array_equal(pos_p, g["pos_embed"])
    # same size: untouched
    assert nets.interpolate_pos_embed(sd["pos_embed"], 14, 14) is sd["pos_embed"]


def test_vit_oracle_variants_match_reference_golden():
    """global_pool='' (all normed tokens; diff_retrieval.py:258-263 + dino_vits.py:255-256) and
    get_intermediate_layers(x, 3)[0] (--layer 3; utils_ret.py:732,745) against the reference module's own outputs."""
    g = np.load(os.path.join(GOLD, "dino_vits16_seed0_variants.npz"))
    x = _golden_inputs(0)
    assert abs(float(x.double().sum()) - float(g["in_checksum"])) < 1e-6
    sd = om.make_vit_state_dict(0)
    tok = om.vit_forward(sd, x, global_pool="").reshape(2, 197, 384)
    np.testing.assert_allclose(tok[:, ::14, :].numpy(), g["tokens_rows"], rtol=0, atol=3e-5)
    np.testing.assert_allclose(tok.double().sum(dim=-1).numpy(), g["tokens_sum"], rtol=0, atol=2e-3)
    cls3 = om.vit_forward(sd, x, n_last_layers=3)
    np.testing.assert_allclose(cls3.numpy(), g["layer3_cls"], rtol=0, atol=3e-5)
    tok3 = om.vit_forward(sd, x, n_last_layers=3, global_pool="").reshape(2, 197, 384)
    np.testing.assert_allclose(tok3.double().sum(dim=-1).numpy(), g["layer3_tokens_sum"], rtol=0, atol=2e-3)

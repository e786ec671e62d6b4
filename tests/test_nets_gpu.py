"""GPU parity of the descriptor networks (through dcr_net_* C ABI) against the CPU oracle."""
import numpy as np
import pytest
import torch

from dcr_b200 import nets, synthetic
from oracle import models as om

pytestmark = pytest.mark.gpu


def _imgs(n, seed):
    return synthetic.images(n, seed=seed)


def _report(name, got, ref):
    err = (got - ref).abs().max().item()
    cos = torch.nn.functional.cosine_similarity(got, ref, dim=1).min().item()
    print(f"{name}: max_abs_err={err:.3e} min_cos={cos:.7f}")
    return err, cos


@pytest.mark.parametrize("mean,std", [((0.5, 0.5, 0.5), (0.5, 0.5, 0.5)),
                                      ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))])
def test_sscd_resnet50_parity_mode(mean, std):
    sd = om.make_sscd_state_dict(0)
    img = _imgs(5, 1)
    ref = om.sscd_forward(sd, om.preprocess(img, mean, std))
    net = nets.build_sscd_resnet50(sd, max_batch=4, precision="parity", mean=mean, std=std)   # 5 images: 4 + 1 tail
    got = net(img.cuda()).cpu()
    err, cos = _report("sscd parity", got, ref)
    assert err < 1e-4, err
    # scores (dot products of unit descriptors) within the 1e-4 tolerance of BASELINE.json north_star
    assert (got @ got.T - ref @ ref.T).abs().max().item() < 1e-4


def test_sscd_resnet50_exact_mode():
    sd = om.make_sscd_state_dict(0)
    img = _imgs(3, 7)
    ref = om.sscd_forward(sd, om.preprocess(img))
    net = nets.build_sscd_resnet50(sd, max_batch=2, precision="exact")
    got = net(img.cuda()).cpu()
    err, cos = _report("sscd exact", got, ref)
    assert err < 1e-5, err      # unit-norm descriptors: a few fp32 ulps of summation-order noise in the oracle itself


def test_dino_vits16_exact_mode():
    sd = om.make_vit_state_dict(0)
    img = _imgs(3, 3)
    ref = om.vit_forward(sd, om.preprocess(img))
    net = nets.build_dino_vit(sd, max_batch=2, precision="exact")
    got = net(img.cuda()).cpu()
    err, cos = _report("vit exact", got, ref)
    assert err < 3e-5 * max(1.0, ref.abs().max().item()), err


def test_sscd_resnet50_fast_mode():
    sd = om.make_sscd_state_dict(0)
    img = _imgs(6, 2)
    x = om.preprocess(img)
    ref32 = om.sscd_forward(sd, x)
    refq = om.sscd_forward(sd, x, bf16_points=True)
    net = nets.build_sscd_resnet50(sd, max_batch=8, precision="fast")
    got = net(img.cuda()).cpu()
    err_q, _ = _report("sscd fast vs bf16-point oracle", got, refq)
    err_32, cos = _report("sscd fast vs fp32 oracle", got, ref32)
    assert err_q < 1.5e-2 and cos > 0.995


def test_dino_vits16_parity_mode():
    sd = om.make_vit_state_dict(0)
    img = _imgs(3, 3)
    ref = om.vit_forward(sd, om.preprocess(img))
    net = nets.build_dino_vit(sd, max_batch=2, precision="parity")
    got = net(img.cuda()).cpu()
    err, cos = _report("vit parity", got, ref)
    assert err < 2e-4 * max(1.0, ref.abs().max().item()), err
    gn, rn = torch.nn.functional.normalize(got, dim=1), torch.nn.functional.normalize(ref, dim=1)
    assert (gn @ gn.T - rn @ rn.T).abs().max().item() < 1e-4


def test_dino_vits16_fast_mode():
    sd = om.make_vit_state_dict(1)
    img = _imgs(4, 4)
    x = om.preprocess(img)
    ref32 = om.vit_forward(sd, x)
    refq = om.vit_forward(sd, x, bf16_points=True)
    net = nets.build_dino_vit(sd, max_batch=4, precision="fast")
    got = net(img.cuda()).cpu()
    err_q, _ = _report("vit fast vs bf16-point oracle", got, refq)
    err_32, cos = _report("vit fast vs fp32 oracle", got, ref32)
    assert err_q < 6e-2 * max(1.0, refq.abs().max().item()) and cos > 0.99


def test_forward_rejects_cpu_and_wrong_size():
    from dcr_b200._lib import DcrError
    net = nets.build_dino_vit(om.make_vit_state_dict(0, depth=1), max_batch=2)
    with pytest.raises(DcrError):
        net(_imgs(1, 0))                                  # CPU tensor
    with pytest.raises(DcrError):
        net(torch.zeros(1, 224, 224, 3, dtype=torch.uint8, device="cuda"))


def test_maxpool_fast_path_is_bit_identical_to_generic(monkeypatch):
    """The packed-bf16 3x3 max pool (two outputs per thread, clamped loads) against the generic pool kernel: maxima of
    bf16 values are exact, so whole-network outputs must not differ in a single bit (ResNet stem pool: stride 2 pad 1;
    Inception pools: stride 2 pad 0, odd widths)."""
    img = _imgs(3, 9).cuda()
    net = nets.build_sscd_resnet50(om.make_sscd_state_dict(2), max_batch=4, precision="fast")
    fast = net(img).clone()
    monkeypatch.setenv("DCR_B200_TUNING", "1")
    monkeypatch.setenv("DCR_POOL_GENERIC", "1")
    generic = net(img).clone()
    monkeypatch.delenv("DCR_POOL_GENERIC")
    assert torch.equal(fast, generic)
    img2 = torch.randint(0, 256, (2, 299, 299, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(3)).cuda()
    inc = nets.build_fid_inception(om.make_inception_state_dict(1), max_batch=2, precision="fast")
    fast = inc(img2).clone()
    monkeypatch.setenv("DCR_B200_TUNING", "1")
    monkeypatch.setenv("DCR_POOL_GENERIC", "1")
    generic = inc(img2).clone()
    assert torch.equal(fast, generic)


def test_sscd_grouped_trunk_1024d():
    """ResNeXt-style trunk with a 1024-d head (the sscd_disc_large family, `--arch resnet50_disc`): grouped 3x3 convs as
    dense block-diagonal GEMMs; descriptors wider than 512 then go through the streamed-query similarity kernel."""
    from dcr_b200 import similarity
    from oracle import similarity as osim
    sd = om.make_sscd_state_dict(4, dims=1024, arch="resnext_tiny")
    img = _imgs(6, 12)
    ref = om.sscd_forward(sd, om.preprocess(img))
    net = nets.build_sscd_resnet50(sd, max_batch=4, precision="exact")
    got = net(img.cuda())
    err, cos = _report("sscd resnext exact", got.cpu(), ref)
    assert got.shape == (6, 1024) and err < 1e-5
    fast = nets.build_sscd_resnet50(sd, max_batch=8, precision="fast")(img.cuda()).cpu()
    _, cosf = _report("sscd resnext fast", fast, ref)
    assert cosf > 0.995
    v, i = similarity.sim_topk(got[:2].contiguous(), got.contiguous(), 3)
    ov, oi = osim.sim_topk(got[:2].cpu().numpy(), got.cpu().numpy(), 3)
    assert np.array_equal(i.cpu().numpy(), oi)


def test_dino_vit_base_width():
    """vit_base/16 geometry (768-d, 12 heads; `--arch vit_base`, dino_vits.py:366-378) at reduced depth: heads and
    patch size are inferred from the state_dict; the 768-d descriptors exceed the resident-query tile of the similarity
    kernel (streamed mode)."""
    sd = om.make_vit_state_dict(2, dim=768, depth=2, heads=12)
    img = _imgs(3, 5)
    ref = om.vit_forward(sd, om.preprocess(img), heads=12)
    net = nets.build_dino_vit(sd, max_batch=4, precision="exact")
    got = net(img.cuda()).cpu()
    err, cos = _report("vit-b exact", got, ref)
    assert got.shape == (3, 768) and err < 3e-5 * max(1.0, ref.abs().max().item())
    fast = nets.build_dino_vit(sd, max_batch=4, precision="fast")(img.cuda()).cpu()
    _, cosf = _report("vit-b fast", fast, ref)
    assert cosf > 0.99


def test_sscd_multiscale():
    """--multiscale (utils_ret.py:676-698): three bilinearly rescaled forward passes (224, 158, 112), averaged."""
    from dcr_b200 import retrieval
    sd = om.make_sscd_state_dict(6)
    img = _imgs(3, 21)
    x = om.preprocess(img)
    ref = om.sscd_forward_multiscale(sd, x)
    nets_ms = [nets.build_sscd_resnet50(sd, max_batch=4, precision="exact", scale_factor=s) for s in retrieval.MULTI_SCALES]
    # each scale on its own against the oracle's F.interpolate + forward
    import torch.nn.functional as F
    for s, net in zip(retrieval.MULTI_SCALES, nets_ms):
        inp = x if s == 1 else F.interpolate(x, scale_factor=s, mode="bilinear", align_corners=False)
        err, _ = _report(f"sscd scale {s:.3f} ({inp.shape[-1]} px)", net(img.cuda()).cpu(), om.sscd_forward(sd, inp))
        assert err < 2e-5, (s, err)
    got = retrieval.extract_features_multiscale(nets_ms, img.cuda()).cpu()
    err, cos = _report("sscd multiscale", got, ref)
    assert err < 2e-5


def test_fused_stem_variants_agree():
    """stem_fused.cu: the overlapping-window (Toeplitz) stem against the space-to-depth stem of conv_gemm.cu -- same bf16
    inputs, another summation order (fp32 noise only) -- and its pooled variant, which must be bit-identical to the
    unpooled one (a maximum of bf16 values is exact).  Also at the multi-scale input sizes (odd 79 x 79 stem output)."""
    sd = om.make_sscd_state_dict(0)
    img = _imgs(5, 3)
    ref = om.sscd_forward(sd, om.preprocess(img), bf16_points=True)
    out = {st: nets.build_sscd_resnet50(sd, max_batch=4, precision="fast", stem=st)(img.cuda()).cpu()
           for st in ("s2d", "toeplitz", "toeplitz_pool")}
    assert torch.equal(out["toeplitz"], out["toeplitz_pool"])
    assert (out["toeplitz"] - out["s2d"]).abs().max().item() < 2e-3
    assert (out["toeplitz_pool"] - ref).abs().max().item() < 1.5e-2
    for sf in (0.5, 1 / 2 ** 0.5):
        a = nets.build_sscd_resnet50(sd, max_batch=8, precision="fast", stem="s2d", scale_factor=sf)(img.cuda())
        b = nets.build_sscd_resnet50(sd, max_batch=8, precision="fast", stem="toeplitz", scale_factor=sf)(img.cuda())
        c = nets.build_sscd_resnet50(sd, max_batch=8, precision="fast", stem="toeplitz_pool", scale_factor=sf)(img.cuda())
        assert torch.equal(b, c) and (a - b).abs().max().item() < 2e-3
    assert nets.build_sscd_resnet50(sd, max_batch=2, precision="parity").meta[0][0] == nets.OP_STEM_S2D   # other modes keep s2d


def test_block_fusion_is_bit_identical(monkeypatch):
    """bottleneck_fuse.cu: conv3 + residual + ReLU fused with the next block's conv1 (and the expansion-only variant)
    against the separate launches of conv_gemm.cu: same K order and epilogue arithmetic -> the same bits."""
    sd = om.make_sscd_state_dict(1)
    img = _imgs(6, 5).cuda()
    net = nets.build_sscd_resnet50(sd, max_batch=4, precision="fast")
    from dcr_b200 import similarity
    l0 = similarity.kernel_launch_count()
    fused = net(img).clone()
    n_fused = similarity.kernel_launch_count() - l0
    monkeypatch.setenv("DCR_B200_TUNING", "1")
    monkeypatch.setenv("DCR_NO_BLOCK_FUSION", "1")
    l0 = similarity.kernel_launch_count()
    plain = net(img).clone()
    n_plain = similarity.kernel_launch_count() - l0
    assert torch.equal(fused, plain)
    assert n_plain - n_fused == 2 * 6          # two forward chunks (4 + 2 images), six fused pairs each

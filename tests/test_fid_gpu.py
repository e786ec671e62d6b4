"""GPU: FID Inception forward and streaming statistics vs the oracle."""
import numpy as np
import pytest
import torch

from dcr_b200 import fid as dfid
from dcr_b200 import nets
from oracle import fid as ofid
from oracle import models as om

pytestmark = pytest.mark.gpu


def _imgs(n, seed):
    return torch.randint(0, 256, (n, 299, 299, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(seed))


def test_inception_parity_mode():
    sd = om.make_inception_state_dict(0)
    img = _imgs(3, 1)
    ref = om.fid_inception_forward(sd, om.fid_preprocess(img))
    net = nets.build_fid_inception(sd, max_batch=2, precision="parity")
    got = net(img.cuda()).cpu()
    err = (got - ref).abs().max().item()
    print(f"inception parity: max_abs_err={err:.3e} ref_max={ref.abs().max().item():.3f}")
    # split-bf16 on tensor cores: the truncating fp32 accumulator leaves ~4e-6 relative per conv, ~3e-4 after 94 layers
    assert err < 5e-4 * max(1.0, ref.abs().max().item())


def test_inception_exact_mode_matches_reference_golden():
    """Exact mode (float64 accumulation) against the activations the REFERENCE's own metrics/inception.py produced for
    the same seeded weights and images (tests/golden/make_golden.py) -- and against the oracle."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "fid_inception_seed0.npz"))
    sd = om.make_inception_state_dict(0)
    img = torch.randint(0, 256, (2, 299, 299, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(8000))
    net = nets.build_fid_inception(sd, max_batch=2, precision="exact")
    got = net(img.cuda()).cpu().numpy()
    gold = g["out"]
    err = np.abs(got - gold).max()
    ref = om.fid_inception_forward(sd, om.fid_preprocess(img)).numpy()
    err_o = np.abs(got - ref).max()
    print(f"inception exact: vs reference golden {err:.3e}, vs oracle {err_o:.3e}, max|act|={np.abs(gold).max():.3f}")
    # fp32 evaluations of a 94-convolution network differ by their own summation order (the oracle is within 2e-5 of the
    # golden vector, tests/test_oracle_fid.py); the exact path must sit inside the same band
    assert err < 3e-5 * max(1.0, np.abs(gold).max()) and err_o < 3e-5 * max(1.0, np.abs(ref).max())


def test_fid_value_exact_mode():
    """FID of two small image sets, exact mode vs the oracle pipeline: the north-star tolerance (1e-4) on the value."""
    sd = om.make_inception_state_dict(2)
    net = nets.build_fid_inception(sd, max_batch=8, precision="exact")
    real, gen = _imgs(12, 20), _imgs(12, 21)
    got = dfid.fid_from_images(net, real, gen, batch_size=8)
    a = om.fid_inception_forward(sd, om.fid_preprocess(real)).numpy()
    b = om.fid_inception_forward(sd, om.fid_preprocess(gen)).numpy()
    ref = ofid.frechet_distance(*ofid.activation_statistics(a), *ofid.activation_statistics(b))
    print(f"fid exact: got {got:.6f} oracle {ref:.6f}")
    assert abs(got - ref) < 1e-4 * max(1.0, abs(ref)), (got, ref)


def test_inception_fast_mode():
    sd = om.make_inception_state_dict(0)
    img = _imgs(3, 2)
    x = om.fid_preprocess(img)
    ref32 = om.fid_inception_forward(sd, x)
    refq = om.fid_inception_forward(sd, x, bf16_points=True)
    net = nets.build_fid_inception(sd, max_batch=4, precision="fast")
    got = net(img.cuda()).cpu()
    eq = (got - refq).abs().max().item()
    e32 = (got - ref32).abs().max().item()
    print(f"inception fast: vs bf16-point oracle {eq:.3e}, vs fp32 oracle {e32:.3e}, ref_max={ref32.abs().max().item():.3f}")
    assert eq < 3e-2 * max(1.0, refq.abs().max().item())


def test_streaming_statistics_match_numpy():
    rng = torch.Generator().manual_seed(5)
    x = torch.randn(1000, 256, generator=rng) * 0.7 + 3.0          # large mean: exercises the shift
    st = dfid.ActivationStatistics(256)
    for s in range(0, 1000, 170):                                  # ragged batches
        st.update(x[s:s + 170].cuda())
    mu, sig = st.finalize()
    rmu, rsig = ofid.activation_statistics(x.numpy())
    np.testing.assert_allclose(mu, rmu, rtol=0, atol=1e-12)
    np.testing.assert_allclose(sig, rsig, rtol=0, atol=1e-11)
    # Frechet distance of the two halves, GPU eigh formulation vs the sqrtm oracle
    a, b = x[:500].numpy(), (x[500:] * 1.1 + 0.05).numpy()
    ref = ofid.frechet_distance(*ofid.activation_statistics(a), *ofid.activation_statistics(b))
    sa, sb = dfid.ActivationStatistics(256), dfid.ActivationStatistics(256)
    sa.update(torch.from_numpy(a).cuda())
    sb.update(torch.from_numpy(b).cuda())
    got = dfid.frechet_distance(*sa.finalize(), *sb.finalize())
    assert abs(got - ref) < 1e-4, (got, ref)


def test_fid_pipeline_small():
    """End to end on 2 x 24 images (d = 2048 > N: covariances singular, as in the reference's small-N use): the
    activation statistics must match the oracle's; the FID value is compared through them."""
    sd = om.make_inception_state_dict(1)
    net = nets.build_fid_inception(sd, max_batch=8, precision="parity")
    real, gen = _imgs(24, 10), _imgs(24, 11)
    m1, s1 = dfid.statistics_of_images(net, real, batch_size=8)
    act = om.fid_inception_forward(sd, om.fid_preprocess(real)).numpy()
    rm, rs = ofid.activation_statistics(act)
    tol = 1e-3 * np.abs(act).max()      # parity mode is ~3e-4 relative on Inception (see test_inception_parity_mode)
    assert np.abs(m1 - rm).max() < tol and np.abs(s1 - rs).max() < tol * np.abs(act).max()
    v = dfid.fid_from_images(net, real, gen, batch_size=8)
    assert np.isfinite(v) and v >= -1e-6


def test_calculate_fid_given_paths_mirror(tmp_path, capsys):
    """metrics/fid.py:224-275 call surface: folders of images, .npz statistics files, error behaviour."""
    from PIL import Image
    sd = om.make_inception_state_dict(3)
    a, b = tmp_path / "real" / "sub", tmp_path / "gen"
    a.mkdir(parents=True)
    b.mkdir()
    ia, ib = _imgs(10, 30), _imgs(9, 31)
    for i in range(10):
        Image.fromarray(ia[i].numpy()).save(a / f"{i}.png")
    for i in range(9):
        Image.fromarray(ib[i].numpy()).save(b / f"{i}.png")
    net = nets.build_fid_inception(sd, max_batch=8, precision="fast")
    v = dfid.calculate_fid_given_paths([str(tmp_path / "real"), str(b)], 8, "cuda", 2048, weights=net)
    assert capsys.readouterr().out.strip().splitlines()[0] == "2048"              # fid.py:244
    assert abs(v - dfid.fid_from_images(net, dfid.load_resized(str(tmp_path / "real")), dfid.load_resized(str(b)), 8)) < 1e-6
    npz = str(tmp_path / "real_stats.npz")
    dfid.save_fid_stats([str(tmp_path / "real"), npz], 8, "cuda", 2048, weights=net)
    v2 = dfid.calculate_fid_given_paths([npz, str(b)], 8, "cuda", 2048, weights=net)
    assert abs(v - v2) < 1e-6
    with pytest.raises(RuntimeError, match="Invalid path"):
        dfid.calculate_fid_given_paths([str(tmp_path / "missing"), str(b)], 8, "cuda", 2048, weights=net)
    with pytest.raises(FileNotFoundError):
        dfid.calculate_fid_given_paths([str(tmp_path / "real"), str(b)], 8, "cuda", 2048, weights=str(tmp_path / "none.pth"))


@pytest.mark.parametrize("dims,stop", [(64, "pool1"), (192, "pool2"), (768, "Mixed_6e")])
def test_fid_feature_blocks_by_dims(dims, stop, tmp_path):
    """InceptionV3.BLOCK_INDEX_BY_DIM (metrics/inception.py:22-28): dims 64 / 192 / 768 select earlier blocks, average
    pooled to 1x1 (metrics/fid.py:130-133); calculate_fid_given_paths(dims=...) builds the matching network."""
    from PIL import Image
    sd = om.make_inception_state_dict(0)
    net = dfid._build_inception(sd, 4, "exact", dims)
    assert net.out_dim == dims
    img = _imgs(3, 30 + dims)
    ref = om.fid_inception_forward(sd, om.fid_preprocess(img), stop_after=stop)
    got = net(img.cuda()).cpu()
    assert (got - ref).abs().max().item() < 3e-5 * max(1.0, ref.abs().max().item())
    if dims == 64:      # the path-level entry with a non-default dims, on PNG folders
        for name, seed in (("real", 1), ("gen", 2)):
            d = tmp_path / name
            d.mkdir()
            for i, im in enumerate(_imgs(70, seed)):
                Image.fromarray(im.numpy()).save(d / f"{i}.png")
        val = dfid.calculate_fid_given_paths([str(tmp_path / "real"), str(tmp_path / "gen")], 50, "cuda", 64, 1, weights=sd,
                                             precision="exact")
        acts = [om.fid_inception_forward(sd, om.fid_preprocess(dfid.load_resized(str(tmp_path / n))), stop_after="pool1").numpy()
                for n in ("real", "gen")]
        want = ofid.frechet_distance(*ofid.activation_statistics(acts[0]), *ofid.activation_statistics(acts[1]))
        assert abs(val - want) < 1e-4 * max(1.0, abs(want)), (val, want)
    with pytest.raises(KeyError):
        dfid._build_inception(sd, 4, "fast", 100)


def test_frechet_raises_on_imaginary_component():
    """metrics/fid.py:187-191: a covariance product with a clearly negative eigenvalue has no real square root."""
    rng = np.random.default_rng(0)
    a = rng.standard_normal((8, 8))
    s1 = a @ a.T + np.eye(8)
    s2 = -np.eye(8)                                    # not a covariance: sqrtm(s1 s2) is purely imaginary
    with pytest.raises(ValueError, match="Imaginary component"):
        dfid.frechet_distance(np.zeros(8), s1, np.zeros(8), s2)
    assert dfid.frechet_distance(np.zeros(8), s1, np.zeros(8), s1) < 1e-8

"""GPU parity: dcr_sim_topk (through the C ABI) vs the CPU oracle -- indices bit-exact, scores to fp32 rounding."""
import numpy as np
import pytest
import torch

from dcr_b200 import similarity, synthetic
from oracle import similarity as osim

pytestmark = pytest.mark.gpu


def _run(q, g, k, **kw):
    v, i = similarity.sim_topk(q.cuda(), g.cuda(), k, **kw)
    torch.cuda.synchronize()
    return v.cpu().numpy(), i.cpu().numpy()


def _check(q, g, k, rows=None):
    v, i = _run(q, g, k)
    qn = q.numpy() if rows is None else q.numpy()[rows]
    ov, oi = osim.sim_topk(qn, g.numpy(), k)
    if rows is not None:
        v, i = v[rows], i[rows]
    bad = np.nonzero((i != oi).any(axis=1))[0]
    assert bad.size == 0, f"{bad.size} query rows differ, first {bad[:5]}: got {i[bad[:3]]} want {oi[bad[:3]]}"
    np.testing.assert_allclose(v, ov, rtol=0, atol=1.2e-7)   # fp64 dot rounded to fp32 on both sides
    return similarity.sim_topk_stats()


@pytest.mark.parametrize("nq,ng,d,k", [
    (256, 1000, 512, 1),      # BASELINE config 1 (similarity part)
    (256, 1000, 512, 10),
    (1, 16, 64, 1),           # tiny
    (7, 300, 100, 3),         # d not a multiple of 64, ragged everything
    (130, 257, 384, 2),       # DINO dim, just over one tile in both directions
    (513, 4097, 512, 5),
    (300, 20000, 512, 10),    # several gallery tiles per unit
    (2000, 3000, 256, 16),    # many q-tiles, max k
    (300, 5000, 1024, 10),    # SSCD 'large' descriptor dim: query tile streamed instead of resident
    (260, 3000, 768, 1),      # ViT-B dim
    (40, 700, 2048, 3),       # Inception pool3 dim
])
def test_parity_synthetic(nq, ng, d, k):
    q, g = synthetic.descriptors(nq, ng, d, seed=nq + ng)
    _check(q, g, k)


def test_parity_unnormalised_negative():
    gen = torch.Generator().manual_seed(5)
    q = torch.randn(200, 128, generator=gen) * 3
    g = -torch.rand(5000, 128, generator=gen) * q[:1].abs().mean()    # mostly negative scores for positive q
    _check(q.abs(), g, 4)


def test_duplicates_tie_rule():
    q, g = synthetic.descriptors(128, 2048, 512, seed=9)
    g[100:140] = g[7]                     # 41 identical gallery rows
    g[1500] = q[3]
    g[300] = q[3]                         # exact duplicate pair as best match of query 3
    st = _check(q, g, 10)
    v, i = _run(q, g, 10)
    assert i[3, 0] == 300 and i[3, 1] == 1500


def test_identical_gallery_wide_descriptors():
    """brute-force path with d > 512 (adaptive batch)"""
    g = torch.nn.functional.normalize(torch.ones(1, 1024), dim=1).repeat(300, 1)
    q, _ = synthetic.descriptors(40, 16, 1024, seed=3)
    st = _check(q, g, 5)
    assert st["n_flagged"] == 40


def test_all_identical_gallery_forces_exact_fallback():
    q, _ = synthetic.descriptors(64, 8, 64, seed=11)
    g = q[:1].repeat(3000, 1).contiguous()
    v, i = _run(q, g, 5)
    assert (i == np.arange(5)[None, :]).all()
    assert similarity.sim_topk_stats()["n_flagged"] == 64   # certificate cannot hold: everything ties


def test_concentrated_descriptors():
    """Descriptors sharing a common component (un-whitened networks).  Gallery centring keeps the bf16 error bound
    proportional to the spread of the gallery; results must stay exact either way, and for a moderate common
    component almost no query may need the brute-force path."""
    gen = torch.Generator().manual_seed(21)
    common = torch.randn(1, 512, generator=gen)
    q = torch.nn.functional.normalize(common + 1.0 * torch.randn(600, 512, generator=gen), dim=1)
    g = torch.nn.functional.normalize(common + 1.0 * torch.randn(30000, 512, generator=gen), dim=1)
    st = _check(q, g, 10)
    assert st["n_flagged"] <= 6, st
    # extreme case (all cosines > 0.99): with query AND gallery centring the tensor cores only see the small
    # centred parts, so even this stays on the fast path
    q2 = torch.nn.functional.normalize(common + 0.05 * torch.randn(600, 512, generator=gen), dim=1)
    g2 = torch.nn.functional.normalize(common + 0.05 * torch.randn(30000, 512, generator=gen), dim=1)
    st = _check(q2, g2, 10)
    assert st["n_flagged"] <= 6, st


def test_k_equals_gallery_size():
    q, g = synthetic.descriptors(20, 16, 64, seed=12)
    _check(q, g, 16)


def test_index_base_and_stride():
    q, g = synthetic.descriptors(50, 700, 128, seed=13)
    v0, i0 = _run(q, g, 3)
    v1, i1 = _run(q, g, 3, index_base=5, index_stride=4)
    assert np.array_equal(i1, 5 + 4 * i0) and np.array_equal(v0, v1)


def test_background_top2_is_self_then_neighbour():
    _, g = synthetic.descriptors(8, 3000, 512, seed=14)
    v, i = _run(g, g, 2)
    assert (i[:, 0] == np.arange(3000)).all()
    np.testing.assert_allclose(v[:, 1], osim.background_second_best(g.numpy()), atol=1.2e-7)


def test_linearity_property_full_size():
    """BASELINE config 2 shape (10k x 100k x 512): exact check on a query subsample + permutation invariance."""
    q, g = synthetic.descriptors(10000, 100000, 512, seed=2)
    rows = np.random.default_rng(0).choice(10000, 384, replace=False)
    st = _check(q, g, 10, rows=np.sort(rows))
    assert st["n_flagged"] < 100
    # permuting the gallery permutes the indices and nothing else
    perm = torch.randperm(100000, generator=torch.Generator().manual_seed(1))
    v0, i0 = _run(q[:1024], g, 1)
    v1, i1 = _run(q[:1024], g[perm].contiguous(), 1)
    assert np.array_equal(perm.numpy()[i1[:, 0]], i0[:, 0]) and np.array_equal(v0, v1)


def test_merge_matches_oracle():
    q, g = synthetic.descriptors(300, 4000, 128, seed=15)
    parts_v, parts_i = [], []
    for s in range(4):
        v, i = similarity.sim_topk(q.cuda(), g[s::4].contiguous().cuda(), 10, index_base=s, index_stride=4)
        parts_v.append(v)
        parts_i.append(i)
    mv, mi = similarity.topk_merge(torch.stack(parts_v), torch.stack(parts_i), 10)
    ov, oi = osim.sim_topk(q.numpy(), g.numpy(), 10)
    assert np.array_equal(mi.cpu().numpy(), oi)
    np.testing.assert_allclose(mv.cpu().numpy(), ov, atol=1.2e-7)


def test_l2_normalize():
    x = torch.randn(1000, 512, generator=torch.Generator().manual_seed(3))
    x[5] = 0
    ref = torch.nn.functional.normalize(x, dim=1, p=2)
    got = similarity.l2_normalize_(x.clone().cuda()).cpu()
    np.testing.assert_allclose(got.numpy(), ref.numpy(), rtol=5e-7, atol=1e-12)


def test_rejects_bad_arguments():
    from dcr_b200._lib import DcrError
    q, g = synthetic.descriptors(4, 64, 64)
    with pytest.raises(DcrError):
        similarity.sim_topk(q, g.cuda(), 1)          # CPU tensor: no CPU path
    with pytest.raises(DcrError):
        similarity.sim_topk(q.cuda(), g.cuda(), 17)  # k > 16
    with pytest.raises(DcrError):
        similarity.sim_topk(q.cuda(), g[:3].cuda(), 5)  # k > G


@pytest.mark.parametrize("nq,ng,d,c,k", [(64, 3000, 512, 4, 10), (33, 1000, 96, 3, 1), (20, 500, 512, 32, 5)])
def test_splitloss_topk_matches_oracle(nq, ng, d, c, k):
    """diff_retrieval.py:393-400: score = max over the c descriptor parts of the per-part dot products."""
    from oracle import similarity as osim
    q, g = synthetic.descriptors(nq, ng, d, seed=11 + c)
    g[5] = g[9]                                   # exact duplicate rows: tie -> lowest index first
    v, i = similarity.sim_topk_split(q.cuda(), g.cuda(), k, c)
    ov, oi = osim.sim_topk_split(q.numpy(), g.numpy(), k, c)
    assert np.array_equal(i.cpu().numpy(), oi)
    np.testing.assert_allclose(v.cpu().numpy(), ov, rtol=0, atol=1e-6)
    with pytest.raises(similarity._lib.DcrError):
        similarity.sim_topk_split(q.cuda(), g.cuda(), k, 7)       # 7 does not divide d


@pytest.mark.parametrize("nq,ng,d,c,k", [(50, 2000, 512, 8, 1), (21, 600, 128, 4, 3), (9, 300, 96, 2, 8)])
def test_splitloss_cross_matches_oracle(nq, ng, d, c, k):
    """--stype cross (einsum_in_chunks, diff_retrieval.py:643-662): score = max over every (gallery part, query part)."""
    q, g = synthetic.descriptors(nq, ng, d, seed=23 + c)
    g[7] = g[2]
    v, i = similarity.sim_topk_split(q.cuda(), g.cuda(), k, c, cross=True)
    ov, oi = osim.sim_topk_split(q.numpy(), g.numpy(), k, c, cross=True)
    assert np.array_equal(i.cpu().numpy(), oi)
    np.testing.assert_allclose(v.cpu().numpy(), ov, rtol=0, atol=1e-6)
    with pytest.raises(similarity._lib.DcrError):
        similarity.sim_topk_split(q.cuda(), g.cuda(), k, 7, cross=True)        # 7 does not divide d


@pytest.mark.parametrize("nq,ng,d,k", [(1000, 20000, 512, 10), (300, 5000, 384, 1), (129, 3000, 100, 2), (700, 9000, 516, 10)])
def test_rescore_warp_and_block_forms_agree(nq, ng, d, k, monkeypatch):
    """The one-warp-per-query re-score kernel and the one-block-per-query form it replaces for small slot counts: same
    candidates, same fp64 association, same selection -> identical scores (bitwise) and indices, and both equal the oracle."""
    q, g = synthetic.descriptors(nq, ng, d, seed=31)
    g[7] = g[3]                                                   # an exact tie between two gallery rows
    v_w, i_w = _run(q, g, k)
    monkeypatch.setenv("DCR_B200_TUNING", "1")
    monkeypatch.setenv("DCR_SIM_RESCORE_BLOCK", "1")
    v_b, i_b = _run(q, g, k)
    assert np.array_equal(i_w, i_b) and np.array_equal(v_w, v_b)
    monkeypatch.delenv("DCR_SIM_RESCORE_BLOCK")
    _check(q, g, k)

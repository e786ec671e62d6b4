"""Oracle self-checks (CPU): the restated similarity/top-k agrees with the reference's literal inline ops."""
import numpy as np
import pytest
import torch

from dcr_b200 import synthetic
from oracle import similarity as osim


def test_normalize_matches_torch():
    x = torch.randn(37, 96, generator=torch.Generator().manual_seed(0))
    x[3] = 0
    ref = torch.nn.functional.normalize(x, dim=1, p=2).numpy()
    got = osim.l2_normalize(x.numpy())
    np.testing.assert_allclose(got, ref, rtol=5e-7, atol=1e-12)  # 1-2 ulp: reduction order of the norm


@pytest.mark.parametrize("nq,ng,d,k", [(64, 500, 128, 1), (33, 1000, 512, 10), (256, 1000, 512, 1)])
def test_topk_matches_literal_reference(nq, ng, d, k):
    q, g = synthetic.descriptors(nq, ng, d, seed=1)
    v, i = osim.sim_topk(q.numpy(), g.numpy(), k)
    rv, ri = osim.sim_topk_reference_fp32(q.numpy(), g.numpy(), k)
    np.testing.assert_allclose(v, rv, atol=2e-6)
    # indices agree wherever the reference is well defined (gap to the next score above fp32 noise)
    S = q.numpy().astype(np.float64) @ g.numpy().astype(np.float64).T
    for r in range(nq):
        if not np.array_equal(i[r], ri[r]):
            srt = np.sort(S[r])[::-1][:k + 1]
            assert np.min(np.abs(np.diff(srt))) < 1e-6, f"row {r}: {i[r]} vs {ri[r]}"


def test_tie_rule_lowest_index_first():
    g = np.zeros((10, 8), dtype=np.float32)
    g[:, 0] = 1.0            # all identical
    q = np.zeros((2, 8), dtype=np.float32)
    q[:, 0] = 1.0
    v, i = osim.sim_topk(q, g, 3)
    assert i.tolist() == [[0, 1, 2], [0, 1, 2]]
    g[7, 0] = 2.0
    v, i = osim.sim_topk(q, g, 3)
    assert i.tolist() == [[7, 0, 1], [7, 0, 1]]


def test_merge_equals_concat():
    q, g = synthetic.descriptors(40, 900, 64, seed=3)
    full_v, full_i = osim.sim_topk(q.numpy(), g.numpy(), 5)
    parts_v, parts_i = [], []
    for s in range(3):
        v, i = osim.sim_topk(q.numpy(), g.numpy()[s::3], 5)
        parts_v.append(v)
        parts_i.append(s + 3 * i)
    mv, mi = osim.merge_topk(np.stack(parts_v), np.stack(parts_i), 5)
    assert np.array_equal(mi, full_i)
    np.testing.assert_array_equal(mv, full_v)


def test_background_second_best_is_topk2():
    _, g = synthetic.descriptors(8, 300, 64, seed=4)
    bg = osim.background_second_best(g.numpy())
    sim2 = torch.mm(g, g.T)
    ref = sim2.T.topk(2, dim=1)[0][:, -1].numpy()
    np.testing.assert_allclose(bg, ref, atol=2e-6)


def test_stats_keys():
    st = osim.retrieval_stats(np.linspace(0, 1, 101), np.linspace(0, 0.5, 50))
    assert set(st) == {"sim_mean", "sim_std", "sim_75pc", "sim_90pc", "sim_95pc", "sim_gt_05pc", "bg_mean", "bg_std",
                       "bg_75pc", "bg_90pc", "bg_95pc"}
    assert abs(st["sim_gt_05pc"] - 50 / 101) < 1e-12


def test_splitloss_oracle_is_the_einsum_max():
    """Literal restatement of diff_retrieval.py:396-400 with torch/einops-free numpy on a tiny case."""
    rng = np.random.default_rng(4)
    q = rng.standard_normal((7, 24)).astype(np.float32)
    g = rng.standard_normal((40, 24)).astype(np.float32)
    c = 3
    v = g.reshape(40, c, 8)
    qq = q.reshape(7, c, 8)
    chunk_dp = np.einsum("ncp,mcp->nmc", v.astype(np.float64), qq.astype(np.float64))   # :398
    sim = chunk_dp.max(axis=2)                                                            # :399  [G, Q]
    ref_idx = np.argsort(-sim.T, axis=1, kind="stable")[:, :5]
    vals, idx = osim.sim_topk_split(q, g, 5, c)
    assert np.array_equal(idx, ref_idx)
    np.testing.assert_allclose(vals, np.take_along_axis(sim.T, ref_idx, axis=1), atol=1e-6)
    # one part == the plain dot product
    v1, i1 = osim.sim_topk_split(q, g, 5, 1)
    v0, i0 = osim.sim_topk(q, g, 5)
    assert np.array_equal(i1, i0) and np.array_equal(v1, v0)

"""Oracle for the similarity / top-k step (test infrastructure only; see oracle/__init__.py).

Reference lines restated:
  diff_retrieval.py:388-389   nn.functional.normalize(x, dim=1, p=2)
  diff_retrieval.py:402       sim = torch.mm(values_features, query_features.T)
  diff_retrieval.py:411,417   simscores = sim.T ; simscores.topk(k, axis=1, largest=True)
  diff_retrieval.py:403,418-419  sim2 = mm(values, values.T); bg = sim2.T.topk(2)[0][:, -1]
  diff_retrieval.py:442-454   summary statistics
  embedding_search/similarity_search.py:39-83  chunked top-1 with running merge

torch.topk does not define the order of equal scores, and MKL's fp32 summation order
depends on the thread count, so the reference itself is only defined up to fp32 rounding noise.  The oracle's
contract, which the CUDA path reproduces bit-exactly on indices:
    score(q, g) = sum_i q_i * g_i evaluated in float64 from the float32 inputs, reported as float32;
    ranking by (score descending, gallery index ascending).
`sim_topk_reference_fp32` is the literal fp32 restatement used to show both agree wherever the reference is
well defined (no two candidates within fp32 noise of each other).
"""
from __future__ import annotations

import numpy as np


def l2_normalize(x: np.ndarray, eps: float = 1e-12) -> np.ndarray:
    """nn.functional.normalize(x, dim=1, p=2): x / max(||x||_2, eps), float32 arithmetic (diff_retrieval.py:388)."""
    x = np.asarray(x, dtype=np.float32)
    n = np.sqrt(np.sum(x.astype(np.float32) ** 2, axis=1, keepdims=True, dtype=np.float32)).astype(np.float32)
    return (x / np.maximum(n, np.float32(eps))).astype(np.float32)


def _rank_row(scores: np.ndarray, k: int) -> np.ndarray:
    """indices of the k largest entries of a 1-D float64 array, ordered by (score desc, index asc)."""
    n = scores.shape[0]
    if k >= n:
        cand = np.arange(n)
    else:
        kth = np.partition(scores, n - k)[n - k]          # k-th largest value
        cand = np.nonzero(scores >= kth)[0]               # every tie of the k-th value is a candidate
    order = np.lexsort((cand, -scores[cand]))             # primary: -score, secondary: index
    return cand[order[:k]]


def sim_topk(q: np.ndarray, g: np.ndarray, k: int, chunk: int = 256):
    """(values f32[Q,k], indices i64[Q,k]) of the k largest dot products per query; see module docstring."""
    q = np.asarray(q, dtype=np.float32)
    g = np.asarray(g, dtype=np.float32)
    assert q.ndim == 2 and g.ndim == 2 and q.shape[1] == g.shape[1]
    assert 1 <= k <= g.shape[0]
    g64 = g.astype(np.float64)
    vals = np.empty((q.shape[0], k), dtype=np.float32)
    idx = np.empty((q.shape[0], k), dtype=np.int64)
    for s in range(0, q.shape[0], chunk):
        S = q[s:s + chunk].astype(np.float64) @ g64.T     # [chunk, G] float64
        for r in range(S.shape[0]):
            top = _rank_row(S[r], k)
            idx[s + r] = top
            vals[s + r] = S[r, top].astype(np.float32)
    return vals, idx


def sim_topk_split(q: np.ndarray, g: np.ndarray, k: int, num_chunks: int, chunk: int = 128, cross: bool = False):
    """'splitloss' similarity, diff_retrieval.py:393-400: v,q -> [b, c, p]; chunk_dp = einsum('ncp,mcp->nmc');
    sim = max over c; then the same top-k as sim_topk (float64 dot products per part, ranked on the float64 value,
    reported as float32, ties by lowest gallery index)."""
    q = np.asarray(q, dtype=np.float32)
    g = np.asarray(g, dtype=np.float32)
    d = q.shape[1]
    assert d % num_chunks == 0
    p = d // num_chunks
    g64 = g.astype(np.float64).reshape(g.shape[0], num_chunks, p)
    vals = np.empty((q.shape[0], k), dtype=np.float32)
    idx = np.empty((q.shape[0], k), dtype=np.int64)
    for s in range(0, q.shape[0], chunk):
        qq = q[s:s + chunk].astype(np.float64).reshape(-1, num_chunks, p)
        if cross:   # einsum_in_chunks stype='cross', diff_retrieval.py:652-654: 'ncp,mdp->nmcd' then max over (c, d)
            S = np.einsum("mdp,ncp->mncd", qq, g64).max(axis=(2, 3))
        else:
            S = np.einsum("mcp,ncp->mnc", qq, g64).max(axis=2)      # [chunk, G]
        for r in range(S.shape[0]):
            top = _rank_row(S[r], k)
            idx[s + r] = top
            vals[s + r] = S[r, top].astype(np.float32)
    return vals, idx


def sim_topk_reference_fp32(q, g, k: int):
    """Literal restatement of diff_retrieval.py:402,411,417 on CPU fp32 torch (tie order unspecified)."""
    import torch
    qt = torch.from_numpy(np.asarray(q, dtype=np.float32))
    gt = torch.from_numpy(np.asarray(g, dtype=np.float32))
    sim = torch.mm(gt, qt.T)
    v, l = sim.T.topk(k, dim=1, largest=True)
    return v.numpy(), l.numpy()


def background_second_best(g: np.ndarray, chunk: int = 256) -> np.ndarray:
    """bg_v of diff_retrieval.py:403,418-419: second largest entry of every row of G.G^T (the largest is assumed
    to be the row itself)."""
    v, _ = sim_topk(g, g, 2, chunk=chunk)
    return v[:, -1]


def merge_topk(scores: np.ndarray, idx: np.ndarray, k_out: int):
    """scores/idx [nlists, Q, k_in] -> [Q, k_out] by (score desc, idx asc); idx < 0 marks empty entries.
    Equivalent of the running best-of-folders merge in embedding_search/similarity_search.py:70-74 (k = 1) and of
    concatenating the shards before topk."""
    nl, nq, k_in = scores.shape
    out_s = np.empty((nq, k_out), dtype=np.float32)
    out_i = np.empty((nq, k_out), dtype=np.int64)
    for r in range(nq):
        s = scores[:, r, :].reshape(-1)
        i = idx[:, r, :].reshape(-1)
        keep = i >= 0
        s, i = s[keep], i[keep]
        order = np.lexsort((i, -s.astype(np.float64)))[:k_out]
        out_s[r] = s[order]
        out_i[r] = i[order]
    return out_s, out_i


def retrieval_stats(main_v: np.ndarray, bg_v: np.ndarray) -> dict:
    """The printed / wandb dictionary of diff_retrieval.py:442-468 (same keys)."""
    x0 = np.asarray(main_v, dtype=np.float32).reshape(-1)
    x1 = np.asarray(bg_v, dtype=np.float32).reshape(-1)
    return {
        "sim_mean": float(np.mean(x0)), "sim_std": float(np.std(x0)),
        "sim_75pc": float(np.percentile(x0, 75)), "sim_90pc": float(np.percentile(x0, 90)),
        "sim_95pc": float(np.percentile(x0, 95)),
        "sim_gt_05pc": float(np.sum(x0 > 0.5) / x0.shape[0]),
        "bg_mean": float(np.mean(x1)), "bg_std": float(np.std(x1)),
        "bg_75pc": float(np.percentile(x1, 75)), "bg_90pc": float(np.percentile(x1, 90)),
        "bg_95pc": float(np.percentile(x1, 95)),
    }

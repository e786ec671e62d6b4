"""Gallery-sharded retrieval over torch.distributed (one process per GPU, NCCL over NVLink/NVSwitch).

The reference's multi-GPU path (diff_retrieval.py:237-246, 288-317, 345-348; utils_ret.py:762-786) shards only the
gallery *loader* with a DistributedSampler, all_gathers (index, feats) after every batch and funnels everything to
rank 0, which then does the similarity alone (and it dead-locks as committed -- SURVEY.md 3.2).  Here each rank keeps
its 1/N of the gallery descriptors resident, every rank scores ALL queries against its shard with the fused kernel,
and one all-gather of the [Q,k] (score, index) pairs + a merge gives every rank the global top-k:

    rank r: embeds gallery rows [r*G/N, (r+1)*G/N) and queries [r*Q/N, (r+1)*Q/N)
    all_gather_into_tensor(query descriptors, f32)      Q*D*4 bytes total            (one collective)
    local fused sim+top-k with global index = base + local
    all_gather_into_tensor(packed (score, index))       N*Q*k*12 bytes               (one collective)
    merge N lists -> top-k by (score desc, index asc)   == top-k over the concatenated gallery

There is no data-path collective inside the kernels: the exchange is two small all-gathers per run (message sizes
are KBs..MBs, latency bound), so NCCL is the right tool.  The functions take the local scorer / merger as arguments
so the sharding logic is testable on CPU with the oracle under the gloo backend (tests/test_dist_cpu.py).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n items for `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_rows(x: torch.Tensor, sizes: Optional[list] = None) -> torch.Tensor:
    """Concatenate row blocks of every rank (blocks may differ in length by one).  One collective into one pre-sized
    buffer (`all_gather_into_tensor`): equal blocks land in place, ragged blocks are padded to the longest and compacted
    afterwards.  Descriptors travel as float32: the scores are float64-accumulated products of the float32 inputs, and
    rounding the queries to bf16 for the wire (3x fewer bytes of a transfer that is already < 0.5 ms over NVLink at
    50k x 512) would change them."""
    world = dist.get_world_size()
    if world == 1:
        return x
    if sizes is None:
        n = torch.tensor([x.shape[0]], device=x.device, dtype=torch.int64)
        ns = torch.empty(world, device=x.device, dtype=torch.int64)
        dist.all_gather_into_tensor(ns, n)
        sizes = [int(v) for v in ns.tolist()]
    mx = max(sizes)
    x = x.contiguous()
    if x.shape[0] != mx:
        pad = torch.zeros((mx,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        pad[:x.shape[0]] = x
        x = pad
    buf = torch.empty((world * mx,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(buf, x)
    if all(s == mx for s in sizes):
        return buf
    return torch.cat([buf[r * mx:r * mx + s] for r, s in enumerate(sizes)], dim=0)


def _pack_topk(s: torch.Tensor, i: torch.Tensor) -> torch.Tensor:
    """(scores f32 [Q,k], indices i64 [Q,k]) -> one int32 [Q,k,3] message: score bits, index low / high words."""
    iv = i.contiguous().view(torch.int32).view(i.shape[0], i.shape[1], 2)
    return torch.cat([s.contiguous().view(torch.int32).unsqueeze(-1), iv], dim=-1).contiguous()


def _unpack_topk(p: torch.Tensor):
    s = p[..., 0].contiguous().view(torch.float32)
    i = p[..., 1:].contiguous().view(torch.int64).squeeze(-1)
    return s, i


def sharded_topk(query_local: torch.Tensor, gallery_local: torch.Tensor, k: int, gallery_base: int,
                 local_topk: Callable, merge: Callable, query_sizes: Optional[list] = None
                 ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Global top-k for ALL queries on every rank.  query_local: this rank's block of query descriptors;
    gallery_local: this rank's gallery shard whose first row has global index `gallery_base`.
    local_topk(q, g, k, index_base) -> (scores [Q,k], idx [Q,k]); merge(scores [N,Q,k], idx [N,Q,k], k) -> ([Q,k],[Q,k])."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    q_all = all_gather_rows(query_local, query_sizes) if world > 1 else query_local
    kk = min(k, gallery_local.shape[0])
    s, i = local_topk(q_all, gallery_local, kk, gallery_base)
    if kk < k:   # a shard smaller than k: pad with empty entries
        pad_s = torch.full((s.shape[0], k - kk), float("-inf"), dtype=s.dtype, device=s.device)
        pad_i = torch.full((i.shape[0], k - kk), -1, dtype=i.dtype, device=i.device)
        s, i = torch.cat([s, pad_s], 1), torch.cat([i, pad_i], 1)
    if world == 1:
        return s, i
    # ONE collective for the per-shard lists: (score, index) packed into 12 bytes per entry
    msg = _pack_topk(s, i)
    allmsg = torch.empty((world * msg.shape[0],) + tuple(msg.shape[1:]), dtype=msg.dtype, device=msg.device)
    dist.all_gather_into_tensor(allmsg, msg)                   # rank-major concatenation along dim 0
    ss, ii = _unpack_topk(allmsg.view((world,) + tuple(msg.shape)))
    return merge(ss.contiguous(), ii.contiguous(), k)


def cuda_local_topk(q, g, k, index_base):
    from .similarity import sim_topk
    return sim_topk(q, g, k, index_base=index_base)


def cuda_merge(scores, idx, k):
    from .similarity import topk_merge
    return topk_merge(scores, idx, k)


def sharded_topk_c(query_all: torch.Tensor, gallery_local: torch.Tensor, k: int, gallery_base: int,
                   allgather: Optional[Callable] = None, world: Optional[int] = None, index_stride: int = 1
                   ) -> Tuple[torch.Tensor, torch.Tensor]:
    """The same exchange through the C entry `dcr_sim_topk_sharded` (include/dcr_b200.h): local fused top-k, ONE all-gather
    of the packed lists, merge -- all enqueued by the library on the current stream.  `query_all`: every query descriptor
    (already all-gathered); `allgather(send_ptr, recv_ptr, bytes_per_rank, stream_ptr) -> int` performs the collective
    (default: torch.distributed.all_gather_into_tensor over uint8 views of the two device buffers)."""
    import ctypes as C
    from . import _lib
    from .similarity import _aligned_ptr, _check_cuda_f32
    lib = _lib.load()
    q = _check_cuda_f32("query_all", query_all)
    g = _check_cuda_f32("gallery_local", gallery_local)
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    nq, d = q.shape
    ng = g.shape[0]
    holders = []

    def default_allgather(send, recv, nbytes, stream):
        # zero-copy uint8 views of the library's device buffers
        sv = device_bytes(send, nbytes, q.device)
        rv = device_bytes(recv, nbytes * world, q.device)
        holders.extend([sv, rv])
        dist.all_gather_into_tensor(rv, sv)
        return 0

    fn = allgather or default_allgather

    def trampoline(send, recv, nbytes, ctx, stream):
        try:
            return int(fn(send, recv, nbytes, stream))
        except Exception as e:                      # never unwind through the C frame
            print(f"dcr_b200.dist.sharded_topk_c: all-gather callback raised {e!r}")
            return 1

    cb = _lib.ALLGATHER_FN(trampoline)
    with torch.cuda.device(q.device):
        nbytes = lib.dcr_sim_topk_sharded_workspace_size(nq, ng, d, k, world)
        if nbytes == 0:
            raise _lib.DcrError(f"dcr_sim_topk_sharded_workspace_size: {_lib.last_error()}")
        ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=q.device)
        out_s = torch.empty((nq, k), dtype=torch.float32, device=q.device)
        out_i = torch.empty((nq, k), dtype=torch.int64, device=q.device)
        st = torch.cuda.current_stream().cuda_stream
        rc = lib.dcr_sim_topk_sharded(q.data_ptr(), nq, g.data_ptr(), ng, d, k, gallery_base, index_stride, world,
                                      C.cast(cb, C.c_void_p), None, out_s.data_ptr(), out_i.data_ptr(), _aligned_ptr(ws), nbytes, st)
        _lib.check(rc, "dcr_sim_topk_sharded")
        torch.cuda.current_stream().synchronize()    # the workspace and the views above die with this frame
    return out_s, out_i


def device_bytes(ptr: int, nbytes: int, device: torch.device) -> torch.Tensor:
    """Zero-copy uint8 tensor over a raw device pointer (the library's workspace), for handing it to torch.distributed."""
    iface = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}
    holder = type("_DevicePtr", (), {"__cuda_array_interface__": iface})()
    return torch.as_tensor(holder, device=device)

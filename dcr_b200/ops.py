"""Single-op host wrappers over the C ABI (used by the unit tests and by tools; the networks run through
dcr_b200.nets, which drives the same kernels from the C++ executor)."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib

TERMS_FOR_PLANES = {1: 1, 2: 3, 3: 6}


def split_planes(x: torch.Tensor, planes: int) -> torch.Tensor:
    """fp32 tensor -> bf16 planes [planes, *x.shape] with x ~= sum(planes) (hi, mid, lo)."""
    x = x.float()
    out = []
    r = x
    for _ in range(planes):
        h = r.to(torch.bfloat16)
        out.append(h)
        r = r - h.float()
    return torch.stack(out).contiguous()


def merge_planes(p: torch.Tensor) -> torch.Tensor:
    return p.float().sum(dim=0)


def prepare_conv_weight(w: torch.Tensor, planes: int = 1) -> torch.Tensor:
    """[N, C, kh, kw] fp32 (or [N, K] for Linear) -> bf16 [planes, N, kh*kw*ceil64(C)], tap-major, zero padded."""
    if w.dim() == 2:
        w = w[:, :, None, None]
    n, c, kh, kw = w.shape
    cp = (c + 63) // 64 * 64
    wt = torch.zeros((n, kh, kw, cp), dtype=torch.float32, device=w.device)
    wt[..., :c] = w.float().permute(0, 2, 3, 1)
    return split_planes(wt.reshape(n, kh * kw * cp), planes)


def conv2d(x: torch.Tensor, w_prepared: torch.Tensor, n_out: int, kh: int = 1, kw: int = 1, stride: int = 1,
           pad_h: int = 0, pad_w: int = 0, *, scale: Optional[torch.Tensor] = None,
           bias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None, act: int = 0,
           terms: Optional[int] = None, out_planes: Optional[int] = None, want_f32: bool = False):
    """x: bf16 planes [P, B, H, W, C] (CUDA).  Returns (out planes [Po, B, Ho, Wo, N] bf16, out_f32 or None)."""
    lib = _lib.load()
    assert x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 5 and x.is_contiguous()
    p, b, h, w_, c = x.shape
    wp = w_prepared.shape[0]
    terms = TERMS_FOR_PLANES[min(p, wp)] if terms is None else terms
    out_planes = p if out_planes is None else out_planes
    ho = (h + 2 * pad_h - kh) // stride + 1
    wo = (w_ + 2 * pad_w - kw) // stride + 1
    out = torch.empty((out_planes, b, ho, wo, n_out), dtype=torch.bfloat16, device=x.device)
    out32 = torch.empty((b, ho, wo, n_out), dtype=torch.float32, device=x.device) if want_f32 else None
    if residual is not None:
        assert residual.shape[1:] == out.shape[1:] and residual.dtype == torch.bfloat16 and residual.is_contiguous()
    with torch.cuda.device(x.device):
        st = torch.cuda.current_stream().cuda_stream
        rc = lib.dcr_conv2d_bf16(
            x.data_ptr(), p, x[0].numel(), b, h, w_, c,
            w_prepared.data_ptr(), wp, w_prepared[0].numel(), n_out, kh, kw, stride, pad_h, pad_w, terms,
            scale.data_ptr() if scale is not None else None, bias.data_ptr() if bias is not None else None,
            residual.data_ptr() if residual is not None else None,
            residual.shape[0] if residual is not None else 0, residual[0].numel() if residual is not None else 0,
            act, out.data_ptr(), out_planes, out[0].numel(), n_out, 0,
            out32.data_ptr() if out32 is not None else None, st)
        _lib.check(rc, "dcr_conv2d_bf16")
    return out, out32

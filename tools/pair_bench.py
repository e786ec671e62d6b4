"""Per-layer A/B of the single-CTA GEMM kernel against its CTA-pair (cta_group::2) form on the plain 1x1 / Linear layers of the
networks at batch 256.  DCR_GEMM_CG2 = 0 / 1 forces one or the other (DCR_B200_TUNING=1)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DCR_B200_TUNING"] = "1"
from dcr_b200 import ops   # noqa: E402

torch.cuda.set_device(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
# (label, rows per image, K, N, residual, act)
LAYERS = [
    ("l1 conv1 64<-256", 3136, 256, 64, False, 1), ("l1 expand 256<-64", 3136, 64, 256, True, 1),
    ("l2 conv1 128<-512", 784, 512, 128, False, 1), ("l2 expand 512<-128", 784, 128, 512, True, 1),
    ("l2.0 conv1 128<-256", 3136, 256, 128, False, 1),
    ("l3 conv1 256<-1024", 196, 1024, 256, False, 1), ("l3 expand 1024<-256", 196, 256, 1024, True, 1),
    ("l3.0 conv1 256<-512", 784, 512, 256, False, 1),
    ("l4 conv1 512<-2048", 49, 2048, 512, False, 1), ("l4 expand 2048<-512", 49, 512, 2048, True, 1),
    ("l4.0 conv1 512<-1024", 196, 1024, 512, False, 1),
    ("vit qkv 1152<-384", 197, 384, 1152, False, 0), ("vit proj 384<-384", 197, 384, 384, True, 0),
    ("vit fc1 1536<-384", 197, 384, 1536, False, 2), ("vit fc2 384<-1536", 197, 1536, 384, True, 0),
    ("vitb qkv 2304<-768", 197, 768, 2304, False, 0), ("vitb fc1 3072<-768", 197, 768, 3072, False, 2),
    ("vitb fc2 768<-3072", 197, 3072, 768, True, 0),
]
gen = torch.Generator(device="cuda").manual_seed(0)
for label, rows, k, n, with_res, act in LAYERS:
    m = B * rows
    x = ops.split_planes(torch.randn(m, 1, 1, k, device="cuda", generator=gen), 1)
    w = ops.prepare_conv_weight(torch.randn(n, k, 1, 1, device="cuda", generator=gen) / k ** 0.5, 1)
    res = ops.split_planes(torch.randn(m, 1, 1, n, device="cuda", generator=gen), 1) if with_res else None
    scale = torch.ones(n, device="cuda")
    bias = torch.zeros(n, device="cuda")
    t = {}
    outs = {}
    for mode in ("0", "1", "0", "1"):
        os.environ["DCR_GEMM_CG2"] = mode
        for _ in range(3):
            o, _ = ops.conv2d(x, w, n, 1, 1, 1, 0, 0, scale=scale, bias=bias, residual=res, act=act)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            o, _ = ops.conv2d(x, w, n, 1, 1, 1, 0, 0, scale=scale, bias=bias, residual=res, act=act)
        e1.record()
        torch.cuda.synchronize()
        t[mode] = min(t.get(mode, 1e9), e0.elapsed_time(e1) * 100)
        outs[mode] = o
    fl = 2.0 * m * n * k
    print(f"{label:24s} M={m:7d} one CTA {t['0']:7.1f} us ({fl / t['0'] / 1e6:6.0f} TF)   pair {t['1']:7.1f} us ({fl / t['1'] / 1e6:6.0f} TF)   "
          f"{t['0'] / t['1']:.2f}x  identical: {torch.equal(outs['0'], outs['1'])}", flush=True)
    del x, w, res, o, outs
    torch.cuda.empty_cache()

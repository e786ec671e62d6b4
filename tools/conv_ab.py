"""A/B of one convolution shape under a tuning environment variable: python tools/conv_ab.py VAR v0 v1 B H W C N k stride pad
(a value of "-" leaves the variable unset)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DCR_B200_TUNING"] = "1"
from dcr_b200 import ops   # noqa: E402

var, v0, v1 = sys.argv[1:4]
b, h, w_, c, n, k, stride, pad = [int(a) for a in sys.argv[4:12]]
torch.cuda.set_device(0)
gen = torch.Generator(device="cuda").manual_seed(0)
x = ops.split_planes(torch.randn(b, h, w_, c, device="cuda", generator=gen), 1)
w = ops.prepare_conv_weight(torch.randn(n, c, k, k, device="cuda", generator=gen) / (c * k * k) ** 0.5, 1)
t, outs = {}, {}
for mode in (v0, v1, v0, v1):
    if mode == "-":
        os.environ.pop(var, None)      # flags are tested for presence
    else:
        os.environ[var] = mode
    for _ in range(3):
        o, _ = ops.conv2d(x, w, n, k, k, stride, pad, pad, act=1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        o, _ = ops.conv2d(x, w, n, k, k, stride, pad, pad, act=1)
    e1.record()
    torch.cuda.synchronize()
    t[mode] = min(t.get(mode, 1e9), e0.elapsed_time(e1) * 100)
    outs[mode] = o
print(f"{var}: {v0} -> {t[v0]:.1f} us   {v1} -> {t[v1]:.1f} us   identical: {torch.equal(outs[v0], outs[v1])}")

"""Fused Toeplitz stem (stem_fused.cu) against the space-to-depth stem (conv_gemm.cu): descriptors agree to bf16-network
noise, both agree with the bf16-point CPU oracle, and the forward time at batch 256 with either stem."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcr_b200 import nets, synthetic   # noqa: E402
from oracle import models as om        # noqa: E402

torch.cuda.set_device(0)
sd = om.make_sscd_state_dict(0)
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
small = synthetic.images(5, seed=3)
ref = om.sscd_forward(sd, om.preprocess(small), bf16_points=True)
outs = {}
for stem in ("s2d", "toeplitz", "toeplitz_pool"):
    net = nets.build_sscd_resnet50(sd, max_batch=8, precision="fast", stem=stem)
    outs[stem] = net(small.cuda()).cpu()
    print(stem, "max |d| vs bf16-point oracle", float((outs[stem] - ref).abs().max()), flush=True)
    del net
print("toeplitz vs s2d: max |d|", float((outs["toeplitz"] - outs["s2d"]).abs().max()),
      "min cos", float(torch.nn.functional.cosine_similarity(outs["toeplitz"], outs["s2d"], dim=1).min()), flush=True)
print("toeplitz_pool vs toeplitz: bit-identical", bool(torch.equal(outs["toeplitz_pool"], outs["toeplitz"])),
      "max |d|", float((outs["toeplitz_pool"] - outs["toeplitz"]).abs().max()), flush=True)
for sf in (0.5, 1 / 2 ** 0.5):      # multiscale input size (112 -> 56 x 56 stem output): exercises partial tiles / another pitch
    a = nets.build_sscd_resnet50(sd, max_batch=8, precision="fast", stem="s2d", scale_factor=sf)(small.cuda()).cpu()
    b = nets.build_sscd_resnet50(sd, max_batch=8, precision="fast", stem="toeplitz", scale_factor=sf)(small.cuda()).cpu()
    c = nets.build_sscd_resnet50(sd, max_batch=8, precision="fast", stem="toeplitz_pool", scale_factor=sf)(small.cuda()).cpu()
    print(f"scale {sf:.3f}: toeplitz vs s2d max |d|", float((a - b).abs().max()), " pooled == unpooled:", bool(torch.equal(b, c)), flush=True)
big = synthetic.images(32, seed=4).cuda().repeat((batch + 31) // 32, 1, 1, 1)[:batch].contiguous()
for stem in ("s2d", "toeplitz", "toeplitz_pool"):
    net = nets.build_sscd_resnet50(sd, max_batch=batch, precision="fast", stem=stem)
    for _ in range(3):
        net(big)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        net(big)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"batch {batch} stem={stem}: {ms:.3f} ms ({batch / ms * 1e3:.0f} img/s)", flush=True)
    del net

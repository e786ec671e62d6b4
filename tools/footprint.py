"""Device memory of one executor (activation buffers for max_batch images + weights) per network and precision mode, and of its
fork (activations only): cudaMemGetInfo before / after."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcr_b200 import nets               # noqa: E402
from oracle import models as om         # noqa: E402

torch.cuda.set_device(0)
torch.zeros(1).cuda()


def used():
    torch.cuda.synchronize()
    free, total = torch.cuda.mem_get_info()
    return (total - free) / 2 ** 30


cases = [("SSCD ResNet-50", lambda p: nets.build_sscd_resnet50(om.make_sscd_state_dict(0), max_batch=384, precision=p), 384),
         ("DINO ViT-S/16", lambda p: nets.build_dino_vit(om.make_vit_state_dict(0), max_batch=384, precision=p), 384),
         ("FID Inception", lambda p: nets.build_fid_inception(om.make_inception_state_dict(0), max_batch=200, precision=p), 200)]
for name, build, b in cases:
    for prec in ("fast", "bf16x3", "parity"):
        u0 = used()
        net = build(prec)
        u1 = used()
        twin = net.fork()
        u2 = used()
        print(f"{name:16s} batch {b:3d} {prec:7s}: executor {u1 - u0:6.2f} GiB, fork {u2 - u1:6.2f} GiB", flush=True)
        del net, twin
        torch.cuda.empty_cache()

"""Throughput of a descriptor network on synthetic uint8 images already resident on the GPU."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dcr_b200 import nets, synthetic
from oracle import models as om

kind, batch, precision = sys.argv[1], int(sys.argv[2]), sys.argv[3]
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 5
if kind == "sscd":
    net = nets.build_sscd_resnet50(om.make_sscd_state_dict(0), max_batch=batch, precision=precision)
elif kind == "vit":
    net = nets.build_dino_vit(om.make_vit_state_dict(0), max_batch=batch, precision=precision)
else:
    net = nets.build_fid_inception(om.make_inception_state_dict(0), max_batch=batch, precision=precision)
if kind == "inception":
    img = torch.randint(0, 256, (min(batch, 64), 299, 299, 3), dtype=torch.uint8).cuda()
else:
    img = synthetic.images(min(batch, 64), seed=0).cuda()
img = img.repeat((batch + img.shape[0] - 1) // img.shape[0], 1, 1, 1)[:batch].contiguous()
for _ in range(2):
    net(img)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    net(img)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"NET {kind} batch={batch} precision={precision} ms={ms:.3f} img_per_s={batch / ms * 1e3:.0f} "
      f"gflop_per_img={net.flops_per_image / 1e9:.2f} tflops={net.flops_per_image * batch / ms / 1e9:.1f}", flush=True)

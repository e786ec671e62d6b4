"""A/B timing of the SSCD ResNet-50 forward (fast mode) with and without tuning variables:
   python tools/ab_env.py [batch] VAR=VALUE [VAR=VALUE ...]      (each given assignment is one B variant; DCR_B200_TUNING=1 is implied)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcr_b200 import nets, synthetic   # noqa: E402
from oracle import models as om        # noqa: E402

torch.cuda.set_device(0)
args = sys.argv[1:]
batch = int(args.pop(0)) if args and args[0].isdigit() else 384
sd = om.make_sscd_state_dict(0)
img = synthetic.images(32, seed=4).cuda().repeat((batch + 31) // 32, 1, 1, 1)[:batch].contiguous()
net = nets.build_sscd_resnet50(sd, max_batch=batch, precision="fast")
os.environ["DCR_B200_TUNING"] = "1"


def timeit(reps=10):
    for _ in range(3):
        net(img)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = net(img)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


base_ms, base_out = timeit()
print(f"batch {batch} baseline: {base_ms:.3f} ms ({batch / base_ms * 1e3:.0f} img/s)", flush=True)
for a in args:
    k, v = a.split("=", 1)
    os.environ[k] = v
    ms, out = timeit()
    os.environ.pop(k)
    print(f"  {a}: {ms:.3f} ms ({batch / ms * 1e3:.0f} img/s)  same bits: {bool(torch.equal(out, base_out))}", flush=True)
ms, _ = timeit()
print(f"baseline again: {ms:.3f} ms", flush=True)

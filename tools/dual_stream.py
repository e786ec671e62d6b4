"""Two independent forward passes in flight (two networks, two streams, alternate batches) against one network on one
stream: the persistent kernels of the second stream fill the SMs the first leaves idle at its wave tails and ramps."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcr_b200 import nets, synthetic   # noqa: E402
from oracle import models as om        # noqa: E402

torch.cuda.set_device(0)
sd = om.make_sscd_state_dict(0)
base = synthetic.images(32, seed=4).cuda()


def rate(n_nets, batch, reps=8):
    ns = [nets.build_sscd_resnet50(sd, max_batch=batch, precision="fast") for _ in range(n_nets)]
    streams = [torch.cuda.Stream() for _ in range(n_nets)]
    img = base.repeat((batch + 31) // 32, 1, 1, 1)[:batch].contiguous()
    outs = []
    torch.cuda.synchronize()
    for it in range(reps + 3):
        if it == 3:
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for s in streams:
                s.wait_event(e0)
        for net, s in zip(ns, streams):
            with torch.cuda.stream(s):
                outs.append(net(img))
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    same = all(torch.equal(o, outs[0]) for o in outs)
    return n_nets * batch / ms * 1e3, same


for batch in [int(a) for a in sys.argv[1:]] or [192, 256, 384, 512]:
    r1, _ = rate(1, batch)
    r2, same = rate(2, batch)
    print(f"batch {batch}: one stream {r1:.0f} img/s   two streams {r2:.0f} img/s  ({r2 / r1:.3f}x)  outputs identical: {same}", flush=True)
    torch.cuda.empty_cache()

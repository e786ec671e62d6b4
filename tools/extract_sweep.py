"""images/s of retrieval.extract_features (device-resident uint8 input, SSCD ResNet-50, fast mode) per batch size, one
forward pass at a time against two in flight (network + fork on two streams)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcr_b200 import nets, retrieval, synthetic   # noqa: E402
from oracle import models as om                   # noqa: E402

torch.cuda.set_device(0)
sd = om.make_sscd_state_dict(0)
base = synthetic.images(64, seed=4).cuda()
N = 24576
imgs = base.repeat(N // 64, 1, 1, 1).contiguous()
for batch in [int(a) for a in sys.argv[1:]] or [128, 192, 256, 320, 384, 512]:
    net = nets.build_sscd_resnet50(sd, max_batch=batch, precision="fast")
    res = {False: [], True: []}
    outs = {}
    for rnd in range(3):                      # A/B/A/B/A/B: clocks drift under the power cap, interleave the two arms
        for dual in (False, True):
            out = retrieval.extract_features(net, imgs, batch, two_in_flight=dual)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                out = retrieval.extract_features(net, imgs, batch, two_in_flight=dual)
            e1.record()
            torch.cuda.synchronize()
            res[dual].append(3 * N / e0.elapsed_time(e1) * 1e3)
            outs[dual] = out
    fmt = lambda v: "/".join(f"{x / 1e3:.1f}k" for x in v)
    print(f"batch {batch}: one in flight {fmt(res[False])} img/s   two in flight {fmt(res[True])} img/s   identical rows: "
          f"{torch.equal(outs[False], outs[True])}", flush=True)
    del net, res, out, outs
    torch.cuda.empty_cache()

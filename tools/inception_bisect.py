import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dcr_b200 import nets
from oracle import models as om
sd = om.make_inception_state_dict(0)
img = torch.randint(0, 256, (2, 299, 299, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(1))
x = om.fid_preprocess(img)
for prec in ("parity", "fast"):
    for stop in ["Conv2d_1a_3x3", "Conv2d_2a_3x3", "Conv2d_2b_3x3", "pool1", "Conv2d_3b_1x1", "Conv2d_4a_3x3", "pool2",
                 "Mixed_5b", "Mixed_5c", "Mixed_5d", "Mixed_6a", "Mixed_6b", "Mixed_6e", "Mixed_7a", "Mixed_7b", None]:
        ref = om.fid_inception_forward(sd, x, stop_after=stop)
        net = nets.build_fid_inception(sd, max_batch=2, precision=prec, stop_after=stop)
        got = net(img.cuda()).cpu()
        print(f"{prec:7s} {str(stop):16s} err={(got-ref).abs().max().item():.3e} refmax={ref.abs().max().item():.3f}", flush=True)
        del net

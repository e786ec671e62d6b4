"""One small launch of every hand-written tcgen05 / TMA / mbarrier kernel family, for compute-sanitizer
(tools/sanitize.sh runs it under memcheck, racecheck and synccheck; summaries are committed under profiles/).

Covered: sim_topk_kernel<2,*,1|2> (k = 10 and k = 1), the conversion / re-score kernels, gemm_bf16_kernel in its plain,
im2col, A-resident, residual and fp32-output forms (SSCD ResNet-50 fast + parity forward at batch 2), conv3x3_halo_kernel,
attention_tc_kernel and the LayerNorm / token kernels (ViT-S/16 forward), the pooling kernels and concat-by-offset GEMMs
(FID Inception forward), the streaming fp64 statistics kernels."""
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcr_b200 import fid, nets, similarity, synthetic   # noqa: E402
from oracle import models as om                          # noqa: E402  (seeded weights only)

which = sys.argv[1] if len(sys.argv) > 1 else "all"
torch.cuda.set_device(0)
if which in ("all", "sim"):
    q, g = synthetic.descriptors(300, 5000, 512, seed=7)
    for k in (10, 1):
        v, i = similarity.sim_topk(q.cuda(), g.cuda(), k)
        torch.cuda.synchronize()
        print("sim_topk k=%d ok" % k, similarity.sim_topk_stats())
if which in ("all", "sscd"):
    img = synthetic.images(2, seed=1).cuda()
    for prec in ("fast", "parity"):
        net = nets.build_sscd_resnet50(om.make_sscd_state_dict(0), max_batch=2, precision=prec)
        out = net(img)
        torch.cuda.synchronize()
        print("sscd", prec, "ok", float(out.abs().sum()))
        del net
if which in ("all", "vit"):
    img = synthetic.images(2, seed=2).cuda()
    net = nets.build_dino_vit(om.make_vit_state_dict(0, depth=2), max_batch=2, precision="fast")
    out = net(img)
    torch.cuda.synchronize()
    print("vit ok", float(out.abs().sum()))
    del net
if which in ("all", "fid"):
    img = torch.randint(0, 256, (2, 299, 299, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(3)).cuda()
    net = nets.build_fid_inception(om.make_inception_state_dict(0), max_batch=2, precision="fast")
    st = fid.ActivationStatistics(2048)
    st.update(net(img))
    mu, sigma = st.finalize()
    print("inception + fid stats ok", float(mu.sum()))
print("sanitize_case done")

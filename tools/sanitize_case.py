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
if which in ("all", "pair"):
    # CTA-pair form of the GEMM kernel (cta_group::2), forced on small shapes, against the single-CTA kernel; then two forward
    # passes in flight (network + fork on two streams) through extract_features
    from dcr_b200 import ops, retrieval
    os.environ["DCR_B200_TUNING"] = "1"
    gen = torch.Generator(device="cuda").manual_seed(5)
    for (m, k, n, with_res, act) in [(4 * 196, 1024, 256, False, 1), (98, 512, 2048, True, 1), (600, 384, 1152, False, 2)]:
        x = ops.split_planes(torch.randn(m, 1, 1, k, device="cuda", generator=gen), 1)
        w = ops.prepare_conv_weight(torch.randn(n, k, 1, 1, device="cuda", generator=gen) / k ** 0.5, 1)
        res = ops.split_planes(torch.randn(m, 1, 1, n, device="cuda", generator=gen), 1) if with_res else None
        outs = []
        for mode in ("0", "1"):
            os.environ["DCR_GEMM_CG2"] = mode
            o, _ = ops.conv2d(x, w, n, 1, 1, 1, 0, 0, residual=res, act=act)
            torch.cuda.synchronize()
            outs.append(o)
        print("pair gemm", m, k, n, "identical:", torch.equal(outs[0], outs[1]))
    del os.environ["DCR_GEMM_CG2"]
    net = nets.build_sscd_resnet50(om.make_sscd_state_dict(0), max_batch=2, precision="fast")
    img = synthetic.images(5, seed=4)
    a = retrieval.extract_features(net, img.cuda(), 2, two_in_flight=False)
    b = retrieval.extract_features(net, img.pin_memory(), 2, two_in_flight=True)
    torch.cuda.synchronize()
    print("two in flight identical:", torch.equal(a, b))
print("sanitize_case done")

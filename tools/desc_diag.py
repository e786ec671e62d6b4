"""Diagnose the synthetic descriptor distribution used by bench.py (spread, effective rank, top-k gaps)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torchvision
import bench
from dcr_b200 import nets, similarity
from oracle import models as om

dev = torch.device("cuda")
gal = bench.gen_images_cuda(20000, seed=100, device=dev)
qry = bench.gen_images_cuda(2000, seed=200, device=dev, copies_of=gal)

def calibrated_sd(seed):
    sd = om.make_sscd_state_dict(seed)
    m = torchvision.models.resnet50(weights=None)
    tv = {k[len("backbone."):]: v for k, v in sd.items() if k.startswith("backbone.")}
    tv["fc.weight"], tv["fc.bias"] = m.fc.weight.detach(), m.fc.bias.detach()
    m.load_state_dict(tv)
    m = m.to(dev).train()
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.momentum = None
            mod.reset_running_stats()
    x = om.preprocess(gal[:512].cpu()).to(dev)
    with torch.no_grad():
        for s in range(0, 512, 128):
            m(x[s:s + 128])
    out = dict(sd)
    for k, v in m.state_dict().items():
        if "running_" in k:
            out["backbone." + k] = v.detach().cpu()
    return out

for name, sd in (("bench weights", bench.synthetic_sscd_weights(dev)),):
    net = nets.build_sscd_resnet50(sd, max_batch=128, precision="fast")
    g = net(gal); q = net(qry)
    sv = torch.linalg.svdvals(g[:4096] - g[:4096].mean(0))
    erank = float((sv.sum() ** 2) / (sv ** 2).sum())
    S = q @ g.T
    top = S.topk(11, dim=1).values
    gaps = (top[:, :-1] - top[:, 1:])
    v, i = similarity.sim_topk(q, g, 10)
    st = similarity.sim_topk_stats()
    print(f"{name}: mean|cos| {S.abs().mean():.3f} top1 mean {top[:,0].mean():.3f} top10 mean {top[:,9].mean():.3f} "
          f"median gap(1-2) {gaps[:,0].median():.2e} median gap(10-11) {gaps[:,9].median():.2e} eff.rank {erank:.1f} "
          f"sv[0..4] {[round(float(x),2) for x in sv[:5]]} second={st['n_second']} flagged={st['n_flagged']}")
    del net

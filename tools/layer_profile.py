"""Per-op device times of one forward pass (run under `ncu --metrics gpu__time_duration.sum --csv`), joined with
the op list: prints M, N, K, algorithmic FLOPs and bytes and the achieved rates.
usage: python tools/layer_profile.py run <kind> <batch>           (the process ncu wraps)
       python tools/layer_profile.py report <csv> <kind> <batch>  (offline; needs the op list -> GPU-free rebuild not possible,
                                                                   so `run` also dumps gpurun_out/ops_<kind>.json)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def describe(meta, batch):
    out = []
    for kind, a in meta:
        if kind == 1:
            h, w, c, n, kh, kw, st, ph, pw = a[2], a[3], a[4], a[6], a[7], a[8], a[9], a[10], a[11]
            ho, wo = (h + 2 * ph - kh) // st + 1, (w + 2 * pw - kw) // st + 1
            m = batch * ho * wo
            k = kh * kw * c
            byt = 2 * (batch * h * w * c + m * n + (m * n if a[14] >= 0 else 0)) + 2 * n * k
            out.append(dict(op="conv", M=m, N=n, K=k, kh=kh, kw=kw, stride=st, flops=2.0 * m * n * k, bytes=byt))
        else:
            names = {0: "im2col_u8", 10: "stem_s2d", 2: "maxpool", 3: "avgpool", 4: "gem", 5: "gap", 6: "layernorm", 7: "tokens", 8: "attention", 9: "l2norm", 11: "embed", 12: "stem_rows", 13: "stem_conv"}
            out.append(dict(op=names.get(kind, str(kind)), M=0, N=0, K=0, flops=0.0, bytes=0))
    return out

if sys.argv[1] == "run":
    import torch
    from dcr_b200 import nets, synthetic
    from oracle import models as om
    kind, batch = sys.argv[2], int(sys.argv[3])
    if kind == "sscd":
        net = nets.build_sscd_resnet50(om.make_sscd_state_dict(0), max_batch=batch, stem=(sys.argv[4] if len(sys.argv) > 4 else None),
                                       precision=os.environ.get("DCR_LP_PRECISION", "fast"))
        img = synthetic.images(32, seed=0)
    elif kind == "vit":
        net = nets.build_dino_vit(om.make_vit_state_dict(0), max_batch=batch)
        img = synthetic.images(32, seed=0)
    else:
        net = nets.build_fid_inception(om.make_inception_state_dict(0), max_batch=batch)
        img = torch.randint(0, 256, (32, 299, 299, 3), dtype=torch.uint8)
    img = img.cuda().repeat((batch + 31) // 32, 1, 1, 1)[:batch].contiguous()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(describe(net.meta, batch), open(f"gpurun_out/ops_{kind}.json", "w"))
    net(img); net(img)
    torch.cuda.synchronize()
    net(img)
    torch.cuda.synchronize()
else:
    import csv
    path, kind = sys.argv[2], sys.argv[3]
    ops = json.load(open(f"gpurun_out/ops_{kind}.json"))
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if "dcr::" in r["Kernel Name"]]
    per = len(rows) // 3
    last = rows[-per:]
    last = [r for r in last]
    tot = 0.0
    print(f"{'op':12s} {'M':>9s} {'N':>5s} {'K':>6s} {'us':>8s} {'TFLOP/s':>8s} {'GB/s':>7s}")
    oi = 0
    for r in last:
        us = float(r["Metric Value"].replace(",", "")) / 1000
        tot += us
        o = dict(ops[oi])
        if "expand_reduce_kernel<0>" in r["Kernel Name"].replace("(int)", ""):
            o["op"] = "conv3(x3)"                               # expansion-only variant of the fused kernel (in-place residual)
        elif "expand_reduce_kernel" in r["Kernel Name"]:     # conv3 (+residual) fused with the next block's conv1
            o2 = ops[oi + 1]
            # the expanded activation is written once and never re-read: drop its read from the second conv's bytes
            o = dict(op="conv3+conv1", M=o["M"], N=o["N"], K=o["K"], flops=o["flops"] + o2["flops"],
                     bytes=o["bytes"] + o2["bytes"] - 2 * o["M"] * o["N"])
            oi += 1
        oi += 1
        print(f"{o['op']:12s} {o['M']:9d} {o['N']:5d} {o['K']:6d} {us:8.1f} {o['flops'] / us / 1e6 if us else 0:8.1f} {o['bytes'] / us / 1e3 if us else 0:7.0f}")
    assert oi == len(ops), (oi, len(ops), len(last))
    print("total us", tot)

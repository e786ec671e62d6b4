#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config, on N GPUs of one node.

    metric   : embed+top-k query images/sec  (whole job: embed gallery + queries with the SSCD ResNet-50 descriptor
               network, L2-normalise, all-pairs dot-product similarity, per-query top-k)
    workload : configs[1]  "10k query x 100k gallery, SSCD ResNet-50 embed+top-k on 1 B200"; at N > 1 every rank holds
               a 100k-image gallery shard and a 10k block of queries (weak scaling; configs[4] is the N = 8 case up to
               a factor 1.25 in gallery size), all queries are scored against every shard, per-shard top-k lists are
               all-gathered and merged (dcr_b200/dist.py).
    one step : embed G_local + Q_local synthetic 256x256 uint8 images, normalise, sharded top-k (k = 10).

`value` times the step with the images already resident in HBM; `e2e` times the same step through the public API
from pinned HOST memory (H2D of every image batch and D2H of the result inside the timed region).
`roofline` is the fused similarity kernel (tensor bound), timed by CUDA events inside dcr_sim_topk.
`cpu_baseline` / `--impl reference` time the CPU restatement of the reference path (oracle/) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D_DESC = 512
K_TOP = 10
IMG = 256


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"burst": float(d["bf16_tflops"]), "sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                "hbm": float(d["hbm_gbs"]), "src": "measured"}
    return {"burst": 1590.0, "sustained": 1400.0, "hbm": 6650.0, "src": "fallback"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        reasons = []
        for i, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
            if any(s[3 + i].lower().startswith("active") for s in self.samples):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(sm)}


def gen_images_cuda(n: int, seed: int, device, copies_of=None, copy_frac: float = 0.1, chunk: int = 2048):
    """uint8 [n,256,256,3] on `device`: low-frequency random fields; a fraction are brightness/shift-augmented copies
    of `copies_of` rows (planted matches, so similarities span [0,1] as in DCR's use)."""
    out = torch.empty((n, IMG, IMG, 3), dtype=torch.uint8, device=device)
    g = torch.Generator(device=device).manual_seed(seed)
    for s in range(0, n, chunk):
        b = min(chunk, n - s)
        img = 0.4 * torch.rand((b, 3, 1, 1), device=device, generator=g)            # per-image colour offset
        for res, amp in ((4, 0.35), (16, 0.3), (64, 0.25)):                          # three noise scales, random mixing
            field = torch.rand((b, 3, res, res), device=device, generator=g)
            gain = amp * torch.rand((b, 1, 1, 1), device=device, generator=g)
            img = img + gain * torch.nn.functional.interpolate(field, size=(IMG, IMG), mode="bilinear",
                                                              align_corners=False)
        img = img + 0.04 * torch.randn((b, 3, IMG, IMG), device=device, generator=g)
        out[s:s + b] = (img.clamp_(0, 1) * 255.0).round_().to(torch.uint8).permute(0, 2, 3, 1)
    if copies_of is not None and n > 0 and copy_frac > 0:
        n_c = int(round(copy_frac * n))
        dst = torch.randperm(n, device=device, generator=g)[:n_c]
        src = torch.randint(0, copies_of.shape[0], (n_c,), device=device, generator=g)
        gain = 0.8 + 0.4 * torch.rand((n_c, 1, 1, 1), device=device, generator=g)
        sh = int(torch.randint(-8, 9, (1,), device=device, generator=g).item())
        base = torch.roll(copies_of[src].float(), shifts=(sh, -sh), dims=(1, 2)) * gain
        out[dst] = base.clamp_(0, 255).round_().to(torch.uint8)
    return out


def synthetic_sscd_weights(dev):
    """Seeded random-init SSCD ResNet-50 weights (no network access, no checkpoints), made data-consistent the way
    a freshly initialised PyTorch model becomes after its first training-mode batches: the BatchNorm running
    statistics are set from 512 synthetic images (torch ops, set-up only -- nothing of this runs in a timed region),
    and the head Linear is PCA-whitened on 2048 synthetic images.  Without this a random trunk maps every image
    to nearly the same direction; trained SSCD descriptors are spread over the sphere by construction."""
    import torchvision
    from dcr_b200 import nets
    from oracle import models as om
    sd = om.make_sscd_state_dict(0)
    cal_imgs = gen_images_cuda(512, seed=999, device=dev)
    m = torchvision.models.resnet50(weights=None)
    tv = {k[len("backbone."):]: v for k, v in sd.items() if k.startswith("backbone.")}
    tv["fc.weight"], tv["fc.bias"] = m.fc.weight.detach(), m.fc.bias.detach()
    m.load_state_dict(tv)
    m = m.to(dev).train()
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.momentum = None            # cumulative average over the calibration batches
            mod.reset_running_stats()
    x = om.preprocess(cal_imgs.cpu()).to(dev)
    with torch.no_grad():
        for s in range(0, x.shape[0], 128):
            m(x[s:s + 128])
    for k, v in m.state_dict().items():
        if "running_mean" in k or "running_var" in k:
            sd["backbone." + k] = v.detach().float().cpu()
    del m, x
    # PCA-whiten the head on 2048 synthetic images: SSCD is trained (entropy regulariser) to spread its descriptors
    # uniformly over the sphere; a random head on a random trunk concentrates them in a few directions instead.
    cal = nets.build_sscd_resnet50(sd, max_batch=128, precision="fast", l2_normalize=False)
    emb = torch.cat([cal(cal_imgs), cal(gen_images_cuda(1536, seed=998, device=dev))]).double().cpu()
    del cal
    mean = emb.mean(dim=0)
    cov = torch.cov((emb - mean).T)
    lam, u = torch.linalg.eigh(cov)
    lam = lam.clamp_min(lam.max() * 1e-6)
    wh = (u / lam.sqrt()).T                                    # Lambda^-1/2 U^T
    w, bias = sd["embeddings.1.weight"].double(), sd["embeddings.1.bias"].double()
    sd["embeddings.1.weight"] = (wh @ w).float()
    sd["embeddings.1.bias"] = (wh @ (bias - mean)).float()
    torch.cuda.empty_cache()
    return sd


def cpu_reference_sample(embed_imgs: int, sim_q: int, g_total: int, q_total: int, seed: int = 0):
    """Times the oracle (CPU restatement of the reference path) on a bounded sample and extrapolates to the workload.
    Returns (value queries/s, details)."""
    from oracle import models as om
    from oracle import similarity as osim  # noqa: F401  (documented dependency; torch.mm/topk is the literal path)
    from dcr_b200 import synthetic
    cores = torch.get_num_threads()   # torch's default: one thread per physical core of the host
    sd = om.make_sscd_state_dict(0)
    imgs = synthetic.images(embed_imgs, seed=seed)
    x = om.preprocess(imgs)
    om.sscd_forward(sd, x[:2])                       # warm the thread pool / allocator
    t0 = time.perf_counter()
    for s in range(0, embed_imgs, 64):               # loader batch 64, diff_retrieval.py:352
        om.sscd_forward(sd, x[s:s + 64])
    t_img = (time.perf_counter() - t0) / embed_imgs
    q, g = synthetic.descriptors(sim_q, g_total, D_DESC, seed=seed)
    t0 = time.perf_counter()
    sim = torch.mm(g, q.T)                           # diff_retrieval.py:402 (fp32, CPU)
    sim.T.topk(K_TOP, dim=1, largest=True)           # diff_retrieval.py:417/621
    t_sim = (time.perf_counter() - t0) * (q_total / sim_q)
    total = t_img * (g_total + q_total) + t_sim
    details = {"cores": cores, "embed_img_per_s": 1.0 / t_img, "sim_topk_s_full": t_sim,
               "sample": f"oracle SSCD ResNet-50 fp32 forward on {embed_imgs} images (batch 64) + torch.mm/topk({K_TOP}) "
                         f"on {sim_q} x {g_total} descriptors, extrapolated linearly to {q_total} queries + {g_total} gallery"}
    return q_total / total, details


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--queries", type=int, default=10000, help="queries per rank")
    ap.add_argument("--gallery", type=int, default=100000, help="gallery images per rank")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-embed-sample", type=int, default=128)
    ap.add_argument("--cpu-sim-sample", type=int, default=1000)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    q_total, g_total = args.queries * world, args.gallery * world
    config = {"workload": f"SSCD ResNet-50 embed + dot-product top-{K_TOP}: {q_total} query x {g_total} gallery "
                          f"synthetic 256x256 images ({args.queries} q + {args.gallery} g per GPU)",
              "queries": q_total, "gallery": g_total, "descriptor_dim": D_DESC, "k": K_TOP,
              "images_embedded_per_step": q_total + g_total, "parallelism": f"gallery-shard x{world}",
              "l2": f"inputs ({(args.queries + args.gallery) * IMG * IMG * 3 / 1e9:.1f} GB of images per GPU) are larger than "
                    "the 126 MB L2; no explicit flush"}

    if args.impl == "reference":
        # the reference's own CPU path, restated (the reference scripts cannot be imported/installed: torch._six,
        # clip, natsort, NCCL-only init -- SURVEY.md 8c); rank 0 only, bounded sample per step
        if rank != 0:
            return
        vals = []
        det = None
        for i in range(args.warmup + args.steps):
            v, det = cpu_reference_sample(max(32, args.cpu_embed_sample // 2), args.cpu_sim_sample, g_total, q_total, seed=i)
            if i >= args.warmup:
                vals.append(v)
        v = float(np.mean(vals))
        line = {"impl": "reference", "metric": "embed+top-k query images/sec", "value": v, "unit": "query images/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * q_total / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": "query images/s", "cores": det["cores"], "kind": "port",
                                 "sample": det["sample"]},
                "e2e": {"value": v, "unit": "query images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch.distributed as dist
    from dcr_b200 import dist as ddist
    from dcr_b200 import nets, retrieval, similarity
    from oracle import models as om   # only for the seeded weight generator and the cpu_baseline leg

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    gal_u8 = gen_images_cuda(args.gallery, seed=100 + rank, device=dev)
    net = nets.build_sscd_resnet50(synthetic_sscd_weights(dev), max_batch=args.batch, precision="fast")
    qry_u8 = gen_images_cuda(args.queries, seed=200 + rank, device=dev, copies_of=gal_u8)
    g_base, _ = ddist.shard_bounds(g_total, rank, world) if world > 1 else (0, 0)
    g_base = rank * args.gallery

    def step(gal, qry):
        gf = retrieval.extract_features(net, gal, args.batch)
        qf = retrieval.extract_features(net, qry, args.batch)
        similarity.l2_normalize_(gf)
        similarity.l2_normalize_(qf)
        return ddist.sharded_topk(qf, gf, K_TOP, g_base, ddist.cuda_local_topk, ddist.cuda_merge,
                                  query_sizes=[args.queries] * world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        kms = []
        for _ in range(steps):
            fn()
            kms.append(similarity.sim_topk_stats()["kernel_ms"])
        ev1.record()
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), kms

    for _ in range(args.warmup):
        step(gal_u8, qry_u8)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = similarity.kernel_launch_count()
    ms_total, kernel_ms = timed(lambda: step(gal_u8, qry_u8), args.steps)
    launches = similarity.kernel_launch_count() - l0
    ms_per_step = ms_total / args.steps
    value = q_total / (ms_per_step / 1e3)
    st = similarity.sim_topk_stats()

    # ---- end to end through the public API from pinned host memory ------------------------------------------------
    e2e = None
    if not args.no_e2e:
        host_kind = "pinned"
        try:
            gal_h = torch.empty(gal_u8.shape, dtype=torch.uint8, pin_memory=True)
            qry_h = torch.empty(qry_u8.shape, dtype=torch.uint8, pin_memory=True)
        except RuntimeError:        # page-locking ~22 GB per rank can fail on a crowded host: pageable copies still work
            host_kind = "pageable"
            gal_h = torch.empty(gal_u8.shape, dtype=torch.uint8)
            qry_h = torch.empty(qry_u8.shape, dtype=torch.uint8)
        gal_h.copy_(gal_u8)
        qry_h.copy_(qry_u8)
        torch.cuda.synchronize()

        def e2e_step():
            v, i = step(gal_h, qry_h)
            return v.cpu(), i.cpu()           # D2H of the step's result

        e2e_step()
        ms_e2e, _ = timed(e2e_step, args.steps)
        e2e = {"value": q_total / (ms_e2e / args.steps / 1e3), "unit": "query images/s",
               "h2d_bytes_per_step": int(gal_h.numel() + qry_h.numel()) * world,
               "d2h_bytes_per_step": int(q_total * K_TOP * 12) * world, "host_memory": host_kind}
        del gal_h, qry_h
    if rank == 0:
        sampler.stop_flag.set()
        sampler.join(timeout=2)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    k_ms = float(np.mean(kernel_ms))
    flops = 2.0 * q_total * args.gallery * D_DESC          # one launch: all queries x this rank's gallery shard
    achieved = flops / (k_ms * 1e-3) / 1e12
    traffic = None
    prof = os.path.join(ROOT, "profiles", "sim_topk_traffic.json")
    if os.path.exists(prof):
        with open(prof) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    roofline = {"kernel": "sim_topk_kernel<2> (fused Q.G^T + per-query top-k, tcgen05 cta_group::2)",
                "bound": "tensor", "achieved": achieved, "peak": peaks["sustained"], "unit": "TFLOP/s",
                "frac": achieved / peaks["sustained"], "frac_of_burst_peak": achieved / peaks["burst"],
                "peak_source": f"{peaks['src']} bf16_tflops_sustained (kernel timed inside a long step)",
                "kernel_ms": k_ms, "flops_per_launch": flops, "traffic": traffic, "launch": st}
    line = {"metric": "embed+top-k query images/sec", "value": value, "unit": "query images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config,
            "images_embedded_per_s": (q_total + g_total) / (ms_per_step / 1e3),
            "embed_tflops": net.flops_per_image * (q_total + g_total) / (ms_per_step / 1e3) / 1e12,
            "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roofline}
    # the step's time is dominated by the descriptor network (conv/linear GEMM family, many shapes): its aggregate
    # tensor throughput over the whole step, against the same measured peak, reported beside the graded kernel's roofline
    line["roofline_embed"] = {"kernel": "gemm_bf16_kernel family (SSCD ResNet-50 forward, all layers, per GPU)",
                              "bound": "tensor", "achieved": line["embed_tflops"] / world, "peak": peaks["sustained"],
                              "unit": "TFLOP/s", "frac": line["embed_tflops"] / world / peaks["sustained"]}
    if e2e is not None:
        line["e2e"] = e2e
    if world == 1:
        v, det = cpu_reference_sample(args.cpu_embed_sample, args.cpu_sim_sample, g_total, q_total)
        line["cpu_baseline"] = {"value": v, "unit": "query images/s", "cores": det["cores"], "kind": "port",
                                "sample": det["sample"], "embed_img_per_s": det["embed_img_per_s"],
                                "sim_topk_s_full": det["sim_topk_s_full"]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

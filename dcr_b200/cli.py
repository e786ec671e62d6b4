"""Drop-in command line for the retrieval hot path of diff_retrieval.py (same flags, same defaults):

    python -m dcr_b200.cli --arch resnet50_disc --similarity_metric dotproduct --pt_style sscd \
        --query_dir <generations> --val_dir <training images>          (README.md:55 of the reference)

What it does is the `if args.rank == 0:` block diff_retrieval.py:375-483 restricted to the hot path: embed both folders,
L2-normalise, top-1 (and top-`num_matches`... the reference hard-codes top-10 for its galleries, :621) matches, background
top-2, the printed statistics dictionary, and optionally FID (:597-600).  Plots, CLIP score, complexity statistics and
wandb are out of scope.  Flags that the reference parses but never reads are accepted and ignored.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch


def bool_flag(s):                                   # utils_ret.py:463-474
    falsy, truthy = {"off", "false", "0"}, {"on", "true", "1"}
    if s.lower() in falsy:
        return False
    if s.lower() in truthy:
        return True
    raise argparse.ArgumentTypeError("invalid value for a boolean flag")


def build_parser() -> argparse.ArgumentParser:
    """Mirror of the module-level parser, diff_retrieval.py:124-181."""
    p = argparse.ArgumentParser("Generic image retrieval given a path")
    p.add_argument("--query_dir", type=str, required=True, help="The inferences")
    p.add_argument("--val_dir", type=str, required=True, help="The train data")
    p.add_argument("--pt_style", default="sscd", type=str)
    p.add_argument("-a", "--arch", metavar="ARCH", default="resnet50")
    p.add_argument("-j", "--workers", default=4, type=int, metavar="N")
    p.add_argument("-b", "--batch-size", default=128, type=int, metavar="N")
    p.add_argument("--world-size", default=-1, type=int)
    p.add_argument("--rank", default=-1, type=int)
    p.add_argument("--dist-url", default="tcp://224.66.41.62:23456", type=str)
    p.add_argument("--dist-backend", default="nccl", type=str)
    p.add_argument("--seed", default=None, type=int)
    p.add_argument("--gpu", default=None, type=int)
    p.add_argument("--multiprocessing-distributed", action="store_true")
    p.add_argument("--multiscale", default=False, type=bool_flag)
    p.add_argument("--pretrained", default="", type=str)
    p.add_argument("--similarity_metric", default="dotproduct", type=str)
    p.add_argument("--num_loss_chunks", default=1, type=int)
    p.add_argument("--numpatches", default=1, type=int)
    p.add_argument("--isvit", action="store_true")
    p.add_argument("--layer", default=1, type=int)
    p.add_argument("--stype", default="", type=str, choices=["", "cross"])
    p.add_argument("--keephead", action="store_true")
    p.add_argument("--keeppredictor", action="store_true")
    p.add_argument("-ssp", "--sim_save_path", type=str, default="./similarityscores/")
    p.add_argument("--einsum_chunks", default=30, type=int)
    p.add_argument("--dontsave", action="store_true")
    p.add_argument("--num_matches", default=4, type=int)
    p.add_argument("--imsize", default=224, type=int)
    p.add_argument("--noeval", action="store_true")
    # additions of this implementation (all optional)
    p.add_argument("--weights", default="", type=str, help="state_dict / TorchScript file of the descriptor model "
                   "(default: the reference's hard-coded ./pretrainedmodels/ paths)")
    p.add_argument("--precision", default="fast", choices=["fast", "parity", "exact"],
                   help="fast: bf16 tensor cores; parity: split-bf16 tensor cores (fp32-level); exact: float64 accumulation")
    p.add_argument("--topk", default=10, type=int, help="matches kept per query (reference: 1 for the statistics, 10 for the galleries)")
    p.add_argument("--fid_weights", default="", type=str, help="pt_inception-2015-12-05 state_dict; enables FID")
    return p


SSCD_FILES = {  # diff_retrieval.py:277-283
    "resnet50": "./pretrainedmodels/sscd_disc_mixup.torchscript.pt",
    "resnet50_im": "./pretrainedmodels/sscd_imagenet_mixup.torchscript.pt",
    "resnet50_disc": "./pretrainedmodels/sscd_disc_large.torchscript.pt",
}


def load_state_dict(path: str):
    if not os.path.exists(path):
        raise FileNotFoundError(f"model weights not found: {path} (pass --weights)")
    try:
        return torch.jit.load(path, map_location="cpu").state_dict()
    except Exception:
        obj = torch.load(path, map_location="cpu")
        return obj.get("state_dict", obj) if isinstance(obj, dict) else obj.state_dict()


def build_model(args):
    from . import nets
    if args.pt_style == "sscd":
        if args.arch not in SSCD_FILES:
            raise NotImplementedError("This model type does not exist/supported for SSCD")      # :285
        sd = load_state_dict(args.weights or SSCD_FILES[args.arch])
        if args.multiscale:                                                                   # utils_ret.py:676-698
            from . import retrieval
            return [nets.build_sscd_resnet50(sd, max_batch=256, precision=args.precision, scale_factor=s)
                    for s in retrieval.MULTI_SCALES]
        return nets.build_sscd_resnet50(sd, max_batch=256, precision=args.precision)
    if args.pt_style == "dino":
        if args.arch not in ("vit_small", "vit_base"):                                          # :251-257
            raise NotImplementedError("--pt_style dino: --arch vit_small (dino_vits16) and vit_base (dino_vitb16) are "
                                      "implemented; vit_base8 / resnet50 / vit_base_cifar10 are not")
        if args.multiscale:
            raise NotImplementedError("--multiscale needs interpolated position embeddings for ViTs (dino_vits.py:213-233); "
                                      "it is implemented for the convolutional SSCD trunks")
        sd = load_state_dict(args.weights or args.pretrained)
        return nets.build_dino_vit(sd, max_batch=256, precision=args.precision)
    raise NotImplementedError(f"--pt_style {args.pt_style} is outside the embed->match hot path (DESIGN.md section 9)")


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    assert os.path.isdir(args.query_dir)                                                      # :187
    if args.similarity_metric == "splitlosscross":                                            # :188-190
        args.similarity_metric, args.stype = "splitloss", "cross"
    if args.similarity_metric not in ("dotproduct", "splitloss"):
        raise NotImplementedError(f"--similarity_metric {args.similarity_metric}")
    split = args.num_loss_chunks if args.similarity_metric == "splitloss" else 1
    cross = split > 1 and args.stype == "cross"
    from . import data, retrieval
    if args.gpu is not None:
        torch.cuda.set_device(args.gpu)
    net = build_model(args)
    query_u8, q_files = data.load_folder_u8(args.query_dir, workers=args.workers)
    values_u8, v_files = data.load_folder_u8(args.val_dir, workers=args.workers)
    # splitloss: the reference never defines sim2 on that branch (:393-403) and stops at :412; background statistics
    # are only produced for the dot-product metric
    k = min(args.topk, len(v_files))
    if cross and (k - 1) * split + 1 > 16:
        k = 15 // split + 1
        print(f"--stype cross with {split} parts: keeping the {k} best matches per query (kernel limit (k-1)*parts+1 <= 16)")
    out = retrieval.run_retrieval(net, query_u8, values_u8, k=k, with_background=(split == 1), num_loss_chunks=split,
                                  cross=cross)
    dp = os.sep.join(os.path.normpath(args.query_dir).split(os.sep)[-3:])                      # :378
    save = f"ret_plots/{dp}/images/{args.pt_style}_{args.arch}_{args.similarity_metric}{args.stype}/"   # :408
    os.makedirs(save, exist_ok=True)
    torch.save({"values": out["values"].cpu(), "indices": out["indices"].cpu(), "query_files": q_files,
                "gallery_files": v_files}, os.path.join(save, "topk.pth"))
    print("Simscores @x% part done")                                                          # :470
    print(out["stats"])
    if args.fid_weights:
        from . import fid, nets
        inc = nets.build_fid_inception(load_state_dict(args.fid_weights), max_batch=50)
        val = fid.fid_from_images(inc, fid.load_resized(args.val_dir), fid.load_resized(args.query_dir))   # :597-600
        print({"fid": val})
        out["stats"]["fid"] = val
    with open(os.path.join(save, "stats.json"), "w") as f:
        json.dump(out["stats"], f)
    return 0


if __name__ == "__main__":
    sys.exit(main())

"""LAION-scale embedding search: the host-side mirror of the reference's `embedding_search/` scripts.

    write_embedding_pkl / read_embedding_pkl   the `embedding.pkl` format {'features': f32[N,D], 'indexes': [str]*N}
                                               download_and_generate_embedding.py:93-96
    list_image_files                           ImageswithFilename.__init__   embedding_search/utils.py:115-123
                                               (flat folder, .png/.jpg only, plain str sort -- NOT natsort)
    generate_embeddings                        get_transform + extract_features_custom + dump
                                               embedding_search/utils.py:35-50,78-113; download_and_..._embedding.py:88-96
                                               (ImageNet mean/std, no L2 normalisation after the model)
    similarity_search                          similarity_search.py:39-88: per generated image the best match over all
                                               gallery folders, keys "folder:key", scores float64

The reference multiplies every gallery folder with every query chunk (`--num-chunks`, default 100, to bound the
[G_f, Q_c] matrix) and therefore re-reads every `embedding.pkl` num_chunks times (similarity_search.py:46-52).  The fused
similarity + top-1 kernel never builds that matrix, so ALL queries stay resident and every folder is read, copied
and scanned once: a reader thread unpickles folder i+1 into pinned memory and a copy stream moves it to the GPU while
the kernel scans folder i.  `num_chunks` is accepted for command-line compatibility and does not change the result.

Tie rules (the reference has none: `Tensor.max` / `ndarray.argmax`): inside a folder the lowest row wins, across
folders the first folder in sorted order wins (that is what `argmax` over `vstack([previous, current])` does,
similarity_search.py:70-74).  Queries that no folder beats keep the reference's initial values: score -1.0, key "0.0"
(:47-48; the numeric 0 becomes the string "0.0" when numpy stacks it with the string keys, :71).

As committed the reference script cannot run (`args.laion_embeddings_folders` :34,
folder paths joined without their root :52, `pkl.dump(f, dump_dict)` :90-91); this module implements what it evidently
means, the same way oracle/embedding_search.py restates it.
"""
from __future__ import annotations

import argparse
import os
import pickle as pkl
import queue
import threading
import time
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

IMAGENET_MEAN = (0.485, 0.456, 0.406)      # embedding_search/utils.py:37-39
IMAGENET_STD = (0.229, 0.224, 0.225)
_UNMATCHED_KEY = "0.0"                      # similarity_search.py:48,71


# ---------------------------------------------------------------------------------------------------------------
# embedding.pkl
def write_embedding_pkl(path: str, features, indexes: Sequence[str]) -> None:
    feats = features.detach().cpu().numpy() if isinstance(features, torch.Tensor) else np.asarray(features)
    if feats.ndim != 2 or feats.shape[0] != len(indexes):
        raise _lib.DcrError(f"embedding.pkl: features {feats.shape} do not match {len(indexes)} indexes")
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        pkl.dump({"features": np.ascontiguousarray(feats, dtype=np.float32), "indexes": list(indexes)}, f)


def read_embedding_pkl(path: str) -> Tuple[np.ndarray, List[str]]:
    with open(path, "rb") as f:
        data = pkl.load(f)
    feats = data["features"]
    if isinstance(feats, torch.Tensor):
        feats = feats.detach().cpu().numpy()
    feats = np.ascontiguousarray(feats, dtype=np.float32)
    keys = list(data["indexes"])
    if feats.ndim != 2 or feats.shape[0] != len(keys):
        raise _lib.DcrError(f"{path}: features {feats.shape} do not match {len(keys)} indexes")
    return feats, keys


def list_embedding_folders(root: str) -> List[str]:
    """Sub-folders of `root` in the order the reference visits them (sorted names, similarity_search.py:34-35,50)."""
    return sorted(x for x in os.listdir(root) if os.path.isdir(os.path.join(root, x)))


# ---------------------------------------------------------------------------------------------------------------
# image folder -> embedding.pkl
def list_image_files(root_dir: str) -> List[str]:
    return sorted(x for x in os.listdir(root_dir) if x.endswith(".png") or x.endswith(".jpg"))


def generate_embeddings(net, image_folder: str, dump_path: Optional[str] = None, batch_size: int = 128,
                        workers: int = 8) -> Tuple[torch.Tensor, List[str]]:
    """Embed every .png/.jpg of a flat folder; returns (f32 [N,D] on the GPU, file names) and, when `dump_path` is
    given, writes <dump_path>/embedding.pkl.  `net` must have been built with the ImageNet statistics
    (build_sscd_resnet50(..., mean=IMAGENET_MEAN, std=IMAGENET_STD)); descriptors are NOT re-normalised here."""
    from . import data, retrieval
    names = list_image_files(image_folder)
    if not names:
        raise _lib.DcrError(f"no .png/.jpg files in {image_folder}")
    size = 256
    imgs = torch.empty((len(names), size, size, 3), dtype=torch.uint8)
    if torch.cuda.is_available():
        imgs = imgs.pin_memory()
    paths = [os.path.join(image_folder, n) for n in names]
    if workers > 1 and len(paths) > 64:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(workers) as ex:
            for i, t in enumerate(ex.map(lambda p: data.load_image_u8(p, size), paths)):
                imgs[i] = t
    else:
        for i, p in enumerate(paths):
            imgs[i] = data.load_image_u8(p, size)
    feats = retrieval.extract_features(net, imgs, batch_size)
    if dump_path is not None:
        write_embedding_pkl(os.path.join(dump_path, "embedding.pkl"), feats, names)
    return feats, names


def expand_tar_urls(patterns: Sequence[str]) -> List[str]:
    """webdataset-style brace ranges: 'dir/{000000..000012}.tar' -> 13 paths (download_and_generate_embedding.py:85-87
    builds exactly this pattern); plain paths and globs pass through."""
    import glob
    import re
    out: List[str] = []
    for pat in patterns:
        m = re.search(r"\{(\d+)\.\.(\d+)\}", pat)
        if m:
            lo, hi, width = int(m.group(1)), int(m.group(2)), len(m.group(1))
            out.extend(pat[:m.start()] + str(i).zfill(width) + pat[m.end():] for i in range(lo, hi + 1))
        elif any(ch in pat for ch in "*?["):
            out.extend(sorted(glob.glob(pat)))
        else:
            out.append(pat)
    return out


def iter_tar_samples(tar_paths: Sequence[str], size: int = 256):
    """(uint8 [size,size,3] image, key) per webdataset sample, in shard order: the `jpg` member decoded with PIL and the
    `key` field of the `json` member -- `.decode("pil").rename(image="jpg", json="json")` + `json_preproc`
    (embedding_search/utils.py:52-62).  Resize(256) + centre crop happen here on uint8; ToTensor / Normalize are fused
    into the network's first kernel."""
    import io
    import json
    import tarfile

    import numpy as np
    from PIL import Image, ImageFile
    from torchvision import transforms
    ImageFile.LOAD_TRUNCATED_IMAGES = True                                   # utils.py:13
    tf = transforms.Compose([transforms.Resize(size), transforms.CenterCrop(size)])
    for path in tar_paths:
        with tarfile.open(path, "r") as tar:
            cur, img, key = None, None, None
            for m in tar:
                if not m.isfile():
                    continue
                base, _, ext = m.name.partition(".")
                if base != cur:
                    if cur is not None and img is not None:
                        yield img, (key if key is not None else cur)
                    cur, img, key = base, None, None
                ext = ext.lower()
                if ext in ("jpg", "jpeg"):
                    pil = Image.open(io.BytesIO(tar.extractfile(m).read())).convert("RGB")
                    img = torch.from_numpy(np.asarray(tf(pil)).copy())
                elif ext == "json":
                    key = json.loads(tar.extractfile(m).read().decode("utf-8")).get("key")
            if cur is not None and img is not None:
                yield img, (key if key is not None else cur)


def generate_embeddings_from_tars(net, tars: Sequence[str], dump_path: Optional[str] = None, batch_size: int = 128,
                                  chunk: int = 4096) -> Tuple[torch.Tensor, List[str]]:
    """The `--tars` branch of get_dataloader (embedding_search/utils.py:56-62): webdataset shards streamed through the
    descriptor network in chunks of `chunk` decoded images (bounded host memory), keys from the json members."""
    from . import retrieval
    paths = expand_tar_urls(tars)
    feats, keys = [], []
    buf = torch.empty((chunk, 256, 256, 3), dtype=torch.uint8)
    if torch.cuda.is_available():
        buf = buf.pin_memory()
    n = 0
    for img, key in iter_tar_samples(paths):
        buf[n] = img
        keys.append(key)
        n += 1
        if n == chunk:
            feats.append(retrieval.extract_features(net, buf, batch_size).clone())
            torch.cuda.synchronize()      # the pinned chunk is refilled next
            n = 0
    if n:
        feats.append(retrieval.extract_features(net, buf[:n], batch_size).clone())
        torch.cuda.synchronize()
    if not feats:
        raise _lib.DcrError(f"no jpg samples in {list(paths)}")
    out = torch.cat(feats, dim=0)
    if dump_path is not None:
        write_embedding_pkl(os.path.join(dump_path, "embedding.pkl"), out, keys)
    return out, keys


# ---------------------------------------------------------------------------------------------------------------
# search
def merge_folder_best(best_scores: torch.Tensor, best_folder: torch.Tensor, best_row: torch.Tensor,
                      cur_scores: torch.Tensor, cur_row: torch.Tensor, folder_id: int) -> None:
    """Running merge of similarity_search.py:70-74, in place: the current folder replaces the previous best only where
    it is STRICTLY larger (argmax returns the first of equal maxima, i.e. the previous best).  NaN never wins."""
    take = cur_scores > best_scores
    best_scores[take] = cur_scores[take]
    best_folder[take] = folder_id
    best_row[take] = cur_row[take]


def keys_from_matches(folders: Sequence[str], folder_keys: Sequence[Sequence[str]], best_folder, best_row) -> np.ndarray:
    """'folder:key' strings of similarity_search.py:66-67; "0.0" where nothing matched (:48)."""
    bf = np.asarray(best_folder).reshape(-1)
    br = np.asarray(best_row).reshape(-1)
    out = [(_UNMATCHED_KEY if f < 0 else folders[f] + ":" + str(folder_keys[f][r])) for f, r in zip(bf, br)]
    return np.array(out)


class _FolderReader(threading.Thread):
    """Unpickles gallery folders ahead of the GPU: puts (folder index, pinned f32 [N,D] tensor, keys) on a queue."""

    def __init__(self, root: str, folders: Sequence[str], depth: int = 2):
        super().__init__(daemon=True)
        self.root, self.folders = root, list(folders)
        self.q: "queue.Queue" = queue.Queue(maxsize=depth)

    def run(self):
        for i, name in enumerate(self.folders):
            try:
                feats, keys = read_embedding_pkl(os.path.join(self.root, name, "embedding.pkl"))
                t = torch.from_numpy(feats)
                if torch.cuda.is_available():
                    t = t.pin_memory()
                self.q.put((i, t, keys, None))
            except Exception as e:          # the reference prints the exception and skips the folder (:54-56)
                self.q.put((i, None, None, e))
        self.q.put(None)


@torch.no_grad()
def search_tensors(gen_embeddings: torch.Tensor, folder_iter: Iterable, device=None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Best match of every query row over a stream of gallery folders.  `folder_iter` yields
    (folder_id, f32 [N,D] tensor on the host or the GPU); host tensors are copied on a side stream so that the copy of
    folder i+1 overlaps the scan of folder i.  Returns (scores f32 [Q], folder id i64 [Q] (-1 = unmatched), row i64 [Q])."""
    from .similarity import sim_topk
    dev = torch.device("cuda") if device is None else torch.device(device)
    q = gen_embeddings.to(dev, torch.float32).contiguous()
    nq, d = q.shape
    best_s = torch.full((nq,), -1.0, dtype=torch.float32, device=dev)        # similarity_search.py:47
    best_f = torch.full((nq,), -1, dtype=torch.int64, device=dev)
    best_r = torch.zeros((nq,), dtype=torch.int64, device=dev)
    compute = torch.cuda.current_stream(dev)
    copy = torch.cuda.Stream(device=dev)
    pending = None      # (folder_id, device tensor, ready event)

    def stage(item):
        fid, feats = item
        if feats.dim() != 2 or feats.shape[1] != d:
            raise _lib.DcrError(f"gallery folder {fid}: features {tuple(feats.shape)} do not match query dim {d}")
        if feats.is_cuda:
            return fid, feats.float().contiguous(), None
        with torch.cuda.stream(copy):
            g = feats.to(dev, torch.float32, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy)
        return fid, g, ev

    def scan(fid, g, ev):
        if ev is not None:
            compute.wait_event(ev)
            g.record_stream(compute)
        if g.shape[0] == 0:
            return
        s, r = sim_topk(q, g, 1)
        merge_folder_best(best_s, best_f, best_r, s[:, 0], r[:, 0], fid)

    for item in folder_iter:
        nxt = stage(item)               # enqueue the copy of the next folder first ...
        if pending is not None:
            scan(*pending)              # ... then scan the previous one underneath it
        pending = nxt
    if pending is not None:
        scan(*pending)
    return best_s, best_f, best_r


def similarity_search(laion_embedding_folder: str, generation_embedding_path: str, dump_path: Optional[str] = None,
                      num_chunks: int = 100, verbose: bool = True) -> Dict[str, object]:
    """similarity_search.py:22-92.  Returns (and, with `dump_path`, pickles) {'scores': f64[Q], 'keys': str[Q],
    'gen_images': the query file's 'indexes'}."""
    del num_chunks                      # no [G_f, Q_c] matrix exists here; see the module docstring
    gen_embeddings, gen_images_fname = read_embedding_pkl(generation_embedding_path)
    folders = list_embedding_folders(laion_embedding_folder)
    if verbose:
        print(f"Number of generated images to test: {gen_embeddings.shape[0]}")
        print(f"Number of LAION chunks to test: {len(folders)}")
    reader = _FolderReader(laion_embedding_folder, folders)
    reader.start()
    folder_keys: List[Optional[List[str]]] = [None] * len(folders)
    start_time = time.time()

    def stream():
        while True:
            item = reader.q.get()
            if item is None:
                return
            i, feats, keys, err = item
            if err is not None:
                print(err)
                continue
            folder_keys[i] = keys
            yield i, feats

    best_s, best_f, best_r = search_tensors(torch.from_numpy(gen_embeddings), stream())
    torch.cuda.synchronize()
    if verbose:
        print(f"Matching took: {time.time() - start_time:.2f} secs")
    scores = best_s.cpu().numpy().astype(np.float64)        # float64 as the reference's -np.ones accumulator (:47)
    keys = keys_from_matches(folders, folder_keys, best_f.cpu().numpy(), best_r.cpu().numpy())
    out = {"scores": scores, "keys": keys, "gen_images": gen_images_fname}
    if dump_path:
        os.makedirs(os.path.dirname(os.path.abspath(dump_path)), exist_ok=True)
        with open(dump_path, "wb") as f:
            pkl.dump(out, f)
    return out


# ---------------------------------------------------------------------------------------------------------------
# command lines (flags of similarity_search.py:14-20 and download_and_generate_embedding.py:14-38)
def build_search_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser("embedding similarity search")
    p.add_argument("--laion-embedding-folder", type=str, required=True)
    p.add_argument("--generation-embedding-path", type=str, required=True)
    p.add_argument("--dump-path", type=str, required=True)
    p.add_argument("--num-chunks", "--chunks", type=int, default=100)
    return p


def build_embed_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser("generate embedding.pkl from an image folder")
    p.add_argument("--parquet-fname", type=str, default=None)
    p.add_argument("--image-folder", type=str, default=None)
    p.add_argument("--tars", nargs="+", default=[])
    p.add_argument("--dump-path", type=str, default="./data/laion_sd_v2p1/data")
    p.add_argument("--wandb", action="store_true")
    p.add_argument("--skip-download", action="store_true")
    p.add_argument("--skip-img-embed", action="store_true")
    p.add_argument("--skip-image-delete", action="store_true")
    p.add_argument("--pt-style", default="sscd", type=str)
    p.add_argument("--arch", default="resnet50", type=str)
    p.add_argument("--batch-size", type=int, default=128)
    p.add_argument("--workers", type=int, default=8)
    p.add_argument("--multiscale", action="store_true")
    p.add_argument("--gpu", type=int, default=0)
    p.add_argument("--similarity-metric", type=str, default="d")
    p.add_argument("--data-dir", type=str, default=None)
    p.add_argument("--weights", default="", type=str, help="SSCD state_dict / TorchScript file (default: the reference's "
                   "./pretrained_models/ paths, embedding_search/utils.py:18-23)")
    return p


SSCD_FILES = {  # embedding_search/utils.py:18-23
    "resnet50": "./pretrained_models/sscd_disc_mixup.torchscript.pt",
    "resnet50_im": "./pretrained_models/sscd_imagenet_mixup.torchscript.pt",
    "resnet50_disc": "./pretrained_models/sscd_disc_large.torchscript.pt",
}


def embed_main(argv=None) -> int:
    args = build_embed_parser().parse_args(argv)
    if args.parquet_fname and not args.skip_download:
        # img2dataset download (download_and_generate_embedding.py:55-83) is data acquisition over the network
        raise _lib.DcrError("--parquet-fname download is out of scope (no network): pass the downloaded shards with --tars")
    if args.image_folder is None and not args.tars:
        raise RuntimeError("Either tar files or image folder must be specified")    # embedding_search/utils.py:66
    if args.pt_style != "sscd" or args.arch not in SSCD_FILES:
        raise NotImplementedError("This model type does not exist for SSCD")        # utils.py:25-27 (constructed there, raised here)
    if args.skip_img_embed:
        return 0
    from . import cli, nets
    sd = cli.load_state_dict(args.weights or SSCD_FILES[args.arch])
    torch.cuda.set_device(args.gpu)
    net = nets.build_sscd_resnet50(sd, max_batch=min(256, max(1, args.batch_size)), mean=IMAGENET_MEAN, std=IMAGENET_STD)
    start = time.time()
    if args.tars:
        generate_embeddings_from_tars(net, args.tars, args.dump_path, args.batch_size)
    else:
        generate_embeddings(net, args.image_folder, args.dump_path, args.batch_size, args.workers)
    print(f"Embedding processing + dumping took: {time.time() - start:.2f}s")
    return 0


def search_main(argv=None) -> int:
    args = build_search_parser().parse_args(argv)
    similarity_search(args.laion_embedding_folder, args.generation_embedding_path, args.dump_path, args.num_chunks)
    return 0


if __name__ == "__main__":
    import sys
    if len(sys.argv) > 1 and sys.argv[1] in ("search", "embed"):
        sys.exit({"search": search_main, "embed": embed_main}[sys.argv[1]](sys.argv[2:]))
    print("usage: python -m dcr_b200.embedding_search {search|embed} [flags]", file=sys.stderr)
    sys.exit(2)

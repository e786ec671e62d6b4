"""Image folders -> uint8 tensors, following the reference's dataset conventions.

    SynthDataset        diff_retrieval.py:61-111   walk leaf folders, keep .JPG/.JPEG/.jpg/.png, `natsorted` full paths
                                                   (this order DEFINES the index space of every top-k result), prompts
                                                   from <dir>/prompts.txt (:87-90)
    ret_transform       diff_retrieval.py:325-330  Resize(256) -> CenterCrop(224) -> ToTensor -> Normalize
Only the PIL part runs here (decode, Resize(256) on uint8, centre crop to 256x256 -- the centre 224 crop of that equals
the reference's CenterCrop(224) of the resized image); crop-to-224, ToTensor and Normalize are fused into the first
GPU kernel (dcr_b200.nets).  natsort is not installed in this image, so the digit-aware ordering is restated here.
"""
from __future__ import annotations

import itertools
import os
import re
from typing import List, Optional, Tuple

import torch

_IMG_EXT = (".JPG", ".JPEG", ".jpg", ".png")
_num = re.compile(r"(\d+)")


def natural_key(s: str):
    """natsort's default ordering (ns.INT, no locale): runs of digits compare as integers, the rest as text."""
    parts = _num.split(s)
    return [int(p) if i % 2 else p for i, p in enumerate(parts)]


def caption_json_for(main_dir: str) -> Optional[str]:
    """diff_retrieval.py:64-69: a folder WITHOUT prompts.txt takes its file list (and prompts) from a caption json next to
    the `train` part of its path: `*combined_captions.json` for laion folders, `*blip_captions.json` for imagenette."""
    import glob
    if os.path.exists(f"{main_dir}/prompts.txt"):
        return None
    p = main_dir.split("train")[0]
    found = []
    if "laion" in main_dir or "l100kaion" in main_dir:
        found = glob.glob(os.path.join(p, "*combined_captions.json"))
    elif "imagenette" in main_dir:
        found = glob.glob(os.path.join(p, "*blip_captions.json"))
    else:
        return None
    return found[0]          # IndexError when the json is missing, as in the reference


def dataset_index(main_dir: str, capjson: Optional[str] = None) -> Tuple[List[str], Optional[List[str]]]:
    """(total_imgs, prompts) exactly as SynthDataset.__init__ builds them (diff_retrieval.py:59-97): the caption-json
    branch keeps the json's key order; the folder branch is the natsorted walk + prompts.txt lines."""
    import json
    if capjson is None:
        capjson = caption_json_for(main_dir)
    if capjson is None:
        return list_images(main_dir), load_prompts(main_dir)
    with open(capjson) as f:
        all_prompts = json.load(f)
    files = list(all_prompts.keys())
    return files, [all_prompts[k][0] for k in files]


def list_images(main_dir: str) -> List[str]:
    """The file list of SynthDataset.__init__ (diff_retrieval.py:72-86), or the caption json's keys (:92-96)."""
    capjson = caption_json_for(main_dir)
    if capjson is not None:
        return dataset_index(main_dir, capjson)[0]
    total = []
    for root, dirs, files in os.walk(main_dir, topdown=True):
        if len(dirs) > 0:
            dirs.sort()
            continue
        temp = sorted(x for x in files if x.endswith(_IMG_EXT))
        total.append([f"{root}/{w}" for w in temp])
    return sorted(itertools.chain(*total), key=natural_key)


def load_prompts(main_dir: str) -> Optional[List[str]]:
    p = os.path.join(main_dir, "prompts.txt")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        return [line for line in f]


def load_image_u8(path: str, size: int = 256) -> torch.Tensor:
    """PIL RGB -> Resize(size) (bilinear on uint8, as torchvision) -> centre crop size x size -> uint8 [size,size,3]."""
    import numpy as np
    from PIL import Image
    from torchvision import transforms
    img = Image.open(path).convert("RGB")
    img = transforms.CenterCrop(size)(transforms.Resize(size)(img))
    return torch.from_numpy(np.asarray(img).copy())


def load_folder_u8(main_dir: str, size: int = 256, workers: int = 4, pin: bool = True) -> Tuple[torch.Tensor, List[str]]:
    """All images of `main_dir` in SynthDataset order as one uint8 [N,size,size,3] tensor (pinned when possible)."""
    files = list_images(main_dir)
    return load_files_u8(files, size, workers, pin), files


def load_files_u8(files: List[str], size: int = 256, workers: int = 4, pin: bool = True) -> torch.Tensor:
    """The given image files (a rank's shard of the SynthDataset order) as one uint8 [N,size,size,3] tensor."""
    out = torch.empty((len(files), size, size, 3), dtype=torch.uint8)
    if pin and torch.cuda.is_available():
        out = out.pin_memory()
    if workers > 1 and len(files) > 64:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(workers) as ex:
            for i, t in enumerate(ex.map(lambda f: load_image_u8(f, size), files)):
                out[i] = t
    else:
        for i, f in enumerate(files):
            out[i] = load_image_u8(f, size)
    return out

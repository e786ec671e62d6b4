"""Host-side logic that needs no GPU: dataset ordering, CLI surface."""
import os

import numpy as np
import pytest

from dcr_b200 import cli, data


def test_natural_order_matches_natsort_semantics(tmp_path):
    names = ["10.png", "2.png", "1.png", "img12.jpg", "img3.jpg", "a/5.png", "a/40.png", "b/1.JPEG", "skip.txt"]
    for n in names:
        p = tmp_path / "gen" / n
        p.parent.mkdir(parents=True, exist_ok=True)
        p.write_bytes(b"x")
    got = [os.path.relpath(f, tmp_path / "gen") for f in data.list_images(str(tmp_path / "gen"))]
    # only leaf folders contribute (diff_retrieval.py:75-77): the files next to sub-folders a/, b/ are skipped
    assert got == ["a/5.png", "a/40.png", "b/1.JPEG"]
    leaf = tmp_path / "leaf"
    leaf.mkdir()
    for n in ["10.png", "2.png", "1.png", "img12.jpg", "img3.jpg"]:
        (leaf / n).write_bytes(b"x")
    assert [os.path.basename(f) for f in data.list_images(str(leaf))] == ["1.png", "2.png", "10.png", "img3.jpg", "img12.jpg"]


def test_load_image_matches_reference_transform(tmp_path):
    import torch
    from PIL import Image
    from torchvision import transforms
    arr = np.random.default_rng(0).integers(0, 256, (300, 420, 3), dtype=np.uint8)
    f = tmp_path / "x.png"
    Image.fromarray(arr).save(f)
    u8 = data.load_image_u8(str(f))
    assert u8.shape == (256, 256, 3)
    ref = transforms.Compose([transforms.Resize(256), transforms.CenterCrop(224), transforms.ToTensor(),
                              transforms.Normalize([0.5] * 3, [0.5] * 3)])(Image.open(f).convert("RGB"))
    from oracle import models as om
    assert torch.equal(om.preprocess(u8[None])[0], ref)      # centre 224 of our 256 crop == reference crop


def test_cli_flag_surface_matches_reference():
    p = cli.build_parser()
    opts = {a for act in p._actions for a in act.option_strings}
    for flag in ["--query_dir", "--val_dir", "--pt_style", "-a", "--arch", "-j", "--workers", "-b", "--batch-size",
                 "--world-size", "--rank", "--dist-url", "--dist-backend", "--seed", "--gpu",
                 "--multiprocessing-distributed", "--multiscale", "--pretrained", "--similarity_metric",
                 "--num_loss_chunks", "--numpatches", "--isvit", "--layer", "--stype", "--keephead", "--keeppredictor",
                 "-ssp", "--sim_save_path", "--einsum_chunks", "--dontsave", "--num_matches", "--imsize", "--noeval"]:
        assert flag in opts, flag
    d = p.parse_args(["--query_dir", "q", "--val_dir", "v"])
    assert (d.pt_style, d.arch, d.similarity_metric, d.workers, d.batch_size, d.dist_backend, d.layer,
            d.einsum_chunks, d.num_matches, d.imsize) == ("sscd", "resnet50", "dotproduct", 4, 128, "nccl", 1, 30, 4, 224)


def test_dense_expansion_of_grouped_conv_weight():
    """nets._dense_from_grouped: the block-diagonal dense weight computes the same convolution as the grouped one."""
    import torch
    import torch.nn.functional as F
    from dcr_b200 import nets
    g = torch.Generator().manual_seed(0)
    w = torch.randn(32, 4, 3, 3, generator=g)          # 8 groups of 4 -> 4
    x = torch.randn(2, 32, 9, 9, generator=g)
    dense = nets._dense_from_grouped(w, 32)
    assert dense.shape == (32, 32, 3, 3) and (dense != 0).sum() == w.numel()
    assert torch.allclose(F.conv2d(x, dense, padding=1), F.conv2d(x, w, padding=1, groups=8), atol=1e-5)
    assert nets._dense_from_grouped(dense, 32) is dense


def test_fid_call_surface_errors_need_no_gpu(tmp_path):
    """metrics/fid.py:239-255 / :258-275 error behaviour is checked before any device work."""
    from dcr_b200 import fid as dfid
    a = tmp_path / "a"
    a.mkdir()
    with pytest.raises(RuntimeError, match="Invalid path"):
        dfid.calculate_fid_given_paths([str(tmp_path / "missing"), str(a)], 50, "cuda", 2048)
    with pytest.raises(KeyError):                      # InceptionV3.BLOCK_INDEX_BY_DIM[dims] (metrics/fid.py:245)
        dfid.calculate_fid_given_paths([str(a), str(a)], 50, "cuda", 100)
    with pytest.raises(RuntimeError, match="Invalid path"):
        dfid.save_fid_stats([str(tmp_path / "missing"), str(tmp_path / "x.npz")], 50, "cuda", 2048)
    (tmp_path / "exists.npz").write_bytes(b"")
    with pytest.raises(RuntimeError, match="Existing output file"):
        dfid.save_fid_stats([str(a), str(tmp_path / "exists.npz")], 50, "cuda", 2048)


def test_embedding_search_command_line_errors():
    from dcr_b200 import embedding_search as es
    with pytest.raises(es._lib.DcrError):
        es.embed_main(["--parquet-fname", "x.parquet"])                      # download path: out of scope
    with pytest.raises(RuntimeError, match="Either tar files or image folder"):
        es.embed_main([])                                                    # embedding_search/utils.py:66
    with pytest.raises(NotImplementedError):
        es.embed_main(["--image-folder", "x", "--arch", "resnet18"])


def test_cli_rejects_unknown_models_and_metrics(tmp_path):
    q = tmp_path / "q"
    q.mkdir()
    with pytest.raises(NotImplementedError):
        cli.main(["--query_dir", str(q), "--val_dir", str(q), "--similarity_metric", "cosine"])
    with pytest.raises(NotImplementedError):
        cli.main(["--query_dir", str(q), "--val_dir", str(q), "--pt_style", "clip", "--arch", "resnet50"])     # RN50x16 tower
    with pytest.raises(NotImplementedError):
        cli.main(["--query_dir", str(q), "--val_dir", str(q), "--pt_style", "vicregl"])
    with pytest.raises(FileNotFoundError):
        cli.main(["--query_dir", str(q), "--val_dir", str(q), "--pt_style", "sscd", "--weights", str(tmp_path / "none.pt")])


def test_synthdataset_caption_json_branch(tmp_path):
    """diff_retrieval.py:64-69, 92-96: no prompts.txt + 'laion' in the path -> file list and prompts come from
    `*combined_captions.json` next to the part of the path before 'train', in the json's key order."""
    import json
    root = tmp_path / "laion_10k"
    d = root / "train" / "imgs"
    d.mkdir(parents=True)
    files = [str(d / n) for n in ("b10.png", "a2.png", "a1.png")]
    with open(root / "x_combined_captions.json", "w") as f:
        json.dump({files[0]: ["cap b10", "alt"], files[1]: ["cap a2"], files[2]: ["cap a1"]}, f)
    imgs, prompts = data.dataset_index(str(d))
    assert imgs == files and prompts == ["cap b10", "cap a2", "cap a1"]
    assert data.list_images(str(d)) == files
    # with prompts.txt present the folder branch wins
    (d / "prompts.txt").write_text("p1\np2\n")
    for n in ("b10.png", "a2.png"):
        (d / n).write_bytes(b"")
    imgs2, prompts2 = data.dataset_index(str(d))
    assert [os.path.basename(p) for p in imgs2] == ["a2.png", "b10.png"] and prompts2 == ["p1\n", "p2\n"]


def test_load_state_dict_reads_torchscript_and_plain_files(tmp_path):
    """The SSCD models are TorchScript files (diff_retrieval.py:277-283: torch.jit.load).  Script an SSCD-shaped module
    (backbone.* ResNet trunk + embeddings.1 head), save it, and read its tensors back through cli.load_state_dict --
    the loading path a real sscd_disc_mixup.torchscript.pt takes; the builder must see the same keys and values as from
    a plain torch.save'd state_dict."""
    import torch
    import torchvision
    from dcr_b200 import cli
    from oracle import models as om

    class GeM(torch.nn.Module):
        def forward(self, x):
            return x.clamp(min=1e-6).pow(3.0).mean(dim=(2, 3)).pow(1.0 / 3.0)

    class SscdLike(torch.nn.Module):
        def __init__(self):
            super().__init__()
            r = torchvision.models.resnet50(weights=None)
            r.fc = torch.nn.Identity()
            r.avgpool = torch.nn.Identity()
            self.backbone = torch.nn.Sequential()
            for name in ("conv1", "bn1", "relu", "maxpool", "layer1", "layer2", "layer3", "layer4"):
                self.backbone.add_module(name, getattr(r, name))
            self.embeddings = torch.nn.Sequential(GeM(), torch.nn.Linear(2048, 512))

        def forward(self, x):
            return torch.nn.functional.normalize(self.embeddings(self.backbone(x)), dim=1)

    sd = om.make_sscd_state_dict(5)
    m = SscdLike()
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all("num_batches_tracked" in k for k in missing)
    m.eval()
    ts_path, pt_path = str(tmp_path / "sscd_like.torchscript.pt"), str(tmp_path / "sscd_like.pth")
    torch.jit.script(m).save(ts_path)
    torch.save(sd, pt_path)
    got_ts, got_pt = cli.load_state_dict(ts_path), cli.load_state_dict(pt_path)
    for k, v in sd.items():
        assert k in got_ts and torch.equal(got_ts[k].cpu(), v), k
        assert torch.equal(got_pt[k], v)
    # and the scripted module itself agrees with the oracle restatement on an input (the only SSCD pin available)
    x = torch.randn(1, 3, 224, 224, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        ref = torch.jit.load(ts_path)(x)
    assert (om.sscd_forward(sd, x) - ref).abs().max().item() < 2e-5

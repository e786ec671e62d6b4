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
    with pytest.raises(NotImplementedError):
        dfid.calculate_fid_given_paths([str(a), str(a)], 50, "cuda", 768)
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
        cli.main(["--query_dir", str(q), "--val_dir", str(q), "--pt_style", "clip"])
    with pytest.raises(FileNotFoundError):
        cli.main(["--query_dir", str(q), "--val_dir", str(q), "--pt_style", "sscd", "--weights", str(tmp_path / "none.pt")])

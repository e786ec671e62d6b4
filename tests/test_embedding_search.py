"""embedding_search front end: pkl format, folder order, key mapping and tie rules on the
CPU; the streamed GPU search against the oracle under -m gpu."""
import os
import pickle as pkl

import numpy as np
import pytest
import torch

from dcr_b200 import embedding_search as es
from oracle import embedding_search as oes


def _unit(x):
    return (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)


def _make_tree(tmp_path, nq=37, d=64, sizes=(50, 1, 0, 80), seed=0, broken=True):
    rng = np.random.default_rng(seed)
    q = _unit(rng.standard_normal((nq, d)))
    root = tmp_path / "laion"
    root.mkdir()
    gal = []
    for i, n in enumerate(sizes):
        g = _unit(rng.standard_normal((max(n, 1), d)))[:n]
        if n >= 10:                       # planted near-copies of some queries, and one exact duplicate pair for the tie rule
            g[3] = _unit(q[i:i + 1] + 0.05 * rng.standard_normal((1, d)))[0]
            g[7] = q[5]
        gal.append(g)
        es.write_embedding_pkl(str(root / f"part_{i:02d}" / "embedding.pkl"), g, [f"k{i}_{j}" for j in range(n)])
    if broken:                            # unreadable folder: printed and skipped, similarity_search.py:54-56
        (root / "part_zz").mkdir()
        (root / "part_zz" / "embedding.pkl").write_bytes(b"not a pickle")
        (root / "stray_file.txt").write_text("ignored")
    gen_path = tmp_path / "gen" / "embedding.pkl"
    es.write_embedding_pkl(str(gen_path), q, [f"{j}.png" for j in range(nq)])
    return str(root), str(gen_path), q, gal


def test_pkl_round_trip(tmp_path):
    f = np.arange(12, dtype=np.float64).reshape(4, 3)
    p = str(tmp_path / "a" / "embedding.pkl")
    es.write_embedding_pkl(p, torch.from_numpy(f), ["a", "b", "c", "d"])
    with open(p, "rb") as fh:
        raw = pkl.load(fh)
    assert set(raw) == {"features", "indexes"} and raw["features"].dtype == np.float32   # download_and_..._embedding.py:93
    feats, keys = es.read_embedding_pkl(p)
    assert feats.dtype == np.float32 and np.array_equal(feats, f.astype(np.float32)) and keys == ["a", "b", "c", "d"]
    with pytest.raises(es._lib.DcrError):
        es.write_embedding_pkl(p, f, ["only-one"])


def test_image_file_order(tmp_path):
    for n in ["10.png", "9.png", "a.jpg", "b.JPG", "c.jpeg", "2.png"]:
        (tmp_path / n).write_bytes(b"")
    # plain string sort of .png/.jpg names (embedding_search/utils.py:119-123), not the natsort of diff_retrieval.py
    assert es.list_image_files(str(tmp_path)) == ["10.png", "2.png", "9.png", "a.jpg"]


def test_merge_rule_and_keys():
    best_s = torch.tensor([-1.0, 0.5, 0.5, 0.2])
    best_f = torch.tensor([-1, 0, 0, 0])
    best_r = torch.tensor([0, 4, 4, 4])
    cur_s = torch.tensor([-1.0, 0.5, 0.6, float("nan")])
    es.merge_folder_best(best_s, best_f, best_r, cur_s, torch.tensor([9, 9, 9, 9]), 1)
    assert best_f.tolist() == [-1, 0, 1, 0] and best_r.tolist() == [0, 4, 9, 4]        # ties and NaN keep the earlier folder
    keys = es.keys_from_matches(["f0", "f1"], [["a"] * 5, ["b"] * 10], best_f.numpy(), best_r.numpy())
    assert keys.tolist() == ["0.0", "f0:a", "f1:b", "f0:a"]


def test_oracle_matches_brute_force(tmp_path):
    root, gen, q, gal = _make_tree(tmp_path)
    out = oes.similarity_search(root, gen, num_chunks=5)
    names = ["part_00", "part_01", "part_02", "part_03"]
    allg = np.concatenate(gal).astype(np.float64)
    owner = np.concatenate([np.full(len(g), i) for i, g in enumerate(gal)])
    local = np.concatenate([np.arange(len(g)) for g in gal])
    S = q.astype(np.float64) @ allg.T
    best = S.argmax(axis=1)                         # first maximum in (folder, row) order == the reference's merge order
    exp_keys = [f"{names[owner[b]]}:k{owner[b]}_{local[b]}" for b in best]
    assert out["keys"].tolist() == exp_keys
    assert out["scores"].dtype == np.float64
    np.testing.assert_allclose(out["scores"], S[np.arange(len(best)), best], rtol=0, atol=1e-6)
    assert out["gen_images"] == [f"{j}.png" for j in range(q.shape[0])]
    # chunking of the queries does not change anything
    out1 = oes.similarity_search(root, gen, num_chunks=1)
    assert out1["keys"].tolist() == out["keys"].tolist() and np.array_equal(out1["scores"], out["scores"])
    # query 5 is duplicated exactly in part_00 and part_03: the earlier folder wins
    assert out["keys"][5] == "part_00:k0_7"


def test_oracle_unmatched_and_chunks(tmp_path):
    root = tmp_path / "empty"
    root.mkdir()
    gen = tmp_path / "gen.pkl"
    es.write_embedding_pkl(str(gen), np.ones((3, 4), np.float32), ["a", "b", "c"])
    out = oes.similarity_search(str(root), str(gen))
    assert out["scores"].tolist() == [-1.0, -1.0, -1.0] and out["keys"].tolist() == ["0.0"] * 3     # :47-48,:71
    assert oes.torch_chunk_bounds(10, 4) == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert [b - a for a, b in oes.torch_chunk_bounds(7, 100)] == [1] * 7


def test_parsers_match_reference_flags():
    a = es.build_search_parser().parse_args(["--laion-embedding-folder", "x", "--generation-embedding-path", "y",
                                             "--dump-path", "z"])
    assert a.num_chunks == 100
    b = es.build_embed_parser().parse_args([])
    assert (b.pt_style, b.arch, b.batch_size, b.workers, b.gpu, b.similarity_metric) == ("sscd", "resnet50", 128, 8, 0, "d")


@pytest.mark.gpu
def test_streamed_search_matches_oracle(tmp_path):
    root, gen, q, gal = _make_tree(tmp_path, nq=300, d=512, sizes=(3000, 1, 0, 5000, 257), seed=3)
    dump = str(tmp_path / "out" / "result.pkl")
    out = es.similarity_search(root, gen, dump, verbose=False)
    ref = oes.similarity_search(root, gen, num_chunks=7)
    assert out["keys"].tolist() == ref["keys"].tolist()
    assert out["scores"].dtype == np.float64
    np.testing.assert_allclose(out["scores"], ref["scores"], rtol=0, atol=1e-6)
    with open(dump, "rb") as f:
        saved = pkl.load(f)
    assert saved["keys"].tolist() == out["keys"].tolist() and saved["gen_images"] == out["gen_images"]


@pytest.mark.gpu
def test_streamed_search_no_folders(tmp_path):
    root = tmp_path / "empty"
    root.mkdir()
    gen = tmp_path / "gen.pkl"
    es.write_embedding_pkl(str(gen), np.ones((3, 8), np.float32), ["a", "b", "c"])
    out = es.similarity_search(str(root), str(gen), verbose=False)
    assert out["scores"].tolist() == [-1.0] * 3 and out["keys"].tolist() == ["0.0"] * 3


@pytest.mark.gpu
def test_generate_embeddings_imagenet_norm(tmp_path):
    """a13: flat folder, sorted names, ImageNet mean/std, descriptors as the model returns them."""
    from PIL import Image
    from dcr_b200 import nets
    from oracle import models as omodels
    rng = np.random.default_rng(0)
    folder = tmp_path / "imgs"
    folder.mkdir()
    names = ["b.png", "a.png", "10.png", "9.png"]
    for n in names:
        Image.fromarray(rng.integers(0, 256, (256, 256, 3), dtype=np.uint8)).save(folder / n)
    sd = omodels.make_sscd_state_dict(1)
    net = nets.build_sscd_resnet50(sd, max_batch=4, mean=es.IMAGENET_MEAN, std=es.IMAGENET_STD)
    feats, keys = es.generate_embeddings(net, str(folder), str(tmp_path / "dump"), batch_size=4, workers=1)
    assert keys == sorted(names)
    saved, saved_keys = es.read_embedding_pkl(str(tmp_path / "dump" / "embedding.pkl"))
    assert saved_keys == keys and np.array_equal(saved, feats.cpu().numpy())
    imgs = torch.from_numpy(np.stack([np.asarray(Image.open(folder / n).convert("RGB")) for n in keys]))
    ref = omodels.sscd_forward(sd, omodels.preprocess(imgs, es.IMAGENET_MEAN, es.IMAGENET_STD), bf16_points=True).numpy()
    got = feats.cpu().numpy()
    cos = (got * ref).sum(1) / (np.linalg.norm(got, axis=1) * np.linalg.norm(ref, axis=1))
    assert cos.min() > 0.999, cos


def test_webdataset_tar_reader(tmp_path):
    """embedding_search/utils.py:52-62: samples of a webdataset shard = members sharing a basename; the image is the `jpg`
    member, the index the `key` field of the `json` member; 'x/{000000..000001}.tar' expands to the shard list."""
    import io
    import json
    import tarfile

    import numpy as np
    from PIL import Image
    from dcr_b200 import embedding_search as es
    rng = np.random.default_rng(0)
    want = []
    for shard in range(2):
        with tarfile.open(tmp_path / f"{shard:06d}.tar", "w") as tar:
            for i in range(3):
                base = f"{shard:05d}{i:04d}"
                arr = rng.integers(0, 255, (300, 280, 3), dtype=np.uint8)
                buf = io.BytesIO()
                Image.fromarray(arr).save(buf, format="JPEG")
                for ext, payload in (("jpg", buf.getvalue()), ("json", json.dumps({"key": base, "url": "u"}).encode()),
                                     ("txt", b"caption")):
                    info = tarfile.TarInfo(f"{base}.{ext}")
                    info.size = len(payload)
                    tar.addfile(info, io.BytesIO(payload))
                want.append(base)
    urls = es.expand_tar_urls([str(tmp_path) + "/{000000..000001}.tar"])
    assert urls == [str(tmp_path / "000000.tar"), str(tmp_path / "000001.tar")]
    got = list(es.iter_tar_samples(urls))
    assert [k for _, k in got] == want
    assert all(tuple(img.shape) == (256, 256, 3) and img.dtype == torch.uint8 for img, _ in got)

"""CLIP score (utils_ret.py:1046-1066): oracle vs transformers.CLIPModel, tokenizer vs transformers.CLIPTokenizer (CPU);
CUDA towers vs the oracle (GPU)."""
import json
import os
from collections import Counter

import numpy as np
import pytest
import torch

from oracle import clip as oc

CORPUS = ("a photo of a cat sitting on the mat . the quick brown fox jumps over the lazy dog ! "
          "an astronaut riding a horse in photorealistic style , 4k trending on artstation ; don't you'll it's "
          "stable diffusion generated image of a castle at sunset with dramatic lighting and 12345 67 numbers").split()


def _train_merges(words, n):
    """A tiny BPE trainer: enough merges to exercise multi-level merging."""
    from dcr_b200.clipscore import _byte_symbols
    sym = _byte_symbols()
    vocab = Counter()
    for w in words:
        s = [sym[b] for b in w.lower().encode("utf-8")]
        s[-1] += "</w>"
        vocab[tuple(s)] += 1
    merges = []
    for _ in range(n):
        pairs = Counter()
        for w, c in vocab.items():
            for i in range(len(w) - 1):
                pairs[(w[i], w[i + 1])] += c
        if not pairs:
            break
        best = max(sorted(pairs), key=lambda p: pairs[p])
        merges.append(best)
        new = Counter()
        for w, c in vocab.items():
            out, i = [], 0
            while i < len(w):
                if i + 1 < len(w) and (w[i], w[i + 1]) == best:
                    out.append(w[i] + w[i + 1])
                    i += 2
                else:
                    out.append(w[i])
                    i += 1
            new[tuple(out)] += c
        vocab = new
    return merges


def test_tokenizer_matches_transformers_clip_tokenizer(tmp_path):
    from transformers import CLIPTokenizer
    from dcr_b200.clipscore import SimpleTokenizer
    merges = _train_merges(CORPUS, 120)
    mpath = tmp_path / "merges.txt"
    mpath.write_text("#version: 0.2\n" + "\n".join(" ".join(m) for m in merges) + "\n", encoding="utf-8")
    ours = SimpleTokenizer(str(mpath), n_merges=len(merges))
    assert len(ours.encoder) == 512 + len(merges) + 2
    vpath = tmp_path / "vocab.json"
    vpath.write_text(json.dumps(ours.encoder), encoding="utf-8")
    hf = CLIPTokenizer(str(vpath), str(mpath))
    texts = ["A photo of a cat sitting on the mat.", "the quick brown fox, jumps over the lazy dog!",
             "don't you'll it's 4k 12345", "  stable   diffusion castle at sunset  ", " ".join(CORPUS * 3)]
    got = ours.tokenize(texts, 77, True)
    want = hf(texts, padding="max_length", max_length=77, truncation=True)["input_ids"]
    eot = ours.encoder["<|endoftext|>"]
    for g, w in zip(got.tolist(), want):
        n = g.index(eot) + 1
        assert g[:n] == w[:n], (g[:n], w[:n])
        assert all(v == 0 for v in g[n:])                       # clip.tokenize pads with zeros
    assert got[4, 76].item() == eot                               # truncated: the last position is the end token
    with pytest.raises(RuntimeError):
        ours.tokenize([" ".join(CORPUS * 3)], 77, False)


def _hf_model(sd, layers):
    from transformers import CLIPConfig, CLIPModel
    cfg = CLIPConfig(text_config=dict(hidden_size=512, intermediate_size=2048, num_hidden_layers=layers, num_attention_heads=8,
                                      max_position_embeddings=77, vocab_size=49408, hidden_act="quick_gelu", eos_token_id=49407,
                                      bos_token_id=49406, pad_token_id=0, layer_norm_eps=1e-5),
                     vision_config=dict(hidden_size=768, intermediate_size=3072, num_hidden_layers=layers, num_attention_heads=12,
                                        image_size=224, patch_size=16, hidden_act="quick_gelu", layer_norm_eps=1e-5),
                     projection_dim=512)
    m = CLIPModel(cfg).eval()
    missing, unexpected = m.load_state_dict(oc.to_hf_state_dict(sd), strict=False)
    assert not missing and not unexpected
    return m


def _tokens():
    tok = torch.zeros(3, 77, dtype=torch.long)
    tok[:, 0] = 49406
    tok[0, 1:6] = torch.tensor([320, 1125, 539, 320, 2368]); tok[0, 6] = 49407
    tok[1, 1:3] = torch.tensor([1000, 2000]); tok[1, 3] = 49407
    tok[2, 1:76] = torch.arange(500, 575); tok[2, 76] = 49407
    return tok


def test_clip_oracle_matches_transformers_implementation():
    """The `clip` package is absent (PARITY UNPINNED against it); the restated architecture is held against the independent
    implementation of the same model in transformers, with the weights converted name by name."""
    sd = oc.make_clip_state_dict(0, layers=2)
    m = _hf_model(sd, 2)
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    tok = _tokens()
    with torch.no_grad():
        hi, ht = m.get_image_features(pixel_values=x), m.get_text_features(input_ids=tok)
    hi = hi if torch.is_tensor(hi) else hi.pooler_output
    ht = ht if torch.is_tensor(ht) else ht.pooler_output
    assert (oc.encode_image(sd, x) - hi).abs().max().item() < 2e-5
    assert (oc.encode_text(sd, tok) - ht).abs().max().item() < 2e-5


@pytest.mark.gpu
def test_clip_towers_and_score_on_gpu():
    from dcr_b200 import clipscore, synthetic
    from oracle import models as om
    sd = oc.make_clip_state_dict(1, layers=2)
    img = synthetic.images(3, seed=71)
    x = om.preprocess(img)
    tok = _tokens()
    ref_i, ref_t = oc.encode_image(sd, x), oc.encode_text(sd, tok)
    sc = clipscore.ClipScorer(sd, precision="exact", max_batch=2)
    got_i, got_t = sc.encode_image(img.cuda()).cpu(), sc.encode_text(tok.cuda()).cpu()
    assert (got_i - ref_i).abs().max().item() < 5e-5 * max(1.0, ref_i.abs().max().item())
    assert (got_t - ref_t).abs().max().item() < 5e-5 * max(1.0, ref_t.abs().max().item())
    want = oc.clip_score(ref_i, ref_t)
    got = float(sc.pair_scores(img.cuda(), tok).double().mean())
    assert abs(got - want) < 1e-4
    fast = clipscore.ClipScorer(sd, precision="fast", max_batch=4)
    fq_i, fq_t = oc.encode_image(sd, x, bf16_points=True), oc.encode_text(sd, tok, bf16_points=True)
    fi, ft = fast.encode_image(img.cuda()).cpu(), fast.encode_text(tok.cuda()).cpu()
    assert (fi - fq_i).abs().max().item() < 6e-2 * max(1.0, fq_i.abs().max().item())
    assert (ft - fq_t).abs().max().item() < 6e-2 * max(1.0, fq_t.abs().max().item())


@pytest.mark.gpu
def test_clip_vit_l14_geometry_through_the_cli_builder():
    """--pt_style clip --arch vit_large (diff_retrieval.py:264-271: ViT-L/14): patch 14 -> 257 tokens (the streamed-KV
    attention kernel), width 1024 / 16 heads; a 2-layer seeded model against the oracle."""
    from dcr_b200 import nets, synthetic
    from oracle import models as om
    sd = oc.make_clip_state_dict(3, layers=2, vision_width=1024, embed_dim=768, patch=14)
    img = synthetic.images(2, seed=72)
    ref = oc.encode_image(sd, om.preprocess(img))
    net = nets.build_clip_visual(sd, max_batch=2, precision="exact")
    assert net.tokens == 257
    got = net(img.cuda()).cpu()
    assert (got - ref).abs().max().item() < 5e-5 * max(1.0, ref.abs().max().item())
    fast = nets.build_clip_visual(sd, max_batch=2, precision="fast")(img.cuda()).cpu()
    refq = oc.encode_image(sd, om.preprocess(img), bf16_points=True)
    assert (fast - refq).abs().max().item() < 6e-2 * max(1.0, refq.abs().max().item())

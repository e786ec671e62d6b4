"""Generates the committed golden vectors by running the REFERENCE's own modules (importable parts only) in the
build container:  python tests/golden/make_golden.py        (needs /root/reference; not run on the GPU box)

  dino_vits16_seed{S}.npz : dino_vits.VisionTransformer (vit_small, patch 16) from /root/reference/dino_vits.py with
                            the seeded state_dict of oracle.models.make_vit_state_dict(S) on seeded inputs.
  fid_inception_seed{S}.npz: metrics.inception.InceptionV3 (FID variant) with seeded weights  (added with the FID row)
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def golden_inputs(seed, n=2, size=224):
    return torch.randn(n, 3, size, size, generator=torch.Generator().manual_seed(7000 + seed))


def make_dino(seed):
    from oracle.models import make_vit_state_dict
    dv = _load("ref_dino_vits", os.path.join(REF, "dino_vits.py"))
    model = dv.vit_small(patch_size=16, num_classes=0)       # what dino_vits16 builds (dino_vits.py:346)
    sd = make_vit_state_dict(seed)
    missing, unexpected = model.load_state_dict(sd, strict=True)
    model.eval()
    x = golden_inputs(seed)
    with torch.no_grad():
        y = model(x)
    np.savez_compressed(os.path.join(HERE, f"dino_vits16_seed{seed}.npz"), seed=seed, out=y.numpy(),
                        in_checksum=float(x.double().sum()))
    print("dino", seed, y.shape, float(y.abs().mean()))


def make_dino_other_size(seed, size=160):
    """Same module at a 160 x 160 input: exercises interpolate_pos_encoding (dino_vits.py:213-233)."""
    from oracle.models import make_vit_state_dict
    dv = _load("ref_dino_vits", os.path.join(REF, "dino_vits.py"))
    model = dv.vit_small(patch_size=16, num_classes=0)
    model.load_state_dict(make_vit_state_dict(seed), strict=True)
    model.eval()
    x = golden_inputs(seed, size=size)
    with torch.no_grad():
        y = model(x)
        pos = model.interpolate_pos_encoding(torch.zeros(1, 1 + (size // 16) ** 2, 384), size, size)
    np.savez_compressed(os.path.join(HERE, f"dino_vits16_seed{seed}_{size}.npz"), seed=seed, out=y.numpy(),
                        pos_embed=pos.numpy(), in_checksum=float(x.double().sum()))
    print("dino", seed, size, y.shape, pos.shape)


def make_dino_variants(seed):
    """The forward variants diff_retrieval.py selects: global_pool='' (splitloss on a ViT, :258-263 -> every normed
    token, dino_vits.py:255-256) and get_intermediate_layers(x, n)[0] (--layer n, utils_ret.py:732,745).  Only a
    checksum-like subsample of the [2, 197, 384] token tensor is stored to keep the fixture small."""
    from oracle.models import make_vit_state_dict
    dv = _load("ref_dino_vits", os.path.join(REF, "dino_vits.py"))
    sd = make_vit_state_dict(seed)
    x = golden_inputs(seed)
    m_tok = dv.vit_small(patch_size=16, num_classes=0, global_pool="")
    m_tok.load_state_dict(sd, strict=True)
    m_tok.eval()
    m_cls = dv.vit_small(patch_size=16, num_classes=0)
    m_cls.load_state_dict(sd, strict=True)
    m_cls.eval()
    with torch.no_grad():
        tokens = m_tok(x)                                            # [2, 197, 384]
        inter3 = m_cls.get_intermediate_layers(x, 3)[0]              # normed output of block depth-3
    np.savez_compressed(os.path.join(HERE, f"dino_vits16_seed{seed}_variants.npz"), seed=seed,
                        tokens_rows=tokens[:, ::14, :].numpy(), tokens_sum=tokens.double().sum(dim=-1).numpy(),
                        layer3_cls=inter3[:, 0, :].numpy(), layer3_tokens_sum=inter3.double().sum(dim=-1).numpy(),
                        in_checksum=float(x.double().sum()))
    print("dino variants", seed, tokens.shape, inter3.shape)


def make_fid_inception(seed):
    """metrics/inception.InceptionV3([3]) exactly as metrics/fid.py:245-247 builds it, with the URL weight load
    (inception.py:219) replaced by the seeded state_dict of oracle.models.make_inception_state_dict(seed)."""
    from oracle.models import make_inception_state_dict, fid_preprocess
    sys.path.insert(0, REF)
    inc = _load("ref_inception", os.path.join(REF, "metrics", "inception.py"))
    sd = make_inception_state_dict(seed)
    full = dict(sd)

    def fake_loader(url, progress=True):
        import torchvision
        m = torchvision.models.inception_v3(weights=None, aux_logits=False, init_weights=False, num_classes=1008)
        base = m.state_dict()
        base.update(full)
        return base

    inc.load_state_dict_from_url = fake_loader
    model = inc.InceptionV3([3], resize_input=True, normalize_input=True)
    model.eval()
    img = torch.randint(0, 256, (2, 299, 299, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(8000 + seed))
    with torch.no_grad():
        y = model(fid_preprocess(img))[0].squeeze(-1).squeeze(-1)
    np.savez_compressed(os.path.join(HERE, f"fid_inception_seed{seed}.npz"), seed=seed, out=y.numpy())
    print("inception", seed, y.shape, float(y.abs().mean()))


def make_ipr(seed):
    """metrics/ipr.py's own numpy functions (compute_pairwise_distances, distances2radii, compute_metric, realism) on
    seeded feature sets; the VGG-16 extractor needs CUDA in the reference (ipr.py:139 `.cuda()`) and is not run."""
    sys.path.insert(0, REF)
    ipr = _load("ref_ipr", os.path.join(REF, "metrics", "ipr.py"))
    rng = np.random.default_rng(900 + seed)
    ref = (rng.standard_normal((300, 64)) * 3 + 1).astype(np.float32)
    sub = (rng.standard_normal((200, 64)) * 3.3 + 1.2).astype(np.float32)
    man_ref = ipr.Manifold(ref, ipr.distances2radii(ipr.compute_pairwise_distances(ref), k=3))
    man_sub = ipr.Manifold(sub, ipr.distances2radii(ipr.compute_pairwise_distances(sub), k=3))
    precision = ipr.compute_metric(man_ref, sub)
    recall = ipr.compute_metric(man_sub, ref)
    real = np.array([ipr.realism(man_ref, sub[i:i + 1]) for i in range(8)])
    np.savez_compressed(os.path.join(HERE, f"ipr_seed{seed}.npz"), seed=seed, radii_ref=man_ref.radii, radii_sub=man_sub.radii,
                        precision=precision, recall=recall, realism=real)
    print("ipr", seed, precision, recall, real[:3])


if __name__ == "__main__":
    torch.manual_seed(0)
    for s in (0, 1):
        make_dino(s)
    make_dino_other_size(0)
    make_dino_variants(0)
    make_fid_inception(0)
    make_ipr(0)

"""Oracle for CLIP score (utils_ret.py:1046-1066) -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference calls the un-vendored `clip` package (openai/CLIP, unpinned: `import clip` utils_ret.py:1045, model
"ViT-B/16" :1048).  Its published architecture is restated here functionally over a state_dict with the package's own
parameter names (what `clip.load(...)[0].state_dict()` yields):

  encode_image   clip/model.py VisionTransformer.forward: conv1 (patch 16, no bias) -> [class_embedding | patches] +
                 positional_embedding -> ln_pre -> 12 ResidualAttentionBlocks (LayerNorm eps 1e-5, nn.MultiheadAttention
                 with packed in_proj, QuickGELU MLP) -> ln_post(x[:, 0]) @ proj
  encode_text    clip/model.py CLIP.encode_text: token_embedding + positional_embedding -> 12 blocks with the causal
                 mask -> ln_final -> x[arange, text.argmax(-1)] @ text_projection
  clip_score     utils_ret.py:1055-1066: mean over pairs of <normalised image feature, normalised text feature>
PARITY UNPINNED against the package itself (absent here); pinned instead against the independent implementation of the
same architecture in `transformers.CLIPModel` (tests/test_clip.py converts the weights and compares).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F


def _r(x: torch.Tensor, on: bool) -> torch.Tensor:
    return x.to(torch.bfloat16).float() if on else x


def _blocks(sd: Dict[str, torch.Tensor], prefix: str, x: torch.Tensor, causal: bool, q: bool) -> torch.Tensor:
    n, t, dim = x.shape
    heads = dim // 64
    layers = 1 + max(int(k[len(prefix):].split(".")[1]) for k in sd if k.startswith(prefix + "resblocks."))
    mask = torch.full((t, t), float("-inf")).triu_(1) if causal else None        # CLIP.build_attention_mask
    for i in range(layers):
        p = f"{prefix}resblocks.{i}."
        y = _r(F.layer_norm(x, (dim,), sd[p + "ln_1.weight"].float(), sd[p + "ln_1.bias"].float(), 1e-5), q)
        qkv = _r(F.linear(y, _r(sd[p + "attn.in_proj_weight"].float(), q), sd[p + "attn.in_proj_bias"].float()), q)
        qkv = qkv.reshape(n, t, 3, heads, 64).permute(2, 0, 3, 1, 4)
        att = (qkv[0] @ qkv[1].transpose(-2, -1)) * 64 ** -0.5
        if mask is not None:
            att = att + mask
        y = _r((att.softmax(dim=-1) @ qkv[2]).transpose(1, 2).reshape(n, t, dim), q)
        x = _r(x + F.linear(y, _r(sd[p + "attn.out_proj.weight"].float(), q), sd[p + "attn.out_proj.bias"].float()), q)
        y = _r(F.layer_norm(x, (dim,), sd[p + "ln_2.weight"].float(), sd[p + "ln_2.bias"].float(), 1e-5), q)
        h = F.linear(y, _r(sd[p + "mlp.c_fc.weight"].float(), q), sd[p + "mlp.c_fc.bias"].float())
        h = _r(h * torch.sigmoid(1.702 * h), q)                                       # QuickGELU
        x = _r(x + F.linear(h, _r(sd[p + "mlp.c_proj.weight"].float(), q), sd[p + "mlp.c_proj.bias"].float()), q)
    return x


@torch.no_grad()
def encode_image(sd: Dict[str, torch.Tensor], x: torch.Tensor, bf16_points: bool = False) -> torch.Tensor:
    """x: float32 [N,3,224,224] (whatever the loader produced; gen_clipscore feeds the 0.5/0.5-normalised tensors)."""
    q = bf16_points
    w = sd["visual.conv1.weight"].float()
    patch = w.shape[-1]
    t = F.conv2d(_r(x, q), _r(w, q), stride=patch)
    n, dim = t.shape[0], t.shape[1]
    t = _r(t.reshape(n, dim, -1).permute(0, 2, 1), q)
    cls = sd["visual.class_embedding"].float().reshape(1, 1, dim).expand(n, -1, -1)
    t = _r(torch.cat([cls, t], dim=1) + sd["visual.positional_embedding"].float(), q)
    t = _r(F.layer_norm(t, (dim,), sd["visual.ln_pre.weight"].float(), sd["visual.ln_pre.bias"].float(), 1e-5), q)
    t = _blocks(sd, "visual.transformer.", t, False, q)
    t = F.layer_norm(t[:, 0], (dim,), sd["visual.ln_post.weight"].float(), sd["visual.ln_post.bias"].float(), 1e-5)
    return _r(t, q) @ _r(sd["visual.proj"].float(), q)


@torch.no_grad()
def encode_text(sd: Dict[str, torch.Tensor], tokens: torch.Tensor, bf16_points: bool = False) -> torch.Tensor:
    """tokens: int64 [N,77] from clip.tokenize (start token, BPE ids, end token 49407 = the arg-max id, zero padding)."""
    q = bf16_points
    x = _r(sd["token_embedding.weight"].float()[tokens] + sd["positional_embedding"].float(), q)
    dim = x.shape[-1]
    x = _blocks(sd, "transformer.", x, True, q)
    x = F.layer_norm(x, (dim,), sd["ln_final.weight"].float(), sd["ln_final.bias"].float(), 1e-5)
    x = x[torch.arange(x.shape[0]), tokens.argmax(dim=-1)]
    return _r(x, q) @ _r(sd["text_projection"].float(), q)


def clip_score(img_feat: torch.Tensor, txt_feat: torch.Tensor) -> float:
    a = img_feat / torch.linalg.norm(img_feat, dim=-1, keepdim=True)          # utils_ret.py:1058-1061
    b = txt_feat / torch.linalg.norm(txt_feat, dim=-1, keepdim=True)
    return float((a * b).sum(dim=-1).double().mean())                          # :1062-1066 (np.mean of the list)


def make_clip_state_dict(seed: int = 0, layers: int = 2, vision_width: int = 768, text_width: int = 512,
                         embed_dim: int = 512, vocab: int = 49408, patch: int = 16, ctx: int = 77) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s, std=1.0: torch.randn(*s, generator=g) * std
    sd = {"visual.conv1.weight": rn(vision_width, 3, patch, patch, std=(1.0 / (3 * patch * patch)) ** 0.5),
          "visual.class_embedding": rn(vision_width, std=0.5),
          "visual.positional_embedding": rn((224 // patch) ** 2 + 1, vision_width, std=0.3),
          "visual.ln_pre.weight": 1 + rn(vision_width, std=0.1), "visual.ln_pre.bias": rn(vision_width, std=0.1),
          "visual.ln_post.weight": 1 + rn(vision_width, std=0.1), "visual.ln_post.bias": rn(vision_width, std=0.1),
          "visual.proj": rn(vision_width, embed_dim, std=vision_width ** -0.5),
          "token_embedding.weight": rn(vocab, text_width, std=0.5), "positional_embedding": rn(ctx, text_width, std=0.3),
          "ln_final.weight": 1 + rn(text_width, std=0.1), "ln_final.bias": rn(text_width, std=0.1),
          "text_projection": rn(text_width, embed_dim, std=text_width ** -0.5), "logit_scale": torch.tensor(4.6052)}
    for prefix, w in (("visual.transformer.", vision_width), ("transformer.", text_width)):
        for i in range(layers):
            p = f"{prefix}resblocks.{i}."
            sd[p + "ln_1.weight"], sd[p + "ln_1.bias"] = 1 + rn(w, std=0.1), rn(w, std=0.1)
            sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"] = rn(3 * w, w, std=1.5 * w ** -0.5), rn(3 * w, std=0.1)
            sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"] = rn(w, w, std=0.5 * w ** -0.5), rn(w, std=0.05)
            sd[p + "ln_2.weight"], sd[p + "ln_2.bias"] = 1 + rn(w, std=0.1), rn(w, std=0.1)
            sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"] = rn(4 * w, w, std=w ** -0.5), rn(4 * w, std=0.1)
            sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"] = rn(w, 4 * w, std=0.5 * (4 * w) ** -0.5), rn(w, std=0.05)
    return sd


def to_hf_state_dict(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """The same weights under transformers.CLIPModel's names (q/k/v projections unpacked from in_proj)."""
    out = {"logit_scale": sd["logit_scale"],
           "text_model.embeddings.token_embedding.weight": sd["token_embedding.weight"],
           "text_model.embeddings.position_embedding.weight": sd["positional_embedding"],
           "text_model.final_layer_norm.weight": sd["ln_final.weight"], "text_model.final_layer_norm.bias": sd["ln_final.bias"],
           "vision_model.embeddings.class_embedding": sd["visual.class_embedding"],
           "vision_model.embeddings.patch_embedding.weight": sd["visual.conv1.weight"],
           "vision_model.embeddings.position_embedding.weight": sd["visual.positional_embedding"],
           "vision_model.pre_layrnorm.weight": sd["visual.ln_pre.weight"], "vision_model.pre_layrnorm.bias": sd["visual.ln_pre.bias"],
           "vision_model.post_layernorm.weight": sd["visual.ln_post.weight"], "vision_model.post_layernorm.bias": sd["visual.ln_post.bias"],
           "visual_projection.weight": sd["visual.proj"].T.contiguous(), "text_projection.weight": sd["text_projection"].T.contiguous()}
    for src, dst in (("visual.transformer.", "vision_model.encoder.layers."), ("transformer.", "text_model.encoder.layers.")):
        i = 0
        while f"{src}resblocks.{i}.ln_1.weight" in sd:
            p, d = f"{src}resblocks.{i}.", f"{dst}{i}."
            w, b = sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"]
            dim = w.shape[1]
            for j, name in enumerate(("q_proj", "k_proj", "v_proj")):
                out[d + f"self_attn.{name}.weight"] = w[j * dim:(j + 1) * dim].clone()
                out[d + f"self_attn.{name}.bias"] = b[j * dim:(j + 1) * dim].clone()
            out[d + "self_attn.out_proj.weight"], out[d + "self_attn.out_proj.bias"] = sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"]
            out[d + "layer_norm1.weight"], out[d + "layer_norm1.bias"] = sd[p + "ln_1.weight"], sd[p + "ln_1.bias"]
            out[d + "layer_norm2.weight"], out[d + "layer_norm2.bias"] = sd[p + "ln_2.weight"], sd[p + "ln_2.bias"]
            out[d + "mlp.fc1.weight"], out[d + "mlp.fc1.bias"] = sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"]
            out[d + "mlp.fc2.weight"], out[d + "mlp.fc2.bias"] = sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"]
            i += 1
    return out

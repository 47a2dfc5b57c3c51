"""ORACLE (test infrastructure only — never imported by the product path).

CPU/PyTorch restatement of the CLIP model that the reference imports as
``src.models.CLIPs.clip_hba.clip`` (reference call sites:
Training/functions/new_cvpr_train_behavior_things_pipeline.py:21, 251-265, 282, 298).

PARITY UNPINNED for the tower arithmetic: the module lives in the un-vendored,
un-pinned repo ``stephenczhao/CLIP-HBA-Official`` (named only at reference
README.md:247; no requirements/lock file).  What is restated here is the published
OpenAI-CLIP ViT architecture (``VisionTransformer`` / ``ResidualAttentionBlock`` with
``nn.MultiheadAttention`` + QuickGELU MLP / text ``Transformer`` with a causal float
mask / cosine logits scaled by ``exp(logit_scale)``), anchored on the reference's
own call sites:
  * module paths ``visual.transformer.resblocks[i].attn.out_proj`` and
    ``transformer.resblocks[i].attn.out_proj`` (NEW:496-498, 508, 666-668);
  * ``clip_model(image, tokens[66,77], pos_embedding) -> [B,66]`` (NEW:298-300);
  * widths 1024 / 768 and rank 32 => 183,040 trainable parameters (RUNLOG:57);
  * ``nn.MultiheadAttention`` seq-first, so ``out_proj.weight``/``.bias`` are read as
    tensors (torch/nn/modules/activation.py:1504-1505), which is how the reference's
    DoRALayer.weight property reaches the graph.
An independent cross-check against ``transformers.CLIPModel`` (weights copied) lives in
tests/test_oracle_cpu.py.
"""
from __future__ import annotations

import hashlib
import math
import os
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

SOT_TOKEN = 49406
EOT_TOKEN = 49407
VOCAB_SIZE = 49408
CONTEXT_LENGTH = 77

# name -> "url" (the reference only passes the value to _download, NEW:252-253)
_MODELS = {
    "ViT-L/14": "synthetic://ViT-L-14.pt",
    "ViT-B/16": "synthetic://ViT-B-16.pt",
    "ViT-tiny/14": "synthetic://ViT-tiny-14.pt",
}

# architecture hyper-parameters of the published checkpoints (+ a tiny test config)
ARCH = {
    "ViT-L/14": dict(embed_dim=768, image_resolution=224, vision_layers=24, vision_width=1024,
                     vision_patch_size=14, context_length=77, vocab_size=VOCAB_SIZE,
                     transformer_width=768, transformer_heads=12, transformer_layers=12),
    "ViT-B/16": dict(embed_dim=512, image_resolution=224, vision_layers=12, vision_width=768,
                     vision_patch_size=16, context_length=77, vocab_size=VOCAB_SIZE,
                     transformer_width=512, transformer_heads=8, transformer_layers=12),
    # small config with the same structure (3 vision / 2 text blocks) for fast parity tests
    "ViT-tiny/14": dict(embed_dim=128, image_resolution=224, vision_layers=3, vision_width=256,
                        vision_patch_size=14, context_length=77, vocab_size=VOCAB_SIZE,
                        transformer_width=128, transformer_heads=2, transformer_layers=2),
}


class QuickGELU(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


class ResidualAttentionBlock(nn.Module):
    def __init__(self, d_model: int, n_head: int, attn_mask=None):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)  # seq-first
        self.ln_1 = nn.LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([
            ("c_fc", nn.Linear(d_model, d_model * 4)),
            ("gelu", QuickGELU()),
            ("c_proj", nn.Linear(d_model * 4, d_model)),
        ]))
        self.ln_2 = nn.LayerNorm(d_model)
        self.attn_mask = attn_mask

    def attention(self, x):
        mask = self.attn_mask
        if mask is not None:
            mask = mask.to(dtype=x.dtype, device=x.device)
        return self.attn(x, x, x, need_weights=False, attn_mask=mask)[0]

    def forward(self, x):
        x = x + self.attention(self.ln_1(x))
        x = x + self.mlp(self.ln_2(x))
        return x


class Transformer(nn.Module):
    def __init__(self, width, layers, heads, attn_mask=None):
        super().__init__()
        self.width, self.layers = width, layers
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads, attn_mask)
                                         for _ in range(layers)])

    def forward(self, x):
        return self.resblocks(x)


def resize_pos_embedding(pos, grid_now):
    """CLIP-HBA's ``pos_embedding=True`` path [upstream-memory]: bicubic resize of the
    patch part of the positional table to the current token grid; identity at 224^2."""
    n = pos.shape[0] - 1
    g = int(round(math.sqrt(n)))
    if g == grid_now:
        return pos
    cls, patch = pos[:1], pos[1:]
    patch = patch.reshape(1, g, g, -1).permute(0, 3, 1, 2)
    patch = F.interpolate(patch, size=(grid_now, grid_now), mode="bicubic", align_corners=False)
    patch = patch.permute(0, 2, 3, 1).reshape(grid_now * grid_now, -1)
    return torch.cat([cls, patch], 0)


class VisionTransformer(nn.Module):
    def __init__(self, input_resolution, patch_size, width, layers, heads, output_dim):
        super().__init__()
        self.input_resolution, self.output_dim = input_resolution, output_dim
        self.conv1 = nn.Conv2d(3, width, patch_size, patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(
            scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = nn.LayerNorm(width)
        self.transformer = Transformer(width, layers, heads)
        self.ln_post = nn.LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))

    def tokens(self, x, pos_embedding=False):
        """conv1 + cls + pos + ln_pre -> [B, T, width] (batch-first)."""
        x = self.conv1(x)
        grid = x.shape[-1]
        x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)
        cls = self.class_embedding.to(x.dtype) + torch.zeros(
            x.shape[0], 1, x.shape[-1], dtype=x.dtype, device=x.device)
        x = torch.cat([cls, x], dim=1)
        pos = self.positional_embedding
        if pos_embedding:
            pos = resize_pos_embedding(pos, grid)
        x = x + pos.to(x.dtype)
        return self.ln_pre(x)

    def forward(self, x, pos_embedding=False):
        x = self.tokens(x, pos_embedding)
        x = x.permute(1, 0, 2)
        x = self.transformer(x)
        x = x.permute(1, 0, 2)
        x = self.ln_post(x[:, 0, :])
        return x @ self.proj


class CLIP(nn.Module):
    def __init__(self, embed_dim, image_resolution, vision_layers, vision_width, vision_patch_size,
                 context_length, vocab_size, transformer_width, transformer_heads, transformer_layers):
        super().__init__()
        self.context_length = context_length
        self.visual = VisionTransformer(image_resolution, vision_patch_size, vision_width,
                                        vision_layers, vision_width // 64, embed_dim)
        self.transformer = Transformer(transformer_width, transformer_layers, transformer_heads,
                                       attn_mask=self.build_attention_mask())
        self.vocab_size = vocab_size
        self.token_embedding = nn.Embedding(vocab_size, transformer_width)
        self.positional_embedding = nn.Parameter(torch.empty(context_length, transformer_width))
        self.ln_final = nn.LayerNorm(transformer_width)
        self.text_projection = nn.Parameter(torch.empty(transformer_width, embed_dim))
        self.logit_scale = nn.Parameter(torch.ones([]) * math.log(1 / 0.07))
        self.initialize_parameters()

    def initialize_parameters(self):
        nn.init.normal_(self.token_embedding.weight, std=0.02)
        nn.init.normal_(self.positional_embedding, std=0.01)
        w, l = self.transformer.width, self.transformer.layers
        proj_std, attn_std, fc_std = (w ** -0.5) * ((2 * l) ** -0.5), w ** -0.5, (2 * w) ** -0.5
        for blk in self.transformer.resblocks:
            nn.init.normal_(blk.attn.in_proj_weight, std=attn_std)
            nn.init.normal_(blk.attn.out_proj.weight, std=proj_std)
            nn.init.normal_(blk.mlp.c_fc.weight, std=fc_std)
            nn.init.normal_(blk.mlp.c_proj.weight, std=proj_std)
        nn.init.normal_(self.text_projection, std=w ** -0.5)

    def build_attention_mask(self):
        mask = torch.empty(self.context_length, self.context_length)
        mask.fill_(float("-inf"))
        mask.triu_(1)
        return mask

    @property
    def dtype(self):
        return self.visual.conv1.weight.dtype

    def encode_image(self, image, pos_embedding=False):
        return self.visual(image.type(self.dtype), pos_embedding)

    def encode_text(self, text):
        x = self.token_embedding(text).type(self.dtype)
        x = x + self.positional_embedding.type(self.dtype)
        x = x.permute(1, 0, 2)
        x = self.transformer(x)
        x = x.permute(1, 0, 2)
        x = self.ln_final(x).type(self.dtype)
        return x[torch.arange(x.shape[0]), text.argmax(dim=-1)] @ self.text_projection

    def forward(self, image, text, pos_embedding=False):
        text = text.reshape(-1, text.shape[-1])
        img = self.encode_image(image, pos_embedding)
        txt = self.encode_text(text)
        img = img / img.norm(dim=1, keepdim=True)
        txt = txt / txt.norm(dim=1, keepdim=True)
        return self.logit_scale.exp() * img @ txt.t()


def arch_from_state_dict(sd):
    """Infer the constructor arguments from tensor shapes (as the published build_model does)."""
    vw = sd["visual.conv1.weight"].shape[0]
    vl = len({k.split(".")[3] for k in sd if k.startswith("visual.transformer.resblocks.")})
    ps = sd["visual.conv1.weight"].shape[-1]
    grid = round((sd["visual.positional_embedding"].shape[0] - 1) ** 0.5)
    tw = sd["ln_final.weight"].shape[0]
    tl = len({k.split(".")[2] for k in sd if k.startswith("transformer.resblocks.")})
    return dict(embed_dim=sd["text_projection"].shape[1], image_resolution=ps * grid, vision_layers=vl,
                vision_width=vw, vision_patch_size=ps, context_length=sd["positional_embedding"].shape[0],
                vocab_size=sd["token_embedding.weight"].shape[0], transformer_width=tw,
                transformer_heads=tw // 64, transformer_layers=tl)


def build_model(state_dict):
    model = CLIP(**arch_from_state_dict(state_dict))
    sd = {k: v for k, v in state_dict.items()
          if k not in ("input_resolution", "context_length", "vocab_size")}
    model.load_state_dict(sd)
    return model.eval()


def synthetic_state_dict(name="ViT-L/14", seed=1):
    """Seeded random-init weights of the named architecture (no network: no checkpoints).
    ``logit_scale`` is set to ln(100), the value the pretrained checkpoints saturate at
    (the reference logs show logits ~ 100*cos, RUNLOG:67-79)."""
    gen_state = torch.get_rng_state()
    try:
        torch.manual_seed(seed)
        model = CLIP(**ARCH[name])
        with torch.no_grad():
            model.logit_scale.fill_(math.log(100.0))
            # default nn.LayerNorm / zero-bias inits make several terms vanish; perturb them so
            # that every parameter of the graph is exercised by the parity tests
            for n, p in model.named_parameters():
                if n.endswith("bias"):
                    p.normal_(0.0, 0.02)
                elif ".ln_" in n or n.startswith("ln_") or "ln_p" in n:
                    p.add_(0.05 * torch.randn_like(p))
        return OrderedDict((k, v.detach().clone()) for k, v in model.state_dict().items())
    finally:
        torch.set_rng_state(gen_state)


def _download(url, root):
    """No network here: materialise the seeded synthetic checkpoint under ``root``."""
    os.makedirs(root, exist_ok=True)
    fname = os.path.join(root, os.path.basename(url))
    if not os.path.exists(fname):
        name = {v: k for k, v in _MODELS.items()}[url]
        torch.save(synthetic_state_dict(name), fname)
    return fname


def tokenize(texts, context_length=CONTEXT_LENGTH):
    """Synthetic stand-in for the BPE tokenizer (vocabulary file not available offline):
    [SOT, one stable pseudo-id per whitespace word ..., EOT, 0-pad].  EOT is the row max, which
    is all the model relies on (``text.argmax(-1)``).  Like the published tokenizer a str
    returns [1,77], so the reference's ``torch.stack([clip.tokenize(c) for c in classnames])``
    (NEW:282) yields [66,1,77]; ``CLIP.forward`` flattens that to [66,77]."""
    single = isinstance(texts, str)
    rows = []
    for t in ([texts] if single else list(texts)):
        ids = [SOT_TOKEN]
        for w in t.lower().replace(",", " , ").split()[: context_length - 2]:
            h = int.from_bytes(hashlib.sha256(w.encode()).digest()[:4], "little")
            ids.append(1 + h % (SOT_TOKEN - 1))
        ids.append(EOT_TOKEN)
        rows.append(ids + [0] * (context_length - len(ids)))
    return torch.tensor(rows, dtype=torch.long)

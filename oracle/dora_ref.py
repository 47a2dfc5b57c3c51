"""ORACLE (test infrastructure only — never imported by the product path).

CPU restatement of the reference's DoRA adapter and of the model surgery around it:
  DoRALayerRef        <- DoRALayer,          NEW:407-481 (init NEW:408-445, merge NEW:447-463)
  apply_dora_ref      <- apply_dora_to_ViT,  NEW:484-513 (vision blocks first, then text; the global
                                             torch RNG is consumed A-then-B per layer, NEW:443-445)
  switch_dora_ref     <- switch_dora_layers, NEW:516-544
Pinned against the reference's own classes by tests/test_oracle_cpu.py (run where /root/reference
exists) and by the golden fixtures under tests/golden/.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn


class DoRALayerRef(nn.Module):
    def __init__(self, original_layer, r=8, dora_alpha=16, dora_dropout=0.1):
        super().__init__()
        self.original_layer = original_layer
        self.r, self.dora_alpha = r, dora_alpha
        self.dora_dropout = nn.Dropout(p=dora_dropout)
        with torch.no_grad():
            Wt = original_layer.weight.data.clone().T
            S = torch.norm(Wt, dim=0)
            D = Wt / S
        self.m = nn.Parameter(S)
        self.register_buffer("D", D)
        self.delta_D_A = nn.Parameter(torch.zeros(r, original_layer.out_features))
        self.delta_D_B = nn.Parameter(torch.zeros(original_layer.in_features, r))
        self.scaling = dora_alpha / r
        nn.init.kaiming_uniform_(self.delta_D_A, a=math.sqrt(5))
        nn.init.kaiming_uniform_(self.delta_D_B, a=math.sqrt(5))
        self.bias = (nn.Parameter(original_layer.bias.data.clone())
                     if original_layer.bias is not None else None)

    @property
    def weight(self):
        d_new = self.D + (self.delta_D_B @ self.delta_D_A) * self.scaling
        norms = torch.norm(d_new, dim=0, keepdim=True) + 1e-8
        return ((d_new / norms) * self.m).T


def apply_dora_ref(model, n_vision_layers=1, n_transformer_layers=1, r=8, dora_dropout=0.1,
                   layer_cls=DoRALayerRef):
    for idx in range(-n_vision_layers, 0):
        blk = model.clip_model.visual.transformer.resblocks[idx]
        blk.attn.out_proj = layer_cls(blk.attn.out_proj, r=r, dora_dropout=dora_dropout)
    for idx in range(-n_transformer_layers, 0):
        blk = model.clip_model.transformer.resblocks[idx]
        blk.attn.out_proj = layer_cls(blk.attn.out_proj, r=r, dora_dropout=dora_dropout)


def switch_dora_ref(model, layer_cls=DoRALayerRef):
    for p in model.parameters():
        p.requires_grad = False
    for mod in model.modules():
        if isinstance(mod, layer_cls):
            mod.m.requires_grad = True
            mod.delta_D_A.requires_grad = True
            mod.delta_D_B.requires_grad = True


class CLIPHBARef(nn.Module):
    """CLIPHBA, NEW:268-304, over an already built CLIP model and pre-tokenised prompts."""

    def __init__(self, clip_model, tokenized_prompts, pos_embedding=True):
        super().__init__()
        self.clip_model = clip_model.float()
        self.pos_embedding = pos_embedding
        for p in self.clip_model.parameters():
            p.requires_grad = False
        self.tokenized_prompts = tokenized_prompts

    def forward(self, image):
        if self.clip_model.training:
            self.clip_model.eval()
        return self.clip_model(image, self.tokenized_prompts.to(image.device), self.pos_embedding).float()

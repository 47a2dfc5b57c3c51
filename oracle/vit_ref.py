"""ORACLE (test infrastructure only — never imported by the product path).

CPU restatement of the ViT-B/16 classifier the reference trains: `timm.create_model(
'vit_base_patch16_224', pretrained=False, num_classes=1000)` (VIT = Training/vit_training/baseline/
train_vit_sgd.py:283) with the training step of VIT:125-165 (cross-entropy, SGD 0.1 / 0.9 / 1e-4,
VIT:294-299) and the schedule of VIT:206-244.

timm is an un-vendored, un-pinned dependency (README.md:265; nothing under /root/reference): the
architecture is restated from its published definition — conv 16/16 patch embedding with bias,
cls token, learned 197-entry positional table, 12 pre-norm blocks [LayerNorm(eps 1e-6), fused qkv
Linear, 12 heads x 64, softmax attention, proj, LayerNorm, fc1 -> GELU(erf) -> fc2], final LayerNorm,
head on the cls token.  **Parity unpinned by the reference**; pinned here by an independent
implementation with copied weights: torchvision.models.vit_b_16 (tests/test_oracle_cpu.py::
test_vit_restatement_matches_torchvision).  Parameter names follow timm so that checkpoints
(VIT:343-349 `model.module.state_dict()`) keep their keys.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

LN_EPS = 1e-6


class Attention(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.num_heads = heads
        self.qkv = nn.Linear(dim, dim * 3)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, T, C = x.shape
        qkv = self.qkv(x).reshape(B, T, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        att = (q @ k.transpose(-2, -1)) * (q.shape[-1] ** -0.5)
        x = (att.softmax(dim=-1) @ v).transpose(1, 2).reshape(B, T, C)
        return self.proj(x)


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class Block(nn.Module):
    def __init__(self, dim, heads, mlp_ratio):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=LN_EPS)
        self.attn = Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim, eps=LN_EPS)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


class PatchEmbed(nn.Module):
    def __init__(self, patch, dim):
        super().__init__()
        self.proj = nn.Conv2d(3, dim, patch, patch)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class VisionTransformerRef(nn.Module):
    def __init__(self, img_size=224, patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0,
                 num_classes=1000):
        super().__init__()
        self.patch_embed = PatchEmbed(patch_size, embed_dim)
        n_tok = (img_size // patch_size) ** 2 + 1
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.randn(1, n_tok, embed_dim) * 0.02)
        self.blocks = nn.ModuleList([Block(embed_dim, num_heads, mlp_ratio) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=LN_EPS)
        self.head = nn.Linear(embed_dim, num_classes)

    def forward_features(self, x):
        x = self.patch_embed(x)
        x = torch.cat([self.cls_token.expand(x.shape[0], -1, -1), x], dim=1) + self.pos_embed
        for blk in self.blocks:
            x = blk(x)
        return self.norm(x)

    def forward(self, x):
        return self.head(self.forward_features(x)[:, 0])


CONFIGS = {"vit_base_patch16_224": dict(embed_dim=768, depth=12, num_heads=12),
           "vit_tiny_test": dict(embed_dim=128, depth=2, num_heads=2)}


def create_model(name="vit_base_patch16_224", num_classes=1000, seed=None, **kw):
    if seed is not None:
        torch.manual_seed(seed)
    m = VisionTransformerRef(num_classes=num_classes, **CONFIGS[name], **kw)
    for mod in m.modules():
        if isinstance(mod, nn.Linear):
            nn.init.trunc_normal_(mod.weight, std=0.02)
            nn.init.zeros_(mod.bias)
    nn.init.normal_(m.cls_token, std=1e-6)
    return m


def cosine_warmup_lr(base_lr, epoch, warmup_epochs, max_epochs, eta_min=0.0):
    """CosineAnnealingLRWithWarmup.step, VIT:221-236 (lr used for epoch index `epoch`)."""
    if epoch < warmup_epochs:
        return base_lr * (epoch + 1) / warmup_epochs
    prog = (epoch - warmup_epochs) / (max_epochs - warmup_epochs)
    return eta_min + (base_lr - eta_min) * 0.5 * (1 + math.cos(math.pi * prog))


def train_steps(model, batches, lr=0.1, momentum=0.9, weight_decay=1e-4):
    """VIT:132-152 without autocast/GradScaler (fp32): per batch forward, CE, backward, SGD step.
    Returns the list of losses."""
    opt = torch.optim.SGD(model.parameters(), lr=lr, momentum=momentum, weight_decay=weight_decay)
    losses = []
    for images, labels in batches:
        opt.zero_grad()
        loss = F.cross_entropy(model(images), labels)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    return losses

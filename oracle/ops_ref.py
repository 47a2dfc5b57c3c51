"""ORACLE (test infrastructure only — never imported by the product path).

CPU restatements (PyTorch fp32/fp64, numpy, scipy) of the individual operators of the hot path,
each citing the reference lines it follows.  NEW = Training/functions/
new_cvpr_train_behavior_things_pipeline.py of the reference.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def quickgelu(x):
    """un-vendored CLIP MLP activation x*sigmoid(1.702x) [upstream-memory]"""
    return x * torch.sigmoid(1.702 * x)


def dora_weight(D, A, Bm, m, scaling, eps=1e-8):
    """DoRALayer.weight, NEW:447-463: returns W [out, in]."""
    delta = (Bm @ A) * scaling
    d_new = D + delta
    norms = torch.norm(d_new, dim=0, keepdim=True) + eps
    return ((d_new / norms) * m).T


def dora_init(weight):
    """DoRALayer.__init__ decomposition, NEW:416-420: (m, D) of an [out,in] weight."""
    Wt = weight.detach().clone().T
    S = torch.norm(Wt, dim=0)
    return S, Wt / S


def attention(qkv, B, T, H, causal=False):
    """F.multi_head_attention_forward core (torch/nn/functional.py:6682 SDPA): qkv [B*T, 3*H*64]
    -> [B*T, H*64]."""
    d = H * 64
    q, k, v = qkv.view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    mask = None
    if causal:
        mask = torch.full((T, T), float("-inf"), dtype=qkv.dtype).triu_(1)
    o = F.scaled_dot_product_attention(q, k, v, attn_mask=mask)
    return o.permute(0, 2, 1, 3).reshape(B * T, d)


def cos_logits(img, txt, logit_scale):
    """tail of the un-vendored CLIP.forward (contract NEW:298-300)."""
    img = img / img.norm(dim=1, keepdim=True)
    txt = txt / txt.norm(dim=1, keepdim=True)
    return logit_scale.exp() * img @ txt.t()


def rdm_and_spearman(emb, reference_rdm):
    """behavioral_RSA tail, NEW:625-652: (rho, p, model_rdm)."""
    from scipy.stats import spearmanr
    model_rdm = 1 - np.corrcoef(np.asarray(emb))
    np.fill_diagonal(model_rdm, 0)
    iu = np.triu_indices_from(reference_rdm, k=1)
    rho, p = spearmanr(reference_rdm[iu], model_rdm[iu])
    return rho, p, model_rdm


def rankdata_average(x):
    """scipy.stats.rankdata(method='average'), the ranking inside spearmanr (NEW:652)."""
    from scipy.stats import rankdata
    return rankdata(np.asarray(x), method="average")


def adamw_reference(params, grads_per_step, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01):
    """torch.optim.AdamW as used at NEW:1181; returns the parameter list after the given steps."""
    ps = [torch.nn.Parameter(p.clone()) for p in params]
    opt = torch.optim.AdamW(ps, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
    for grads in grads_per_step:
        for p, g in zip(ps, grads):
            p.grad = g.clone()
        opt.step()
    return [p.detach() for p in ps], opt

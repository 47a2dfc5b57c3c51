"""DoRA adapter with the merge and its backward fused into single sm_100a kernels.

Drop-in for the reference's ``DoRALayer`` (NEW:407-481): same constructor, attribute names
(``m``, ``D`` buffer, ``delta_D_A``, ``delta_D_B``, ``bias``, ``scaling``, ``original_layer``,
``dora_dropout``), same initialisation order on the global torch RNG (A then B, Kaiming-uniform
a=sqrt(5), NEW:443-445) and the same ``weight`` property semantics:

    W = [ m * (D + s B A) / (||D + s B A||_col + 1e-8) ]^T        (NEW:447-463)

``weight`` is what ``nn.MultiheadAttention`` reads (torch/nn/modules/activation.py:1504-1505), so
it is the hot path; ``forward`` is dead code in the reference and kept only for API parity.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import engine as _engine
from . import ops
from .ops import Operand

DORA_EPS = 1e-8


class _DoraMerge(torch.autograd.Function):
    """(D, A, Bm, m) -> w_t [in, out] fp32 plus the bf16 GEMM operands of W and W^T (emitted by the
    same kernel so that the consuming out_proj GEMM needs no separate staging pass)."""

    @staticmethod
    def forward(ctx, D, A, Bm, m, scale):
        in_f, out_f = D.shape
        split = _engine.get_precision() == "fp32"
        w_t = torch.empty(in_f, out_f, device=D.device, dtype=torch.float32)
        w = Operand.empty(out_f, in_f, split, D.device)
        wt = Operand.empty(in_f, out_f, split, D.device)
        ops.dora_merge_fwd(D, A.detach().contiguous(), Bm.detach().contiguous(),
                           m.detach().contiguous(), scale, DORA_EPS, w_t_f32=w_t, w=w, wt=wt)
        ctx.save_for_backward(D, A, Bm, m)
        ctx.scale = scale
        ctx.mark_non_differentiable(w.buf, wt.buf)
        ctx.set_materialize_grads(False)   # no zero tensors for the two bf16 operand outputs in backward
        return w_t, w.buf, wt.buf

    @staticmethod
    def backward(ctx, g_wt, _g1, _g2):
        D, A, Bm, m = ctx.saved_tensors
        in_f, out_f = D.shape
        if g_wt is None:
            return None, None, None, None, None
        G = g_wt.t()  # dL/dW in [out, in] layout
        if not G.is_contiguous():
            G = G.contiguous()
        G = G.float()
        dm = torch.empty_like(m)
        dA = torch.empty_like(A)
        dB = torch.empty_like(Bm)
        ws = torch.empty(in_f, out_f, device=D.device, dtype=torch.float32)
        ops.dora_merge_bwd(G, D, A.detach().contiguous(), Bm.detach().contiguous(),
                           m.detach().contiguous(), ctx.scale, DORA_EPS, dm, dA, dB, ws)
        return None, dA, dB, dm, None


class DoRALayer(nn.Module):
    def __init__(self, original_layer, r=8, dora_alpha=16, dora_dropout=0.1):
        super().__init__()
        self.original_layer = original_layer
        self.r = r
        self.dora_alpha = dora_alpha
        self.dora_dropout = nn.Dropout(p=dora_dropout)
        with torch.no_grad():
            # The reference builds its adapters while the model is still on the CPU (NEW:1133-1152, .to(device)
            # comes at NEW:1178), so S and D are CPU fp32 results.  A sweep worker re-adapts a frozen CLIP
            # that already lives on the GPU: computing the decomposition there would change S / D in the
            # last bit from condition to condition (tests/test_gpu_sweep.py found exactly that), so it is
            # always done on the CPU and moved to wherever the wrapped layer lives.
            dev = original_layer.weight.device
            Wt = original_layer.weight.data.detach().to("cpu", torch.float32).clone().T      # [in, out]
            S = torch.norm(Wt, dim=0)                      # column magnitudes [out]
            D = Wt / S                                     # unit-norm columns
        self.m = nn.Parameter(S.to(dev))
        self.register_buffer("D", D.contiguous().to(dev))
        self.delta_D_A = nn.Parameter(torch.zeros(self.r, original_layer.out_features))
        self.delta_D_B = nn.Parameter(torch.zeros(original_layer.in_features, self.r))
        self.scaling = self.dora_alpha / self.r
        self.reset_parameters()
        if self.original_layer.bias is not None:
            self.bias = nn.Parameter(original_layer.bias.data.clone())
        else:
            self.bias = None

    @property
    def in_features(self):
        return self.original_layer.in_features

    @property
    def out_features(self):
        return self.original_layer.out_features

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.delta_D_A, a=math.sqrt(5))
        nn.init.kaiming_uniform_(self.delta_D_B, a=math.sqrt(5))

    @property
    def weight(self):
        if not self.D.is_cuda:
            raise RuntimeError("hba.DoRALayer has no CPU path: move the model to a CUDA device")
        D = self.D if self.D.is_contiguous() else self.D.contiguous()
        w_t, w_buf, wt_buf = _DoraMerge.apply(D, self.delta_D_A, self.delta_D_B, self.m,
                                              float(self.scaling))
        W = w_t.t()
        out_f, in_f = W.shape
        lo = lambda K: K if _engine.get_precision() == "fp32" else 0
        W._hba_ops = (Operand(w_buf, out_f, in_f, lo(in_f)), Operand(wt_buf, in_f, out_f, lo(out_f)),
                      _engine.get_precision())
        return W

    def forward(self, x):
        # Dead code on the CLIP path (seq-first MHA reads .weight/.bias as tensors); kept so that
        # the layer still works when called directly.  Dropout on delta_D as in NEW:468.
        if self.training and self.dora_dropout.p > 0:
            delta = self.dora_dropout((self.delta_D_B @ self.delta_D_A) * self.scaling)
            d_new = self.D + delta
            W = (d_new / (torch.norm(d_new, dim=0, keepdim=True) + DORA_EPS) * self.m).T
        else:
            W = self.weight
        return F.linear(x, W, self.bias)

"""Fused multi-tensor optimisers on libhba (one kernel launch per step).

``FusedAdamW`` subclasses ``torch.optim.AdamW`` so that defaults, ``state`` layout
(``step`` / ``exp_avg`` / ``exp_avg_sq``) and ``state_dict()`` are exactly torch's — the reference
saves and restores ``optimizer.state_dict()`` every epoch (NEW:712, NEW:124-126) and those files
must stay interchangeable.  Only ``step`` is replaced (reference: AdamW(model.parameters(), lr),
NEW:1181; .step at NEW:1001).
"""
from __future__ import annotations

import torch

from . import ops


def _table(tensor_lists, device):
    n = len(tensor_lists[0])
    flat = [lst[i].data_ptr() for i in range(n) for lst in tensor_lists]
    sizes = [t.numel() for t in tensor_lists[0]]
    return (torch.tensor(flat, dtype=torch.int64, device=device),
            torch.tensor(sizes, dtype=torch.int64, device=device), n, sum(sizes))


def _check(p):
    if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
        raise RuntimeError("hba optimisers need contiguous fp32 CUDA parameters (no CPU fallback)")


class FusedAdamW(torch.optim.AdamW):
    """The step count lives on the device (one int32 per param group, advanced only by steps that are
    not skipped — exactly the reference, whose NaN guard `continue`s before optimizer.step, NEW:989-998)
    so that a whole training step can be captured into a CUDA graph; torch's per-parameter
    ``state[p]["step"]`` tensors are refreshed from it whenever the state is read through
    ``state_dict()`` (what the reference's checkpoints save, NEW:712)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, foreach=False)
        self._tables = {}
        self._step_dev = {}   # group index -> device int32 [1]

    def sync_state(self):
        """Device step counters -> ``state[p]["step"]`` (one small device read per group)."""
        for gi, ctr in self._step_dev.items():
            n = float(int(ctr))
            for p in self.param_groups[gi]["params"]:
                st = self.state.get(p)
                if st is not None and "step" in st:
                    st["step"] = torch.tensor(n, dtype=torch.float32)

    def state_dict(self):
        self.sync_state()
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._step_dev = {}
        self._tables = {}

    def zero_grad(self, set_to_none=False):
        """Zeroes the gradients in place (stable device pointers for the fused kernel's table);
        parameters that have never received a gradient keep ``grad is None`` as in torch."""
        grads = [p.grad for g in self.param_groups for p in g["params"] if p.grad is not None]
        if set_to_none:
            return super().zero_grad(set_to_none=True)
        if grads:
            torch._foreach_zero_(grads)

    @torch.no_grad()
    def step(self, closure=None, skip_flag=None):
        """``skip_flag``: optional device int32 tensor; a non-zero value skips the update on the
        device (the reference's NaN/Inf guard, NEW:989-998, without a host sync)."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                _check(p)
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p in ps]
            m = [self.state[p]["exp_avg"] for p in ps]
            v = [self.state[p]["exp_avg_sq"] for p in ps]
            key = tuple(t.data_ptr() for lst in (ps, grads, m, v) for t in lst)
            cached = self._tables.get(gi)
            if cached is None or cached[0] != key:
                cached = (key, _table([ps, grads, m, v], ps[0].device))
                self._tables[gi] = cached
            table, sizes, n, total = cached[1]
            ctr = self._step_dev.get(gi)
            if ctr is None:
                ctr = torch.tensor([int(self.state[ps[0]]["step"])], dtype=torch.int32, device=ps[0].device)
                self._step_dev[gi] = ctr
            if skip_flag is None:
                ctr += 1
            else:
                ctr += (skip_flag.reshape(1) == 0).to(torch.int32)
            b1, b2 = group["betas"]
            ops.adamw_multi(table, sizes, n, total, float(group["lr"]), b1, b2, group["eps"],
                            group["weight_decay"], 0, skip_flag, step_dev=ctr)
            self._keepalive = grads
        return loss


class FusedSGD(torch.optim.SGD):
    """SGD(momentum, weight_decay) of VIT:294-299 as one multi-tensor launch."""

    def __init__(self, params, lr, momentum=0.0, weight_decay=0.0):
        super().__init__(params, lr=lr, momentum=momentum, weight_decay=weight_decay, foreach=False)
        self._tables = {}

    @torch.no_grad()
    def step(self, closure=None, skip_flag=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            first = False
            for p in ps:
                _check(p)
                st = self.state[p]
                if st.get("momentum_buffer") is None:
                    st["momentum_buffer"] = torch.zeros_like(p)
                    first = True
            grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p in ps]
            bufs = [self.state[p]["momentum_buffer"] for p in ps]
            key = tuple(t.data_ptr() for lst in (ps, grads, bufs) for t in lst)
            cached = self._tables.get(gi)
            if cached is None or cached[0] != key:
                cached = (key, _table([ps, grads, bufs], ps[0].device))
                self._tables[gi] = cached
            table, sizes, n, total = cached[1]
            ops.sgd_multi(table, sizes, n, total, float(group["lr"]), group["momentum"],
                          group["weight_decay"], first, skip_flag)
            self._keepalive = grads
        return loss

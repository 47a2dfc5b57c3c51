"""Data-parallel plumbing of the ViT-B/16 trainer (reference VIT:287, DistributedDataParallel): gradient
buckets of one flat fp32 buffer, all-reduced asynchronously as the backward pass completes them.

Device-agnostic on purpose (works on `gloo` + CPU tensors as well as `nccl` + CUDA): the N > 1 logic is
tested with world size 2 on CPU (tests/test_dp_cpu.py), the GPU path only swaps the backend.
DDP averages gradients (SUM all-reduce, then / world); here 1/world is folded into dL/dlogits before the
backward pass (`fold_world_size`), so the buckets need a plain SUM."""
from __future__ import annotations

import torch


def layout_buckets(groups, align=4):
    """groups: [(name, [tensor-like with .numel()])] -> (total elements, {id(p): (offset, numel)},
    [(name, start, end)]): every parameter's gradient starts `align`-element aligned (16 bytes for fp32:
    the GEMM epilogue stores float4) and a group's gradients are contiguous = one all-reduce bucket."""
    pad = lambda n: (n + align - 1) // align * align
    offsets, buckets, off = {}, [], 0
    for name, ps in groups:
        start = off
        for p in ps:
            offsets[id(p)] = (off, p.numel())
            off += pad(p.numel())
        buckets.append((name, start, off))
    return off, offsets, buckets


def fold_world_size(d_logits, world):
    """mean over the GLOBAL batch = (1/world) * sum over ranks of the local-batch mean gradient."""
    if world > 1:
        d_logits.mul_(1.0 / world)
    return d_logits


class BucketAllReducer:
    def __init__(self, dist=None, group=None):
        self.dist, self.group = dist, group
        self.handles = []
        self.launched = []
        self.enabled = True   # False: buckets are announced but not reduced (compute-only timing of a multi-rank step)

    @property
    def world(self):
        return self.dist.get_world_size(self.group) if self.dist is not None else 1

    def on_bucket_ready(self, name, flat_slice):
        """Called by the backward pass as soon as every gradient of the bucket has been written."""
        self.launched.append(name)
        if self.enabled and self.dist is not None and self.world > 1:
            self.handles.append(self.dist.all_reduce(flat_slice, group=self.group, async_op=True))

    def wait(self):
        for h in self.handles:
            h.wait()
        self.handles.clear()
        self.launched.clear()

    def broadcast_parameters(self, params, src=0):
        if self.dist is not None and self.world > 1:
            for p in params:
                self.dist.broadcast(p.data if isinstance(p, torch.nn.Parameter) else p, src=src, group=self.group)

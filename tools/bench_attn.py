"""Attention micro-benchmark (tcgen05 path, bf16): CLIP ViT-L/14 vision (B=32, T=257, H=16), CLIP text
(66 x 77, H=12, causal), ViT-B/16 (B=256, T=197, H=12).  Algorithmic FLOPs = 4 * B * H * T^2 * 64."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vit-project_b200")]

import torch  # noqa: E402

import hba  # noqa: E402
from hba import ops  # noqa: E402

dev = torch.device("cuda", 0)
out = {}
for name, B, T, H, causal in (("clip_vision", 32, 257, 16, False), ("clip_text", 66, 77, 12, True),
                              ("vitb16", 256, 197, 12, False)):
    d = H * 64
    qkv = (torch.randn(B * T, 3 * d, device=dev) * 0.5).to(torch.bfloat16)
    o = ops.Operand.empty(B * T, d, False, dev)
    for _ in range(3):
        ops.attention_fwd(qkv, B, T, H, causal=causal, out=o)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    e0.record()
    for _ in range(reps):
        ops.attention_fwd(qkv, B, T, H, causal=causal, out=o)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 4.0 * B * H * T * T * 64
    out[name] = {"us": ms * 1e3, "tflops": fl / ms / 1e9, "exp_per_s_T": B * H * T * T / ms / 1e9}
print(json.dumps(out))

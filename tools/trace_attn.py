"""Phase timeline of the tcgen05 attention kernel (CTA 0) from the hba_debug_attention_trace hook."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vit-project_b200")]
import torch  # noqa: E402
import hba  # noqa: E402
from hba import ops, _lib  # noqa: E402

dev = torch.device("cuda", 0)
B, T, H = (256, 197, 12) if len(sys.argv) < 2 else (32, 257, 16)
d = H * 64
qkv = (torch.randn(B * T, 3 * d, device=dev) * 0.5).to(torch.bfloat16)
o = ops.Operand.empty(B * T, d, False, dev)
for _ in range(3):
    ops.attention_fwd(qkv, B, T, H, out=o)
buf = torch.zeros(512, dtype=torch.int64, device=dev)
lib = _lib.load()
lib.hba_debug_attention_trace.argtypes = [C.c_void_p]
lib.hba_debug_attention_trace.restype = None
lib.hba_debug_attention_trace(C.c_void_p(buf.data_ptr()))
ops.attention_fwd(qkv, B, T, H, out=o)
torch.cuda.synchronize()
lib.hba_debug_attention_trace(None)
t = buf.cpu().view(64, 8)
t0 = int(t[0, 0])
names = ["S_issue", "PV_issue", "S_seen", "max_done", "P_written", "O_seen", "epi_done", "O_in_regs"]
print("item " + " ".join(f"{n:>10s}" for n in names))
for i in range(12):
    print(f"{i:4d} " + " ".join(f"{int(t[i, k]) - t0:10d}" for k in range(8)))

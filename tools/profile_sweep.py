"""Host-side profile of the cached sweep epochs (bench.measure_sweep): cProfile top functions."""
import cProfile
import io
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vit-project_b200")]

import torch  # noqa: E402

import bench  # noqa: E402
import hba  # noqa: E402
from hba import ops  # noqa: E402


class A:
    batch, backbone, precision = 32, "ViT-L/14", "bf16"


dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
hba.set_precision("bf16")
pr = cProfile.Profile()
c0 = ops.COUNTERS["launches"]
pr.enable()
sw = bench.measure_sweep(A, dev, 1)
pr.disable()
print(sw)
print("launches", ops.COUNTERS["launches"] - c0)
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(70)
print(s.getvalue()[:14000])

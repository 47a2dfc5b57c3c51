"""One trunk-cached CLIP-HBA training step (eager launches) inside cudaProfilerStart/Stop, for
`ncu --profile-from-start off` launch lists of the sweep's steady-state step."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vit-project_b200")]

import torch  # noqa: E402

import bench  # noqa: E402
import hba  # noqa: E402
from functions import _pipeline_core as core  # noqa: E402


class A:
    batch, backbone, precision = 32, "ViT-L/14", "bf16"


dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
hba.set_precision("bf16")
model, opt = bench.build_gpu_model(A, dev)
eng = model.clip_model.hba_engine()
eng.cache_text = True
core.enable_trunk_cache(model, 64)
crit = torch.nn.MSELoss()
g = torch.Generator().manual_seed(0)
images = torch.randn(32, 3, 224, 224, generator=g).to(dev)
targets = (torch.randn(32, 66, generator=g) * 9.5 + 5.75).to(dev)
ids = list(range(32))


def step():
    opt.zero_grad()
    eng.batch_ids = ids
    loss = crit(model(images), targets)
    loss.backward()
    opt.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
e0.record()
loss = step()
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(loss), "ms", e0.elapsed_time(e1))

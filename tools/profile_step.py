"""Runs warm-up steps, then exactly `--steps` CLIP-HBA training steps inside
cudaProfilerStart/Stop so that `ncu --profile-from-start off` sees only the steady-state step."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vit-project_b200")]

os.environ.setdefault("HBA_TEXT_STREAM", "0")  # serialised kernels: clean per-kernel durations under ncu
import torch  # noqa: E402

import bench  # noqa: E402
import hba  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--backbone", default="ViT-L/14")
ap.add_argument("--precision", default="bf16")
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
hba.set_precision(a.precision)
model, opt = bench.build_gpu_model(a, dev)
model.clip_model.hba_engine().cache_text = False
crit = torch.nn.MSELoss()
g = torch.Generator().manual_seed(0)
images = torch.randn(a.batch, 3, 224, 224, generator=g).to(dev)
targets = (torch.randn(a.batch, 66, generator=g) * 9.5 + 5.75).to(dev)


def step():
    opt.zero_grad()
    loss = crit(model(images), targets)
    loss.backward()
    opt.step()
    return loss


for _ in range(a.warmup):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(a.steps):
    loss = step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(loss))

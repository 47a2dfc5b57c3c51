"""Wall-clock per phase of the cached sweep epochs (train / eval / RSA / checkpoints / CSV), with a
device synchronize after every phase."""
import collections
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vit-project_b200")]

import torch  # noqa: E402

import bench  # noqa: E402
import hba  # noqa: E402
import functions.new_cvpr_train_behavior_things_pipeline as NEW  # noqa: E402
from functions import _pipeline_core as core  # noqa: E402

T = collections.defaultdict(list)


def wrap(mod, name):
    fn = getattr(mod, name)

    def inner(*a, **k):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize()
        T[name].append(time.perf_counter() - t0)
        return r
    setattr(mod, name, inner)


for n in ("train_one_epoch", "evaluate_model", "behavioral_RSA", "save_dora_parameters", "save_random_states",
          "append_csv_row", "load_random_states"):
    for mod in (NEW, core):
        if hasattr(mod, n):
            wrap(mod, n)


class A:
    batch, backbone, precision = 32, "ViT-L/14", "bf16"


dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
hba.set_precision("bf16")
if len(sys.argv) > 1 and sys.argv[1] == "prelude":
    # what bench.main does before the sweep section: a full-trunk model, a few steps, then delete
    model, opt = bench.build_gpu_model(A, dev)
    model.clip_model.hba_engine().cache_text = False
    crit = torch.nn.MSELoss()
    x = torch.randn(32, 3, 224, 224, device=dev)
    y = torch.randn(32, 66, device=dev)
    for _ in range(6):
        opt.zero_grad()
        loss = crit(model(x), y)
        loss.backward()
        opt.step()
    print("prelude loss", float(loss), "allocated GB", torch.cuda.memory_allocated() / 2**30)
    if len(sys.argv) > 2 and sys.argv[2] == "keep":
        KEEP = (model, opt)
    del model, opt, loss
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    print("after del: allocated GB", torch.cuda.memory_allocated() / 2**30, "reserved", torch.cuda.memory_reserved() / 2**30)
sw = bench.measure_sweep(A, dev, 1)
print("end: allocated GB", torch.cuda.memory_allocated() / 2**30, "reserved", torch.cuda.memory_reserved() / 2**30)
print(sw)
for k, v in T.items():
    print(f"{k:24s} n={len(v):3d} " + " ".join(f"{1e3 * x:7.1f}" for x in v))

"""ViT-B/16 classification baseline (BASELINE.json configs[2]; reference VIT = Training/vit_training/
baseline/train_vit_sgd.py): synthetic 224^2 images / 1000 classes, SGD 0.1 / 0.9 / 1e-4, bf16 tensor-core
GEMMs with fp32 accumulation, data parallel over NCCL (one process per GPU, gradient all-reduce per
block bucket overlapped with the backward pass).

    python tools/bench_vit.py [--batch 256] [--steps 10]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/bench_vit.py --batch 256

Prints one JSON line: whole-job images/s (max-over-ranks device time), algorithmic training FLOPs
(3 x 35.1 GFLOP per image) against the measured sustained bf16 peak.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vit-project_b200")]

import torch  # noqa: E402

import hba  # noqa: E402
from hba import ops, vit  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256, help="per-GPU batch (reference: 256, VIT:250)")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--model", default="vit_base_patch16_224")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host (no CUDA graph)")
    ap.add_argument("--profile", action="store_true",
                    help="bracket the timed steps with cudaProfilerStart/Stop (ncu --profile-from-start off)")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    hba.set_precision(a.precision)
    torch.manual_seed(0)
    model = vit.create_model(a.model, num_classes=1000).to(dev)
    tr = vit.DataParallelTrainer(model, lr=0.1, momentum=0.9, weight_decay=1e-4, use_graph=not a.no_graph)
    tr.broadcast_parameters()
    g = torch.Generator(device=dev).manual_seed(rank)
    images = torch.randn(a.batch, 3, 224, 224, device=dev, generator=g)
    labels = torch.randint(0, 1000, (a.batch,), device=dev, generator=g)
    for _ in range(a.warmup):
        loss, _ = tr.step(images, labels)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    c0 = ops.COUNTERS["launches"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if a.profile:
        torch.cuda.profiler.start()
    e0.record()
    for _ in range(a.steps):
        loss, _ = tr.step(images, labels)
    e1.record()
    if a.profile:
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    value = world * a.batch * a.steps / (ms / 1e3)
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(peaks))["bf16_tflops_sustained"] if os.path.exists(peaks) else 1400.0
    tflops_per_gpu = value / world * 105.3e9 / 1e12
    if rank == 0:
        print(json.dumps({"metric": "ViT-B/16 train imgs/s", "value": value, "unit": "images/s", "n_gpus": world,
                          "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "scaling": "weak",
                          "dtype": "bf16" if a.precision == "bf16" else "bf16x3 (fp32 mode)", "data": "synthetic",
                          "config": {"workload": f"{a.model}, batch {a.batch}/GPU, 1000 classes, SGD 0.1/0.9/1e-4, "
                                                 "fused CE, bucketed NCCL gradient all-reduce", "global_batch": a.batch * world},
                          "loss": float(loss), "gpu_launches": ops.COUNTERS["launches"] - c0,
                          "algorithmic_tflops_per_gpu": tflops_per_gpu, "frac_of_sustained_bf16_peak": tflops_per_gpu / peak}))
    sys.stdout.flush()
    if dist is not None:
        if not a.no_graph:
            # a captured graph holds NCCL kernels of the communicator: tearing the process group down
            # with live graphs hung the ranks at exit (observed at N=2) - leave without the teardown
            os._exit(0)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

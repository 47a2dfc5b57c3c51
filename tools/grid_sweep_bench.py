#!/usr/bin/env python
"""The perturbation sweeps (BASELINE.json configs 2 and 4) run FOR REAL through the scheduler `hba.sweep.run_sweep`:
worker processes pinned one per GPU, conditions handed out longest-first (or LEN's resume chains with --chain), every
condition = the unmodified `run_behavioral_training(config)` of the drop-in pipeline (NEW) resuming from a baseline
checkpoint, trained until the reference's own early stopping ends it (patience 20, frozen inside the perturbation
window, NEW:1049-1063), per-epoch evaluation + RSA + CSV row + DoRA / random-state checkpoints on disk.

    python tools/grid_sweep_bench.py --kind grid   --gpus 0,1,2,3,4,5,6,7 --out gpurun_out/grid_n8.json
    python tools/grid_sweep_bench.py --kind single --gpus 0 --limit 24 --out gpurun_out/single_n1.json

Everything is synthetic and offline: a THINGS-shaped dataset on disk (1,806 training images -> random 66-D targets,
48 RSA images, RDM48_triplet.mat), seeded random-init ViT-L/14 weights, and a baseline run (BASE pipeline) that
writes the per-epoch checkpoints the conditions resume from.  Wall clock includes worker start-up (process spawn,
`import torch`, weight staging, image decoding), the cache-fill epoch and the graph captures of every worker.
The JSON carries, per condition, the number of epochs trained and the sha256 of its result CSV: two runs of the same
slice on different GPU counts must agree hash for hash ("matched loss / RSA trajectories").

Reference: SWEEP = Training/clip_behavioral_finetuning/uniform_sweep/clip_train_behavior_sweep.py:192-223,
LEN = Training/clip_behavioral_finetuning/length_experiments/clip_train_behavior_lengths.py:86-266.
"""
import argparse
import hashlib
import json
import os
import shutil
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vit-project_b200")]


def write_dataset(root, n_train=1806, n_rsa=48, seed=0):
    """THINGS-shaped files: PNG images (decoded and resized to 224^2 once per worker by the pipeline's own
    datasets), the SPoSE csv layout (index, image name, 66 target columns, NEW:191-202), the inference csv and
    RDM48_triplet.mat."""
    import numpy as np
    import pandas as pd
    import scipy.io
    from PIL import Image
    rng = np.random.default_rng(seed)
    img_dir = os.path.join(root, "imgs")
    os.makedirs(img_dir, exist_ok=True)

    def make(prefix, n):
        names = []
        for i in range(n):
            name = f"{prefix}{i:04d}.png"
            Image.fromarray(rng.integers(0, 255, (32, 32, 3), dtype=np.uint8)).save(os.path.join(img_dir, name))
            names.append(name)
        return names
    tr, rs = make("train", n_train), make("rsa", n_rsa)
    cols = {"image": tr}
    for k in range(66):
        cols[f"dim{k}"] = rng.standard_normal(n_train) * 9.5 + 5.75
    pd.DataFrame(cols).to_csv(os.path.join(root, "train.csv"))
    cols = {"image": rs}
    for k in range(66):
        cols[f"dim{k}"] = rng.standard_normal(n_rsa)
    pd.DataFrame(cols).to_csv(os.path.join(root, "rsa.csv"))
    rdm = 1 - np.corrcoef(rng.standard_normal((n_rsa, 66)))
    np.fill_diagonal(rdm, 0)
    scipy.io.savemat(os.path.join(root, "RDM48_triplet.mat"), {"RDM48_triplet": rdm})
    return img_dir


def _baseline_worker(cfg):
    """Runs in a spawned process (so that the parent never initialises CUDA before the sweep workers pin their
    GPUs): the baseline run whose checkpoints the conditions resume from."""
    from hba import sweep
    os.environ["CUDA_VISIBLE_DEVICES"] = sweep.pinned_device_env(cfg.pop("_gpu"))   # index into the launcher's own list
    import torch
    import functions.cvpr_train_behavior_things_pipeline_baseline as BASE
    cfg["criterion"] = torch.nn.MSELoss()
    BASE.run_behavioral_training(cfg)


def csv_digest(path):
    if not os.path.exists(path):
        return None, 0
    data = open(path, "rb").read()
    return hashlib.sha256(data).hexdigest()[:16], max(0, data.count(b"\n") - 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", choices=["grid", "single"], default="grid")
    ap.add_argument("--gpus", default="0")
    ap.add_argument("--limit", type=int, default=0, help="only the first N conditions in LPT order (0 = all)")
    ap.add_argument("--per-gpu", type=int, default=0, help="weak-scaling slice: N conditions per GPU")
    ap.add_argument("--max-start", type=int, default=0, help="only conditions whose start epoch is <= this")
    ap.add_argument("--chain", action="store_true")
    ap.add_argument("--workers-per-gpu", type=int, default=1,
                    help="worker processes per GPU (SURVEY N1: several conditions per GPU; they share the GPU by "
                         "time slicing, or concurrently under an MPS daemon)")
    ap.add_argument("--backbone", default="ViT-L/14")
    ap.add_argument("--batch-size", type=int, default=32)
    ap.add_argument("--n-train", type=int, default=1806)
    ap.add_argument("--patience", type=int, default=20)
    ap.add_argument("--epochs", type=int, default=500)
    ap.add_argument("--root", default="/tmp/hba_grid")
    ap.add_argument("--keep", action="store_true")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    os.environ.setdefault("HBA_SYNTHETIC_OK", "1")
    import multiprocessing as mp
    from hba import sweep
    gpus = [int(x) for x in a.gpus.split(",")]
    devices = [g for g in gpus for _ in range(max(1, a.workers_per_gpu))]
    conds = sweep.length_grid_conditions() if a.kind == "grid" else sweep.single_epoch_conditions(1, 98)
    if a.max_start:
        conds = [c for c in conds if c["training_run"] <= a.max_start]
    conds = sweep.lpt_order(conds)
    limit = a.per_gpu * len(gpus) if a.per_gpu else a.limit
    if limit:
        conds = conds[:limit]
    layout = "length" if a.kind == "grid" else "sweep"
    root = a.root
    shutil.rmtree(root, ignore_errors=True)
    os.makedirs(root)
    t_all = time.time()
    img_dir = write_dataset(root, n_train=a.n_train)
    t_data = time.time() - t_all
    common = {"csv_file": f"{root}/train.csv", "img_dir": img_dir, "inference_csv_file": f"{root}/rsa.csv",
              "RDM48_triplet_dir": f"{root}/RDM48_triplet.mat", "backbone": a.backbone, "batch_size": a.batch_size,
              "lr": 3e-4, "random_seed": 1, "vision_layers": 2, "transformer_layers": 1, "rank": 32, "cuda": 0,
              "logger": None}
    # ---- baseline: one checkpoint per epoch up to the latest resume point of the slice
    base_epochs = max(1, max(c["training_run"] for c in conds) - 1)
    base_cfg = dict(common, epochs=base_epochs, train_portion=0.8, early_stopping_patience=10 ** 6,
                    checkpoint_path=f"{root}/base/model.pth", training_res_path=f"{root}/base/res.csv",
                    dora_parameters_path=f"{root}/base/dora", random_state_path=f"{root}/base/rand",
                    _gpu=devices[0])
    os.makedirs(f"{root}/base", exist_ok=True)
    t0 = time.time()
    ctx = mp.get_context("spawn")
    p = ctx.Process(target=_baseline_worker, args=(base_cfg,))
    p.start()
    p.join()
    if p.exitcode != 0:
        raise SystemExit(f"baseline run failed (exit code {p.exitcode})")
    t_base = time.time() - t0
    # ---- the sweep through the scheduler
    import torch.nn as nn
    sweep_cfg = dict(common, epochs=a.epochs, early_stopping_patience=a.patience, hba_resident=True,
                     criterion=nn.MSELoss(),
                     baseline_dora_directory=f"{root}/base/dora", baseline_random_state_path=f"{root}/base/rand",
                     baseline_split_indices_path=f"{root}/base/rand/dataset_split_indices.pth",
                     perturb_type="random_target", perturb_length=1, perturb_distribution="target",
                     perturb_seed=42, previous_training_res_path=f"{root}/base/res.csv",
                     output_base_directory=f"{root}/out")
    logs = []
    t0 = time.time()
    results = sweep.run_sweep(sweep_cfg, conds, devices, layout=layout, chain=a.chain and layout == "length",
                              log=lambda s: (logs.append(s), print(s, file=sys.stderr, flush=True)))
    wall = time.time() - t0
    per_cond, epochs_total = [], 0
    for r in results:
        c = r["condition"]
        cfg = sweep.condition_config(sweep_cfg, c, layout)
        digest, rows = csv_digest(cfg["training_res_path"])
        resume = max(0, c["training_run"] - 1)
        trained = max(0, rows - resume)
        epochs_total += trained
        per_cond.append({"training_run": c["training_run"], "perturb_length": c.get("perturb_length", 1),
                         "ok": r["ok"], "worker": r["worker"], "seconds": round(r["seconds"], 3), "csv_rows": rows,
                         "epochs_trained": trained, "csv_sha256_16": digest,
                         "error": (r["error"] or "").splitlines()[0][:200] if not r["ok"] else ""})
    n_ok = sum(r["ok"] for r in results)
    busy = {}
    for r in results:
        busy[r["worker"]] = busy.get(r["worker"], 0.0) + r["seconds"]
    all_digest = hashlib.sha256("".join(f"{c['training_run']}:{c['perturb_length']}:{c['csv_sha256_16']};"
                                        for c in sorted(per_cond, key=lambda c: (c["training_run"],
                                                                                 c["perturb_length"]))).encode()
                                ).hexdigest()[:16]
    out = {"metric": "perturbation sweep conditions/hour", "kind": a.kind, "layout": layout, "chain": bool(a.chain),
           "n_gpus": len(gpus), "workers_per_gpu": max(1, a.workers_per_gpu), "conditions": len(conds), "ok": n_ok, "failed": len(conds) - n_ok,
           "wall_s": wall, "conditions_per_hour": 3600.0 * n_ok / wall,
           "epochs_trained": epochs_total, "epochs_per_s": epochs_total / wall,
           "sec_per_epoch_per_gpu": wall * len(gpus) / max(1, epochs_total),
           "worker_busy_s": {str(k): round(v, 2) for k, v in sorted(busy.items(), key=lambda kv: str(kv[0]))},
           "worker_startup_and_idle_s": round(wall - max(busy.values()), 2) if busy else None,
           "balance": (sum(busy.values()) / len(busy)) / max(busy.values()) if busy else None,
           "baseline_epochs": base_epochs, "baseline_s": t_base, "dataset_s": t_data,
           "backbone": a.backbone, "batch_size": a.batch_size, "n_train_images": a.n_train,
           "early_stopping_patience": a.patience, "trajectories_digest": all_digest,
           "what": "hba.sweep.run_sweep: one spawned worker process per GPU (CUDA_VISIBLE_DEVICES), shared queue in "
                   "LPT order, every condition = run_behavioral_training(config) until the reference's early stopping "
                   "ends it; wall clock from the first process spawn to the last result, incl. worker start-up, "
                   "cache fill, graph capture, per-epoch eval + RSA + CSV + 2 checkpoint files",
           "per_condition": per_cond}
    txt = json.dumps(out)
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        with open(a.out, "w") as f:
            f.write(txt + "\n")
    brief = {k: v for k, v in out.items() if k != "per_condition"}
    print(json.dumps(brief))
    if not a.keep:
        shutil.rmtree(root, ignore_errors=True)
    sys.exit(0 if n_ok == len(conds) else 1)


if __name__ == "__main__":
    main()

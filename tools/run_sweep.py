#!/usr/bin/env python
"""Multi-GPU perturbation sweep driver (hba.sweep): one worker process per GPU, conditions handed out
longest-first, each condition = the unmodified run_behavioral_training(config).

  python tools/run_sweep.py --kind single --start 1 --end 98 --gpus 0,1,2,3,4,5,6,7 \
      --perturb-type random_target --baseline-dir /path/to/baseline_run --output /path/to/out \
      --csv-file ... --img-dir ... --inference-csv-file ... --rdm ...
  python tools/run_sweep.py --kind grid ...        # the 136-condition (start, length) grid
  python tools/run_sweep.py --kind grid --plan --gpus 0,1,2,3,4,5,6,7     # print the LPT plan only
  python tools/run_sweep.py --kind grid --chain ...  # LEN's resume chain: a start epoch's windows in increasing length
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vit-project_b200")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", choices=["single", "grid"], default="single")
    ap.add_argument("--start", type=int, default=1)
    ap.add_argument("--end", type=int, default=98)
    ap.add_argument("--gpus", default="0")
    ap.add_argument("--workers-per-gpu", type=int, default=1,
                    help="worker processes per GPU - measured on a B200: 2 / 4 workers run 2.2x / 4.5x SLOWER than one "
                         "(profiles/r02_grid12_workers*.json); keep 1")
    ap.add_argument("--plan", action="store_true")
    ap.add_argument("--chain", action="store_true",
                    help="grid only: LEN's shorter -> longer resume chain, one start epoch per worker at a time")
    ap.add_argument("--perturb-type", default="random_target")
    ap.add_argument("--perturb-distribution", default="target")
    ap.add_argument("--perturb-seed", type=int, default=42)
    ap.add_argument("--baseline-dir", help="baseline run directory holding dora_params/ and random_states/")
    ap.add_argument("--output")
    ap.add_argument("--csv-file")
    ap.add_argument("--img-dir")
    ap.add_argument("--inference-csv-file")
    ap.add_argument("--rdm")
    ap.add_argument("--epochs", type=int, default=500)
    ap.add_argument("--batch-size", type=int, default=32)
    ap.add_argument("--patience", type=int, default=20)
    a = ap.parse_args()
    from hba import sweep
    conds = sweep.single_epoch_conditions(a.start, a.end) if a.kind == "single" else sweep.length_grid_conditions()
    devices = [int(x) for x in a.gpus.split(",") for _ in range(max(1, a.workers_per_gpu))]
    if a.plan and a.chain:
        groups = sweep.chain_groups(conds)
        plan, loads = sweep.lpt_assign(groups, len(devices), cost=sweep.chain_cost)
        for w, (p, load) in enumerate(zip(plan, loads)):
            print(f"worker {w} (GPU {devices[w]}): {len(p)} start epochs, {load} expected epochs: "
                  + " ".join(f"e{g[0]['training_run']}x{len(g)}" for g in p))
        indep = sum(sweep.expected_epochs(c) for c in conds)
        print(f"{len(conds)} conditions in {len(groups)} chains; {sum(loads)} epochs (independent: {indep}); "
              f"makespan {max(loads)} vs ideal {sum(loads) / len(devices):.1f} epochs")
        return
    if a.plan:
        plan, loads = sweep.lpt_assign(conds, len(devices))
        for w, (p, load) in enumerate(zip(plan, loads)):
            print(f"worker {w} (GPU {devices[w]}): {len(p)} conditions, {load} expected epochs: "
                  + " ".join(f"({c['training_run']},{c['perturb_length']})" for c in p))
        print(f"{len(conds)} conditions; makespan {max(loads)} vs ideal {sum(loads) / len(devices):.1f} epochs")
        return
    import torch.nn as nn
    base = {"csv_file": a.csv_file, "img_dir": a.img_dir, "inference_csv_file": a.inference_csv_file,
            "RDM48_triplet_dir": a.rdm, "backbone": "ViT-L/14", "epochs": a.epochs, "batch_size": a.batch_size,
            "train_portion": 0.8, "lr": 3e-4, "logger": None, "early_stopping_patience": a.patience,
            "random_seed": 1, "vision_layers": 2, "transformer_layers": 1, "rank": 32, "criterion": nn.MSELoss(),
            "cuda": 0, "baseline_dora_directory": os.path.join(a.baseline_dir, "dora_params"),
            "baseline_random_state_path": os.path.join(a.baseline_dir, "random_states"),
            "baseline_split_indices_path": os.path.join(a.baseline_dir, "random_states", "dataset_split_indices.pth"),
            "perturb_type": a.perturb_type, "perturb_length": 1, "perturb_distribution": a.perturb_distribution,
            "perturb_seed": a.perturb_seed, "output_base_directory": a.output}
    results = sweep.run_sweep(base, conds, devices, layout="sweep" if a.kind == "single" else "length",
                              chain=a.chain and a.kind == "grid")
    sys.exit(0 if all(r["ok"] for r in results) else 1)


if __name__ == "__main__":
    main()

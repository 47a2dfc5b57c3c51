#!/usr/bin/env python
"""RSA score of every checkpoint of a ViT baseline run (hba.vit_train.rsa_over_checkpoints) -> the
`baseline_metrics_csv` of measure_single_epoch_perturbation_effect.py, in the schema of the reference's shipped
Data/vit_results/rsa_results_final.csv.  Checkpoints shard across ranks; no collective on the data path.

  torchrun --nproc_per_node=8 tools/vit_rsa_over_checkpoints.py --checkpoint_dir runs/vit_sgd \
      --things_csv <csv|synthetic> [--things_img_dir D --things_rdm_path RDM48_triplet.mat] --output_csv runs/vit_sgd/rsa_results.csv
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vit-project_b200")]


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--checkpoint_dir", required=True)
    ap.add_argument("--output_csv", required=True)
    ap.add_argument("--things_csv", required=True, help="THINGS inference CSV, or 'synthetic'")
    ap.add_argument("--things_img_dir", default="")
    ap.add_argument("--things_rdm_path", default="")
    ap.add_argument("--num_classes", type=int, default=1000)
    ap.add_argument("--model", default="vit_base_patch16_224", help=argparse.SUPPRESS)
    a = ap.parse_args(argv)
    import importlib.util
    import torch
    import torch.distributed as dist
    from hba import vit_train as vt
    spec = importlib.util.spec_from_file_location("_measure_script", os.path.join(
        ROOT, "vit-project_b200", "vit_training", "single_epoch", "measure_single_epoch_perturbation_effect.py"))
    measure = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(measure)
    rank, world_size, local_rank = vt.setup_distributed()
    device = torch.device("cuda", local_rank)
    things, rdm = measure.load_things(a.things_csv, a.things_img_dir, a.things_rdm_path, device)
    rows = vt.rsa_over_checkpoints(a.checkpoint_dir, things, rdm, a.output_csv, model_name=a.model,
                                   num_classes=a.num_classes, rank=rank, world_size=world_size)
    if rank == 0:
        print(f"{len(rows)} checkpoints -> {a.output_csv}")
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

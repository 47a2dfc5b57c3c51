#!/usr/bin/env python3
"""Immediate effect of single-epoch perturbations on ViT training, on libhba (drop-in for the reference's
Training/vit_training/single_epoch/measure_single_epoch_perturbation_effect.py: same command line, same
result CSV).

For each perturbation epoch N and type (gaussian, uniform_gray, label_shuffle, target_noise): load the
baseline checkpoint of epoch N-1, train ONLY epoch N on the perturbed data, evaluate, compute the RSA of the
CLS features on the 48 THINGS images, and record the differences to the baseline run's epoch N.

`--data_path synthetic:NTRAIN:NVAL[:C]` and `--things_csv synthetic` replace ImageNet / THINGS by HBM-resident
synthetic stand-ins (there are no datasets offline).  `--reference_row_order` reproduces the reference's
rank-interleaved RSA rows for world sizes > 1; the default keeps dataset order.
"""
import argparse
import os
import sys

_PKG = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from hba import vit_train as vt  # noqa: E402
from hba.vit import CosineAnnealingLRWithWarmup  # noqa: E402,F401
from hba.vit_train import (GaussianNoiseTransform, ShuffledLabelsDataset, TargetNoiseDataset,  # noqa: E402,F401
                           UniformGrayTransform, compute_rsa_score, measure_perturbation_effect,
                           train_one_epoch, validate)


class THINGSInferenceDataset(torch.utils.data.Dataset):
    """MEAS:95-115: (image_name, transformed image) for every row of the THINGS inference CSV."""

    def __init__(self, csv_file, img_dir, rdm_path, transform=None):
        import pandas as pd
        self.data = pd.read_csv(csv_file)
        self.img_dir, self.rdm_path, self.transform = img_dir, rdm_path, transform

    def __len__(self):
        return len(self.data)

    def __getitem__(self, idx):
        from PIL import Image
        name = self.data.iloc[idx]["image_name"]
        image = Image.open(os.path.join(self.img_dir, name)).convert("RGB")
        if self.transform:
            image = self.transform(image)
        return name, image


def setup_distributed():
    return vt.setup_distributed()


def get_dataloaders(data_path, batch_size, num_workers, world_size, rank, perturbation_type=None, epsilon=0.1,
                    shuffle_seed=42, device=None):
    """MEAS:139-227 -> (train_loader, val_loader, train_sampler), the training data perturbed as requested."""
    device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    syn = vt.parse_synthetic(data_path)
    if syn is None:
        return vt.imagenet_loaders(data_path, batch_size, num_workers, world_size, rank, device,
                                   perturbation_type=perturbation_type, epsilon=epsilon, shuffle_seed=shuffle_seed)
    n_train, n_val, classes = syn
    train = vt.synthetic_imagenet(n_train, classes, seed=0, device=device)
    val = vt.synthetic_imagenet(n_val, classes, seed=1, device=device)
    train_loader = vt.ShardedLoader(train, batch_size, world_size, rank, shuffle=True, perturbation_type=perturbation_type,
                                    epsilon=epsilon, shuffle_seed=shuffle_seed, num_classes=classes)
    val_loader = vt.ShardedLoader(val, batch_size, world_size, rank, shuffle=False, num_classes=classes)
    return train_loader, val_loader, train_loader.sampler


def load_things(things_csv, things_img_dir, things_rdm_path, device):
    """-> (ResidentImageSet of the RSA images, reference RDM)."""
    if str(things_csv).startswith("synthetic"):
        return vt.synthetic_things(device)
    from torchvision import transforms
    transform = transforms.Compose([transforms.Resize(256), transforms.CenterCrop(224), transforms.ToTensor(),
                                    transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    ds = THINGSInferenceDataset(things_csv, things_img_dir, things_rdm_path, transform)
    return vt.ResidentImageSet.from_dataset(ds, device), vt.load_reference_rdm(things_rdm_path)


def build_parser():
    p = argparse.ArgumentParser(description="Measure single-epoch perturbation effects on ViT")
    p.add_argument("--baseline_checkpoint_dir", type=str, required=True, help="Directory containing baseline checkpoints")
    p.add_argument("--baseline_metrics_csv", type=str, required=True, help="Baseline metrics CSV (epoch, val_loss, rsa_score)")
    p.add_argument("--data_path", type=str, required=True, help="Path to ImageNet data, or synthetic:NTRAIN:NVAL[:C]")
    p.add_argument("--output_csv", type=str, required=True, help="Output CSV file for results")
    p.add_argument("--things_csv", type=str, required=True, help="THINGS inference CSV, or 'synthetic'")
    p.add_argument("--things_img_dir", type=str, default="", help="Directory containing THINGS images")
    p.add_argument("--things_rdm_path", type=str, default="", help="Behavioural RDM .mat file")
    p.add_argument("--perturbation_types", type=str, nargs="+", default=list(vt.PERTURBATION_TYPES))
    p.add_argument("--perturb_epochs", type=int, nargs="+", default=list(vt.DEFAULT_PERTURB_EPOCHS))
    p.add_argument("--epsilon", type=float, default=0.1, help="Perturbation strength for gaussian noise")
    p.add_argument("--batch_size", type=int, default=256)
    p.add_argument("--lr", type=float, default=0.1)
    p.add_argument("--momentum", type=float, default=0.9)
    p.add_argument("--weight_decay", type=float, default=1e-4)
    p.add_argument("--warmup_epochs", type=int, default=5)
    p.add_argument("--total_epochs", type=int, default=100)
    p.add_argument("--num_workers", type=int, default=8)
    p.add_argument("--reference_row_order", action="store_true", help="rank-interleaved RSA rows as in the reference")
    p.add_argument("--noise_seed", type=int, default=None, help="seed of the device generator for gaussian noise")
    p.add_argument("--model", type=str, default="vit_base_patch16_224", help=argparse.SUPPRESS)
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    rank, world_size, local_rank = vt.setup_distributed()
    device = torch.device("cuda", local_rank)
    if rank == 0:
        print("=" * 80 + "\nViT Single-Epoch Perturbation Effect Measurement - libhba / sm_100a\n" + "=" * 80)
        print(f"Baseline checkpoint dir: {args.baseline_checkpoint_dir}\nBaseline metrics CSV: {args.baseline_metrics_csv}\n"
              f"Perturbation types: {args.perturbation_types}\nPerturbation epochs: {args.perturb_epochs}\n"
              f"Output CSV: {args.output_csv}")
    if world_size > 1:
        dist.barrier()
    syn = vt.parse_synthetic(args.data_path)
    if syn is None:      # a real ImageFolder tree: streamed through the reference's torchvision pipeline (MEAS:139-227)
        train = val = None
        classes, data_path = 1000, args.data_path
    else:                # synthetic stand-ins, kept resident in HBM
        n_train, n_val, classes = syn
        train = vt.synthetic_imagenet(n_train, classes, seed=0, device=device)
        val = vt.synthetic_imagenet(n_val, classes, seed=1, device=device)
        data_path = None
    things, rdm = load_things(args.things_csv, args.things_img_dir, args.things_rdm_path, device)
    from hba.rsa import RSAEvaluator
    evaluator = RSAEvaluator(rdm, device) if rank == 0 else None
    results = vt.measure_all(
        args.output_csv, args.perturb_epochs, args.perturbation_types, rank=rank,
        baseline_checkpoint_dir=args.baseline_checkpoint_dir, baseline_metrics_csv=args.baseline_metrics_csv,
        train_data=train, val_data=val, things_data=things, things_rdm=rdm, epsilon=args.epsilon,
        batch_size=args.batch_size, lr=args.lr, momentum=args.momentum, weight_decay=args.weight_decay,
        warmup_epochs=args.warmup_epochs, total_epochs=args.total_epochs, world_size=world_size,
        local_rank=local_rank, model_name=args.model, num_classes=classes,
        dataset_order=not args.reference_row_order, noise_seed=args.noise_seed, evaluator=evaluator,
        data_path=data_path, num_workers=args.num_workers)
    if rank == 0:
        import pandas as pd
        print(f"\nSaved results to {args.output_csv}\n\nResults summary:")
        print(pd.DataFrame(results).to_string(index=False))
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

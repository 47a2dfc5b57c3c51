#!/usr/bin/env python3
"""ViT-Base classification training with SGD on libhba (drop-in for the reference's
Training/vit_training/baseline/train_vit_sgd.py: same command line, checkpoint files and metrics CSV).

    torchrun --nproc_per_node=N train_vit_sgd.py --data_path <ImageFolder root | synthetic:NTRAIN:NVAL[:C]>
             --output_dir <dir> [--batch_size 256 --epochs 100 --lr 0.1 --momentum 0.9 --weight_decay 1e-4
             --num_workers 8 --warmup_epochs 5]

Differences from the reference, all on the device side: bf16 tensor-core step without GradScaler
(`hba.vit.DataParallelTrainer`, CUDA-graph captured, NCCL bucket all-reduce overlapped with the backward
pass), synthetic data kept resident in HBM, loss sums accumulated on the device.
"""
import argparse
import os
import sys

_PKG = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from hba import vit  # noqa: E402
from hba import vit_train as vt  # noqa: E402
from hba.vit import CosineAnnealingLRWithWarmup  # noqa: E402,F401  (reference name, VIT:206)
from hba.vit_train import save_checkpoint, train_one_epoch, validate  # noqa: E402,F401


def setup_distributed():
    return vt.setup_distributed()


def get_dataloaders(data_path, batch_size, num_workers, world_size, rank, device=None):
    """VIT:29-87 -> (train_loader, val_loader, train_sampler)."""
    device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    syn = vt.parse_synthetic(data_path)
    if syn is None:
        return vt.imagenet_loaders(data_path, batch_size, num_workers, world_size, rank, device)
    n_train, n_val, classes = syn
    train = vt.synthetic_imagenet(n_train, classes, seed=0, device=device)
    val = vt.synthetic_imagenet(n_val, classes, seed=1, device=device)
    train_loader = vt.ShardedLoader(train, batch_size, world_size, rank, shuffle=True, num_classes=classes)
    val_loader = vt.ShardedLoader(val, batch_size, world_size, rank, shuffle=False, num_classes=classes)
    return train_loader, val_loader, train_loader.sampler


def build_parser():
    parser = argparse.ArgumentParser(description="Train ViT-Base on ImageNet")
    parser.add_argument("--data_path", type=str, required=True, help="Path to ImageNet data, or synthetic:NTRAIN:NVAL[:C]")
    parser.add_argument("--output_dir", type=str, required=True, help="Output directory for checkpoints")
    parser.add_argument("--batch_size", type=int, default=256, help="Batch size per GPU")
    parser.add_argument("--epochs", type=int, default=100, help="Number of epochs")
    parser.add_argument("--lr", type=float, default=0.1, help="Learning rate")
    parser.add_argument("--momentum", type=float, default=0.9, help="SGD momentum")
    parser.add_argument("--weight_decay", type=float, default=1e-4, help="Weight decay")
    parser.add_argument("--num_workers", type=int, default=8, help="Number of data loading workers")
    parser.add_argument("--warmup_epochs", type=int, default=5, help="Warmup epochs")
    parser.add_argument("--model", type=str, default="vit_base_patch16_224", help=argparse.SUPPRESS)
    return parser


def main(argv=None):
    args = build_parser().parse_args(argv)
    rank, world_size, local_rank = setup_distributed()
    device = torch.device("cuda", local_rank)
    say = print if rank == 0 else (lambda *_: None)
    say("=" * 60 + "\nViT-Base ImageNet Training (SGD) - libhba / sm_100a\n" + "=" * 60)
    say(f"World size: {world_size} GPUs\nBatch size per GPU: {args.batch_size}\n"
        f"Effective batch size: {args.batch_size * world_size}\nTotal epochs: {args.epochs}\n"
        f"Learning rate: {args.lr}\nMomentum: {args.momentum}\nWeight decay: {args.weight_decay}\n"
        f"Warmup epochs: {args.warmup_epochs}\nOutput directory: {args.output_dir}")
    if rank == 0:
        os.makedirs(args.output_dir, exist_ok=True)
    if world_size > 1:
        dist.barrier()

    syn = vt.parse_synthetic(args.data_path)
    num_classes = syn[2] if syn is not None else 1000
    torch.manual_seed(0)   # every rank builds the same initial model; rank 0's is broadcast below anyway
    model = vit.create_model(args.model, pretrained=False, num_classes=num_classes).to(device)
    say(f"Model created. Parameters: {sum(p.numel() for p in model.parameters()) / 1e6:.1f}M")
    trainer = vit.DataParallelTrainer(model, lr=args.lr, momentum=args.momentum, weight_decay=args.weight_decay,
                                      use_graph=True)
    scheduler = CosineAnnealingLRWithWarmup(trainer, warmup_epochs=args.warmup_epochs, max_epochs=args.epochs,
                                            eta_min=0)
    train_loader, val_loader, train_sampler = get_dataloaders(args.data_path, args.batch_size, args.num_workers,
                                                              world_size, rank, device)
    say(f"Data loaded. Train batches: {len(train_loader)}, Val batches: {len(val_loader)}")

    start_epoch = 0
    checkpoint_path = os.path.join(args.output_dir, "checkpoint_latest.pth")
    if os.path.exists(checkpoint_path):                                       # VIT:314-328
        checkpoint = vt.load_checkpoint(checkpoint_path, model, trainer, scheduler, device)
        start_epoch = checkpoint["epoch"] + 1
        say(f"Resumed from epoch {checkpoint['epoch']}")
    else:
        trainer.broadcast_parameters()                                        # DDP's initial broadcast, VIT:287

    for epoch in range(start_epoch, args.epochs):
        say(f"\n{'=' * 60}\nEpoch {epoch}/{args.epochs - 1}\n{'=' * 60}")
        train_sampler.set_epoch(epoch)
        train_loss = train_one_epoch(trainer, train_loader, epoch, local_rank, world_size, log=say)
        scheduler.step()
        val_loss, val_acc = validate(trainer, val_loader, local_rank, world_size)
        say(f"Epoch {epoch}: train_loss={train_loss:.4f} val_loss={val_loss:.4f} val_acc={val_acc:.2f}%")
        save_checkpoint(epoch, model, trainer, scheduler, train_loss, val_loss, val_acc, args.output_dir, local_rank)
    say("\n" + "=" * 60 + "\nTraining Complete!\n" + "=" * 60)
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""ORACLE (test infrastructure only - never imported by the product path).

CPU restatement of the reference's ViT epoch loop and single-epoch perturbation measurement:
  VIT  = Training/vit_training/baseline/train_vit_sgd.py
  MEAS = Training/vit_training/single_epoch/measure_single_epoch_perturbation_effect.py
in fp32 without autocast / GradScaler, for world size 1 and, for the collective tails, any world size as plain
functions of the per-rank values.  The model is oracle/vit_ref.py (timm restated).

Pinned against the reference's own code: `tests/golden/vit_measure.json` (written by
oracle/make_vit_measure_golden.py with the classes and functions imported from MEAS / VIT, `timm` stubbed)
holds the label perturbations, the schedule, the checkpoint / CSV formats and the shipped result rows of
Data/vit_results; `tests/test_vit_measure_cpu.py::test_restatement_equals_reference_classes` compares
directly where /root/reference is mounted.  The model-dependent parts (train / validate / RSA on a real
network) cannot be run from the reference offline (hard-wired `.cuda()`, NCCL, timm): they are restated
here line by line, and their NumPy / SciPy tail is the reference's own arithmetic.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F
from scipy.stats import spearmanr
from torch.utils.data import DataLoader, Dataset, DistributedSampler, TensorDataset


class ShuffledLabelsRef(Dataset):
    """MEAS:57-72: label of sample i = label of sample RandomState(seed).permutation(n)[i]."""

    def __init__(self, base, shuffle_seed=42):
        self.base = base
        self.shuffled_indices = np.random.RandomState(shuffle_seed).permutation(len(base))

    def __len__(self):
        return len(self.base)

    def __getitem__(self, i):
        return self.base[i][0], self.base[self.shuffled_indices[i]][1]


class TargetNoiseRef(Dataset):
    """MEAS:74-93: label of sample i = RandomState(seed).randint(0, C, n)[i]."""

    def __init__(self, base, num_classes=1000, noise_seed=42):
        self.base = base
        self.random_targets = np.random.RandomState(noise_seed).randint(0, num_classes, len(base))

    def __len__(self):
        return len(self.base)

    def __getitem__(self, i):
        return self.base[i][0], int(self.random_targets[i])


class ImagePerturbRef(Dataset):
    """MEAS:36-55 applied to already-transformed tensors: 'gaussian' -> randn_like * eps, 'uniform_gray' -> 0."""

    def __init__(self, base, kind, epsilon=0.1):
        self.base, self.kind, self.epsilon = base, kind, epsilon

    def __len__(self):
        return len(self.base)

    def __getitem__(self, i):
        img, y = self.base[i]
        if self.kind == "gaussian":
            return torch.randn_like(img) * self.epsilon, y
        return torch.zeros_like(img), y


def perturbed_dataset(images, labels, perturbation_type, epsilon=0.1, num_classes=1000, shuffle_seed=42):
    """MEAS:139-181 on tensors."""
    ds = TensorDataset(images, labels)
    if perturbation_type in ("gaussian", "uniform_gray"):
        ds = ImagePerturbRef(ds, perturbation_type, epsilon)
    elif perturbation_type == "label_shuffle":
        ds = ShuffledLabelsRef(ds, shuffle_seed)
    elif perturbation_type == "target_noise":
        ds = TargetNoiseRef(ds, num_classes, shuffle_seed)
    return ds


def rank_loader(dataset, batch_size, world_size, rank, shuffle, epoch=None):
    """VIT:58-84: DistributedSampler + DataLoader (no workers)."""
    sampler = DistributedSampler(dataset, num_replicas=world_size, rank=rank, shuffle=shuffle)
    if epoch is not None:
        sampler.set_epoch(epoch)
    return DataLoader(dataset, batch_size=batch_size, sampler=sampler)


class CosineWarmupRef:
    """VIT:206-244."""

    def __init__(self, optimizer, warmup_epochs, max_epochs, eta_min=0):
        self.optimizer, self.warmup_epochs, self.max_epochs, self.eta_min = optimizer, warmup_epochs, max_epochs, eta_min
        self.base_lrs = [g["lr"] for g in optimizer.param_groups]
        self.current_epoch = 0

    def step(self):
        for g, base in zip(self.optimizer.param_groups, self.base_lrs):
            if self.current_epoch < self.warmup_epochs:
                g["lr"] = base * ((self.current_epoch + 1) / self.warmup_epochs)
            else:
                progress = (self.current_epoch - self.warmup_epochs) / (self.max_epochs - self.warmup_epochs)
                g["lr"] = self.eta_min + (base - self.eta_min) * 0.5 * (1 + math.cos(math.pi * progress))
        self.current_epoch += 1

    def state_dict(self):
        return {"current_epoch": self.current_epoch, "base_lrs": self.base_lrs, "warmup_epochs": self.warmup_epochs,
                "max_epochs": self.max_epochs, "eta_min": self.eta_min}

    def load_state_dict(self, sd):
        self.current_epoch, self.base_lrs = sd["current_epoch"], sd["base_lrs"]
        self.warmup_epochs, self.max_epochs, self.eta_min = sd["warmup_epochs"], sd["max_epochs"], sd["eta_min"]


def train_one_epoch_ref(model, loader, optimizer):
    """VIT:125-153 (one rank): mean over batches of the batch-mean CE; SGD step per batch."""
    model.train()
    total, n = 0.0, 0
    for images, targets in loader:
        optimizer.zero_grad()
        loss = F.cross_entropy(model(images), targets)
        loss.backward()
        optimizer.step()
        total += loss.item()
        n += 1
    return total / n


def reduce_train_loss(per_rank_avg):
    """VIT:154-157: fp32 SUM all-reduce of the per-rank means, divided by the world size."""
    s = torch.zeros((), dtype=torch.float32)
    for v in per_rank_avg:
        s = s + torch.tensor(v)
    return s.item() / len(per_rank_avg)


def validate_rank_ref(model, loader):
    """VIT:167-192 (one rank) -> [avg_loss, accuracy, total, correct]."""
    model.eval()
    correct = total = n = 0
    val_loss = 0.0
    with torch.no_grad():
        for images, targets in loader:
            out = model(images)
            val_loss += F.cross_entropy(out, targets).item()
            n += 1
            total += targets.size(0)
            correct += out.max(1)[1].eq(targets).sum().item()
    return [val_loss / n, 100.0 * correct / total, total, correct]


def reduce_validation(per_rank_metrics):
    """VIT:193-201: fp32 SUM all-reduce of the 4-vectors; the loss is metrics[0] (the SUM of the per-rank means:
    never divided by the world size), the accuracy comes from the global counts."""
    m = torch.zeros(4, dtype=torch.float32)
    for v in per_rank_metrics:
        m = m + torch.tensor(v, dtype=torch.float32)
    return m[0].item(), 100.0 * int(m[3].item()) / int(m[2].item())


def rsa_tail_ref(embeddings, reference_rdm):
    """MEAS:340-353."""
    model_rdm = 1 - np.corrcoef(embeddings)
    np.fill_diagonal(model_rdm, 0)
    iu = np.triu_indices_from(reference_rdm, k=1)
    rho, p = spearmanr(reference_rdm[iu], model_rdm[iu])
    return rho, p


def rank_embeddings_ref(model, images, world_size, rank, batch_size=8):
    """MEAS:304-324 for one rank: CLS row of forward_features over the rank's DistributedSampler(shuffle=False) share."""
    model.eval()
    ds = TensorDataset(images, torch.zeros(len(images), dtype=torch.long))
    out = []
    with torch.no_grad():
        for x, _ in rank_loader(ds, batch_size, world_size, rank, shuffle=False):
            out.extend(model.forward_features(x)[:, 0].cpu().numpy())
    return np.array(out)


def compute_rsa_score_ref(model, images, reference_rdm, world_size=1, dataset_order=False):
    """MEAS:298-355.  world_size > 1: `cat(all_gather)[:48]` - rank-major rows, i.e. the reference's interleave
    (dataset_order=False), or the rows restored to dataset order (what the product does by default)."""
    blocks = [rank_embeddings_ref(model, images, world_size, r) for r in range(world_size)]
    if world_size == 1:
        emb = blocks[0]
    elif dataset_order:
        emb = np.stack(blocks, axis=1).reshape(-1, blocks[0].shape[1])[:len(reference_rdm)]
    else:
        emb = np.concatenate(blocks, axis=0)[:len(reference_rdm)]
    return rsa_tail_ref(emb, reference_rdm)


def measure_ref(checkpoint, model_factory, train, val, things_images, reference_rdm, perturb_epoch,
                perturbation_type, baseline_loss, baseline_rsa, epsilon=0.1, batch_size=256, lr=0.1, momentum=0.9,
                weight_decay=1e-4, warmup_epochs=5, total_epochs=100, num_classes=1000):
    """MEAS:403-555 at world size 1 on CPU tensors.  `train` / `val`: (images, labels)."""
    model = model_factory()
    model.load_state_dict(checkpoint["model_state_dict"])
    opt = torch.optim.SGD(model.parameters(), lr=lr, momentum=momentum, weight_decay=weight_decay)
    sched = CosineWarmupRef(opt, warmup_epochs, total_epochs, eta_min=0)
    opt.load_state_dict(checkpoint["optimizer_state_dict"])
    sched.load_state_dict(checkpoint["scheduler_state_dict"])
    ds = perturbed_dataset(train[0], train[1], perturbation_type, epsilon, num_classes, 42)
    train_loss = train_one_epoch_ref(model, rank_loader(ds, batch_size, 1, 0, True, epoch=perturb_epoch), opt)
    sched.step()
    val_loss, _ = reduce_validation([validate_rank_ref(model, rank_loader(TensorDataset(*val), batch_size, 1, 0, False))])
    rho, _ = compute_rsa_score_ref(model, things_images, reference_rdm)
    return {"perturb_epoch": perturb_epoch, "perturbation_type": perturbation_type,
            "baseline_loss": baseline_loss, "baseline_rsa": baseline_rsa, "perturbed_loss": val_loss,
            "perturbed_rsa": rho, "delta_loss": val_loss - baseline_loss, "delta_rsa": rho - baseline_rsa,
            "train_loss": train_loss, "lr_after": opt.param_groups[0]["lr"], "model": model, "optimizer": opt}

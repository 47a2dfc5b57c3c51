"""Epoch-level host side of the ViT-B/16 study: the training script's epoch loop (reference VIT =
Training/vit_training/baseline/train_vit_sgd.py) and the single-epoch perturbation measurement
(reference MEAS = Training/vit_training/single_epoch/measure_single_epoch_perturbation_effect.py).

Same names, argument meaning and on-disk formats as the reference functions they replace:

    GaussianNoiseTransform / UniformGrayTransform      MEAS:36-55   (PIL pipeline wrappers)
    ShuffledLabelsDataset / TargetNoiseDataset         MEAS:57-93   (label perturbations, NumPy RandomState)
    train_one_epoch / validate / save_checkpoint       VIT:89-204   (MEAS:229-296)
    compute_rsa_score                                  MEAS:298-355
    measure_perturbation_effect / measure_all          MEAS:403-555, 625-651

What changes is where the work happens.  The reference decodes JPEGs in DataLoader workers and synchronises
on `loss.item()` every step; here the epoch's images live in HBM (`ResidentImageSet`), the perturbations are
applied to whole batches on the device, a rank's batches are cut by the same `DistributedSampler` index
stream, every step is the fused / graph-captured `hba.vit.DataParallelTrainer.step`, running sums stay on
the device (float64, the same summation order as the reference's Python floats) and are read once per
epoch, and the RSA tail (RDM, ranking, Spearman) runs in libhba (`hba.rsa.RSAEvaluator`).

Two reference quirks are kept behind explicit switches:
  * `validate(..., reference_rank_sum=True)`: the reported validation loss is the SUM over ranks of the
    per-rank mean losses (VIT:196 reads metrics[0] after a SUM all-reduce) - kept by default so that a
    baseline CSV written here is comparable with the shipped one; False gives the mean.
  * `compute_rsa_score(..., dataset_order=True)`: with W > 1 ranks the reference concatenates the gathered
    per-rank embeddings (rank-strided by `DistributedSampler(shuffle=False)`) without undoing the stride,
    so the model RDM's rows are misaligned with the human RDM (MEAS:326-334).  Default here: rows restored
    to dataset order (the value every rank count agrees on); False reproduces the reference's row order.

This module never touches a CPU model path: the trainer and model it drives raise on non-CUDA tensors.
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch
from torch.utils.data import Dataset, DistributedSampler

PERTURBATION_TYPES = ("gaussian", "uniform_gray", "label_shuffle", "target_noise")      # MEAS:577
DEFAULT_PERTURB_EPOCHS = (5, 10, 15, 16, 20, 25, 30, 35, 45, 70, 98)                      # MEAS:580
RESULT_COLUMNS = ("perturb_epoch", "perturbation_type", "baseline_loss", "baseline_rsa", "perturbed_loss",
                  "perturbed_rsa", "delta_loss", "delta_rsa")                             # MEAS:541-550
GRAD_SCALER_STATE = {"scale": 65536.0, "growth_factor": 2.0, "backoff_factor": 0.5, "growth_interval": 2000,
                     "_growth_tracker": 0}   # a fresh GradScaler's state_dict: bf16 training needs no scaler,
#                                              the key stays in the checkpoint for the reference's loader (VIT:323)


# ------------------------------------------------------------------------------------------------
# Perturbations: the reference's wrapper classes (PIL / Dataset level) ...
# ------------------------------------------------------------------------------------------------
class GaussianNoiseTransform:
    """MEAS:36-45: the image is REPLACED by N(0, epsilon^2) noise of the transformed image's shape."""

    def __init__(self, base_transform, epsilon=0.1):
        self.base_transform, self.epsilon = base_transform, epsilon

    def __call__(self, img):
        img = self.base_transform(img)
        return torch.randn_like(img) * self.epsilon


class UniformGrayTransform:
    """MEAS:47-55: the normalised image is replaced by zeros (= the dataset mean colour)."""

    def __init__(self, base_transform):
        self.base_transform = base_transform

    def __call__(self, img):
        return torch.zeros_like(self.base_transform(img))


def shuffled_label_indices(num_samples, shuffle_seed=42):
    """MEAS:63-64: sample i takes the label of sample perm[i], perm = RandomState(seed).permutation(n)."""
    return np.random.RandomState(shuffle_seed).permutation(num_samples)


def random_targets(num_samples, num_classes=1000, noise_seed=42):
    """MEAS:83-84: one uniform random class per sample, RandomState(seed).randint(0, C, n)."""
    return np.random.RandomState(noise_seed).randint(0, num_classes, num_samples)


class ShuffledLabelsDataset(Dataset):
    """MEAS:57-72."""

    def __init__(self, base_dataset, shuffle_seed=42):
        self.base_dataset = base_dataset
        self.num_samples = len(base_dataset)
        self.shuffled_indices = shuffled_label_indices(self.num_samples, shuffle_seed)

    def __len__(self):
        return self.num_samples

    def __getitem__(self, idx):
        img, _ = self.base_dataset[idx]
        _, shuffled_label = self.base_dataset[self.shuffled_indices[idx]]
        return img, shuffled_label


class TargetNoiseDataset(Dataset):
    """MEAS:74-93."""

    def __init__(self, base_dataset, num_classes=1000, noise_seed=42):
        self.base_dataset, self.num_classes = base_dataset, num_classes
        self.random_targets = random_targets(len(base_dataset), num_classes, noise_seed)

    def __len__(self):
        return len(self.base_dataset)

    def __getitem__(self, idx):
        img, _ = self.base_dataset[idx]
        return img, self.random_targets[idx]


# ------------------------------------------------------------------------------------------------
# ... and their HBM-resident form
# ------------------------------------------------------------------------------------------------
def perturbed_labels(labels, perturbation_type, num_classes=1000, seed=42):
    """The label vector a perturbed epoch trains on, for ALL samples at once: exactly the labels
    `ShuffledLabelsDataset` / `TargetNoiseDataset` return sample by sample (MEAS:178-181)."""
    if perturbation_type == "label_shuffle":
        perm = torch.from_numpy(shuffled_label_indices(len(labels), seed)).to(labels.device)
        return labels[perm]
    if perturbation_type == "target_noise":
        return torch.from_numpy(random_targets(len(labels), num_classes, seed)).to(labels.device, labels.dtype)
    return labels


class ResidentImageSet:
    """`images [N,3,H,W]` (already transformed: what the reference's transform pipeline yields) and
    `labels [N]` int64 on one device.  A fixed image set (synthetic ImageNet-shaped data, the 48 THINGS
    images) is decoded once and stays in HBM instead of being re-decoded every epoch."""

    def __init__(self, images, labels=None, names=None):
        if images.ndim != 4:
            raise ValueError("images must be [N, 3, H, W]")
        if labels is not None and len(labels) != len(images):
            raise ValueError("one label per image")
        self.images = images
        self.labels = labels.long() if labels is not None else None
        self.names = list(names) if names is not None else [f"image_{i:05d}" for i in range(len(images))]

    def __len__(self):
        return self.images.shape[0]

    @classmethod
    def from_dataset(cls, dataset, device):
        """Materialises a map-style dataset of (image tensor, label) or (name, image tensor) items."""
        imgs, second, names = [], [], None
        for i in range(len(dataset)):
            a, b = dataset[i]
            if isinstance(a, str):      # THINGSInferenceDataset yields (img_name, image), MEAS:114
                names = (names or []) + [a]
                imgs.append(b)
            else:
                imgs.append(a)
                second.append(int(b))
        images = torch.stack(imgs).to(device)
        labels = torch.tensor(second, dtype=torch.long, device=device) if second else None
        return cls(images, labels, names)


def synthetic_imagenet(n, num_classes=1000, seed=0, device="cuda", img_size=224):
    """`randn(N,3,224,224)` images (stand-ins for normalised crops) + `randint` labels."""
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(n, 3, img_size, img_size, generator=g)
    labels = torch.randint(0, num_classes, (n,), generator=g)
    return ResidentImageSet(images.to(device), labels.to(device))


class _Range(Dataset):
    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return i


class ShardedLoader:
    """One rank's batches of a `ResidentImageSet`, in the order `DataLoader(dataset, batch_size,
    sampler=DistributedSampler(dataset, world, rank, shuffle))` visits them (VIT:58-84): the sampler object
    is torch's own (`.sampler.set_epoch(e)` re-seeds the shuffle, VIT:343), only the gather of a batch is a
    device-side `index_select` instead of worker processes.

    `perturbation_type` (MEAS:139-181): 'gaussian' replaces every training batch by N(0, eps^2) noise drawn
    from `noise_generator` on the images' device, 'uniform_gray' by zeros, 'label_shuffle' / 'target_noise'
    swap the label vector (`perturbed_labels`)."""

    def __init__(self, data: ResidentImageSet, batch_size, world_size=1, rank=0, shuffle=False, seed=0,
                 perturbation_type=None, epsilon=0.1, shuffle_seed=42, num_classes=1000, noise_generator=None,
                 with_names=False):
        if perturbation_type not in (None, "none") + PERTURBATION_TYPES:
            raise ValueError(f"unknown perturbation_type {perturbation_type!r}")
        self.data, self.batch_size = data, batch_size
        self.perturbation_type = None if perturbation_type == "none" else perturbation_type
        self.epsilon, self.noise_generator, self.with_names = epsilon, noise_generator, with_names
        self.sampler = DistributedSampler(_Range(len(data)), num_replicas=world_size, rank=rank, shuffle=shuffle,
                                          seed=seed)
        self.labels = (perturbed_labels(data.labels, self.perturbation_type, num_classes, shuffle_seed)
                       if data.labels is not None else None)

    def __len__(self):
        return math.ceil(len(self.sampler) / self.batch_size)

    def __iter__(self):
        order = torch.tensor(list(self.sampler), dtype=torch.long)
        dev = self.data.images.device
        order_dev = order.to(dev)
        for s in range(0, len(order), self.batch_size):
            idx = order_dev[s:s + self.batch_size]
            if self.perturbation_type == "gaussian":
                shape = (len(idx),) + tuple(self.data.images.shape[1:])
                images = torch.randn(shape, device=dev, dtype=self.data.images.dtype,
                                     generator=self.noise_generator) * self.epsilon
            elif self.perturbation_type == "uniform_gray":
                images = torch.zeros((len(idx),) + tuple(self.data.images.shape[1:]), device=dev,
                                     dtype=self.data.images.dtype)
            else:
                images = self.data.images.index_select(0, idx)
            if self.with_names:
                yield [self.data.names[i] for i in order[s:s + self.batch_size].tolist()], images
            else:
                yield images, self.labels.index_select(0, idx)


# ------------------------------------------------------------------------------------------------
# Distributed helpers (device-agnostic: `gloo` on CPU tensors in the tests, `nccl` on the GPU)
# ------------------------------------------------------------------------------------------------
def _dist():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized()) else None


def _all_reduce_sum(t):
    d = _dist()
    if d is not None and d.get_world_size() > 1:
        d.all_reduce(t, op=d.ReduceOp.SUM)
    return t


def gather_embeddings(local, world_size, dataset_order=True, n_total=None):
    """all_gather of the per-rank embedding blocks [n_local, D] (MEAS:326-334).  Rank r holds the samples
    r, r+W, r+2W, ... (`DistributedSampler(shuffle=False)`, padded by wrapping to a multiple of W);
    `dataset_order` puts row j of rank r back at position j*W + r, otherwise the blocks are concatenated
    rank after rank as the reference does.  The first `n_total` rows are returned (reference: 48)."""
    if world_size > 1:
        d = _dist()
        if d is None:
            raise RuntimeError("world_size > 1 needs an initialised process group")
        blocks = [torch.zeros_like(local) for _ in range(world_size)]
        d.all_gather(blocks, local.contiguous())
    else:
        blocks = [local]
    return arrange_rank_blocks(blocks, dataset_order, n_total)


def arrange_rank_blocks(blocks, dataset_order=True, n_total=None):
    """Per-rank blocks [n_local, D] of a `DistributedSampler(shuffle=False)` pass -> one matrix.  Rank r's row j is
    sample j*W + r of the padded index list (samples wrapped from the start fill the last positions), so
    interleaving the blocks and keeping the first `n_total` rows restores dataset order exactly, also when W does
    not divide the number of samples; `dataset_order=False` is the reference's rank-major `torch.cat`."""
    if dataset_order:
        out = torch.stack(blocks, dim=1).reshape(-1, blocks[0].shape[1])     # row j*W + r = blocks[r][j]
    else:
        out = torch.cat(blocks, dim=0)
    return out if n_total is None else out[:n_total]


# ------------------------------------------------------------------------------------------------
# Epoch loop (VIT:125-204 = MEAS:229-296)
# ------------------------------------------------------------------------------------------------
def train_one_epoch(trainer, train_loader, epoch, local_rank=0, world_size=1, log=print, log_every=100):
    """VIT:125-165.  `trainer.step(images, targets) -> (loss, hits)` device tensors; the per-step
    `loss.item()` of the reference is a float64 device accumulation read once at the end (the progress line
    every `log_every` batches is the only other read).  Returns the mean over batches, averaged over ranks."""
    total, num_batches = None, 0
    for batch_idx, (images, targets) in enumerate(train_loader):
        loss, _ = trainer.step(images, targets)
        step_loss = loss.detach().double().reshape(())
        total = step_loss.clone() if total is None else total + step_loss
        num_batches += 1
        if log_every and batch_idx % log_every == 0 and local_rank == 0 and log is not None:
            log(f"  [{batch_idx:4d}/{len(train_loader)}] Loss: {float(loss):.4f}")
    if num_batches == 0:
        raise ValueError("train_one_epoch: empty loader")
    avg = (total / num_batches).reshape(1).float()       # torch.tensor(avg_loss) is fp32 in the reference
    _all_reduce_sum(avg)
    return float(avg) / world_size


def validate(trainer, val_loader, local_rank=0, world_size=1, reference_rank_sum=True):
    """VIT:167-204 -> (loss, top-1 accuracy in %).  `trainer.evaluate(images, targets) -> (mean CE, hits)`.
    One SUM all-reduce of [avg_loss, accuracy, total, correct]; accuracy from the global counts.  With
    `reference_rank_sum` the loss is the reference's: the sum over ranks of the per-rank mean."""
    loss_sum, correct, total, num_batches = None, None, 0, 0
    for images, targets in val_loader:
        loss, hits = trainer.evaluate(images, targets)
        l, h = loss.detach().double().reshape(()), hits.detach().long().reshape(())
        loss_sum = l.clone() if loss_sum is None else loss_sum + l
        correct = h.clone() if correct is None else correct + h
        total += int(targets.shape[0])
        num_batches += 1
    if num_batches == 0:
        raise ValueError("validate: empty loader")
    correct_n = int(correct)
    avg_loss = float(loss_sum) / num_batches
    metrics = torch.tensor([avg_loss, 100.0 * correct_n / total, total, correct_n], dtype=torch.float32,
                           device=loss_sum.device)
    _all_reduce_sum(metrics)
    m = metrics.tolist()
    global_loss = m[0] if reference_rank_sum else m[0] / world_size
    return global_loss, 100.0 * int(m[3]) / int(m[2])


def save_checkpoint(epoch, model, trainer, scheduler, train_loss, val_loss, val_acc, output_dir, local_rank=0,
                    scaler_state=None):
    """VIT:89-123: `checkpoint_epoch_{epoch:03d}.pth` + `checkpoint_latest.pth` with the reference's keys,
    and one `training_metrics.csv` row (`epoch,train_loss,val_loss,val_acc`, 6 / 6 / 4 decimals)."""
    if local_rank != 0:
        return None
    os.makedirs(output_dir, exist_ok=True)
    checkpoint = {
        "epoch": epoch,
        "model_state_dict": {k: v.detach().cpu() for k, v in model.state_dict().items()},
        "optimizer_state_dict": _to_cpu(trainer.state_dict()),
        "scheduler_state_dict": scheduler.state_dict(),
        "scaler_state_dict": dict(scaler_state if scaler_state is not None else GRAD_SCALER_STATE),
        "train_loss": train_loss, "val_loss": val_loss, "val_acc": val_acc,
    }
    path = os.path.join(output_dir, f"checkpoint_epoch_{epoch:03d}.pth")
    torch.save(checkpoint, path)
    torch.save(checkpoint, os.path.join(output_dir, "checkpoint_latest.pth"))
    csv_path = os.path.join(output_dir, "training_metrics.csv")
    if not os.path.exists(csv_path):
        with open(csv_path, "w") as f:
            f.write("epoch,train_loss,val_loss,val_acc\n")
    with open(csv_path, "a") as f:
        f.write(f"{epoch},{train_loss:.6f},{val_loss:.6f},{val_acc:.4f}\n")
    return path


def _to_cpu(obj):
    if torch.is_tensor(obj):
        return obj.detach().cpu()
    if isinstance(obj, dict):
        return {k: _to_cpu(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_cpu(v) for v in obj)
    return obj


def load_checkpoint(path, model, trainer, scheduler, device):
    """VIT:316-324 / MEAS:497-509: model, optimizer and scheduler state of a checkpoint written by
    `save_checkpoint` or by the reference script (the GradScaler state is read and ignored: bf16)."""
    checkpoint = torch.load(path, map_location=device, weights_only=False)
    model.load_state_dict(checkpoint["model_state_dict"])
    trainer.load_state_dict(checkpoint["optimizer_state_dict"])
    scheduler.load_state_dict(checkpoint["scheduler_state_dict"])
    return checkpoint


# ------------------------------------------------------------------------------------------------
# RSA of the ViT (MEAS:298-355)
# ------------------------------------------------------------------------------------------------
def load_reference_rdm(rdm):
    """`scipy.io.loadmat(rdm_path)['RDM48_triplet']` (MEAS:344-345), or an array passed through."""
    if isinstance(rdm, (str, os.PathLike)):
        import scipy.io
        return np.asarray(scipy.io.loadmat(rdm)["RDM48_triplet"], dtype=np.float64)
    return np.asarray(rdm, dtype=np.float64)


def pooled_features(model, images):
    """MEAS:308-322: `forward_features`, then the CLS token (or the mean of the patch tokens when the model
    says `global_pool == 'avg'`)."""
    m = model.module if hasattr(model, "module") else model
    features = m.forward_features(images)
    if getattr(m, "global_pool", None) == "avg":
        return features[:, 1:].mean(dim=1)
    return features[:, 0]


def compute_rsa_score(model, things_loader, rdm, local_rank=0, world_size=1, dataset_order=True, evaluator=None):
    """MEAS:298-355 -> (rho, p_value) on rank 0, (None, None) elsewhere.  `things_loader` yields
    (names, images) batches of this rank's share of the RSA images; the embeddings never leave the device:
    gather -> RDM (float64) -> average-tie ranks -> Spearman in libhba (`hba.rsa.RSAEvaluator`, pass
    `evaluator` to reuse the ranked reference RDM across measurements)."""
    with torch.no_grad():                                                     # MEAS:304
        feats = [pooled_features(model, images).float() for _, images in things_loader]
    if not feats:
        raise ValueError("compute_rsa_score: empty loader")
    local = torch.cat(feats, dim=0)
    reference = None
    if evaluator is None:
        reference = load_reference_rdm(rdm)
        n_total = reference.shape[0]
    else:
        n_total = evaluator.N
    emb = gather_embeddings(local, world_size, dataset_order=dataset_order, n_total=n_total)
    if local_rank != 0:
        return None, None
    if evaluator is None:
        from .rsa import RSAEvaluator
        evaluator = RSAEvaluator(reference, emb.device)
    rho, p_value, _ = evaluator(emb)
    return rho, p_value


# ------------------------------------------------------------------------------------------------
# The measurement (MEAS:403-555) and its driver loop (MEAS:625-662)
# ------------------------------------------------------------------------------------------------
def baseline_row(baseline_metrics, perturb_epoch):
    """MEAS:421-433: `val_loss` and `rsa_score` of the baseline run at `perturb_epoch`, or None.
    `baseline_metrics`: path of the CSV or a DataFrame."""
    import pandas as pd
    df = pd.read_csv(baseline_metrics) if isinstance(baseline_metrics, (str, os.PathLike)) else baseline_metrics
    row = df[df["epoch"] == perturb_epoch]
    if row.empty:
        return None
    return row["val_loss"].values[0], row["rsa_score"].values[0]


def assemble_result(perturb_epoch, perturbation_type, baseline_loss, baseline_rsa, val_loss, rsa_score):
    """MEAS:531-550: a missing RSA score (non-zero rank) counts as 0.0; deltas are perturbed - baseline."""
    if rsa_score is None:
        rsa_score = 0.0
    return {"perturb_epoch": perturb_epoch, "perturbation_type": perturbation_type,
            "baseline_loss": baseline_loss, "baseline_rsa": baseline_rsa,
            "perturbed_loss": val_loss, "perturbed_rsa": rsa_score,
            "delta_loss": val_loss - baseline_loss, "delta_rsa": rsa_score - baseline_rsa}


def measure_perturbation_effect(perturb_epoch, perturbation_type, baseline_checkpoint_dir, baseline_metrics_csv,
                                train_data, val_data, things_data, things_rdm, epsilon=0.1, batch_size=256,
                                lr=0.1, momentum=0.9, weight_decay=1e-4, warmup_epochs=5, total_epochs=100,
                                rank=0, world_size=1, local_rank=0, model_name="vit_base_patch16_224",
                                num_classes=1000, use_graph=True, dataset_order=True, noise_seed=None,
                                evaluator=None, log=print, data_path=None, num_workers=8):
    """MEAS:403-555.  Loads the baseline checkpoint of epoch N-1 (model, SGD momentum, scheduler), trains
    ONLY epoch N on the perturbed training data, evaluates on the clean validation data, computes the RSA
    of the CLS features on the THINGS images and returns the row of `RESULT_COLUMNS` (None when the baseline
    CSV has no row for the epoch or the checkpoint is missing, MEAS:427-430, 489-492).

    `train_data` / `val_data` / `things_data`: `ResidentImageSet`s on this rank's device (the reference
    builds ImageFolder loaders from `data_path`; the resident sets are the whole dataset, each rank reads its
    `DistributedSampler` share).  `noise_seed` seeds the device generator of the 'gaussian' perturbation
    (the reference's noise comes from unseeded worker processes).  With `train_data is None` and a
    `data_path`, the epoch streams a real ImageFolder tree through the reference's own torchvision pipeline and
    dataset wrappers instead (`imagenet_loaders`, MEAS:139-227; ImageNet does not fit in HBM)."""
    from . import vit
    say = log if (rank == 0 and log is not None) else (lambda *_: None)
    say(f"\n{'=' * 80}\nMeasuring: {perturbation_type} @ epoch {perturb_epoch}\n{'=' * 80}")
    base = baseline_row(baseline_metrics_csv, perturb_epoch)
    if base is None:
        say(f"No baseline data for epoch {perturb_epoch}")
        return None
    baseline_loss, baseline_rsa = base
    say(f"Baseline @ epoch {perturb_epoch}: loss={baseline_loss:.4f}, RSA={baseline_rsa:.4f}")
    checkpoint_path = os.path.join(baseline_checkpoint_dir, f"checkpoint_epoch_{perturb_epoch - 1:03d}.pth")
    if not os.path.exists(checkpoint_path):
        say(f"Checkpoint not found: {checkpoint_path}")
        return None

    device = train_data.images.device if train_data is not None else things_data.images.device
    model = vit.create_model(model_name, pretrained=False, num_classes=num_classes).to(device)
    trainer = vit.DataParallelTrainer(model, lr=lr, momentum=momentum, weight_decay=weight_decay,
                                      use_graph=use_graph)
    scheduler = vit.CosineAnnealingLRWithWarmup(trainer, warmup_epochs=warmup_epochs, max_epochs=total_epochs,
                                                eta_min=0)
    load_checkpoint(checkpoint_path, model, trainer, scheduler, device)

    if train_data is None:
        if data_path is None:
            raise ValueError("measure_perturbation_effect: pass resident train_data / val_data or a data_path")
        train_loader, val_loader, _ = imagenet_loaders(data_path, batch_size, num_workers, world_size, rank, device,
                                                       perturbation_type=perturbation_type, epsilon=epsilon,
                                                       shuffle_seed=42)
    else:
        gen = None
        if perturbation_type == "gaussian" and noise_seed is not None:
            gen = torch.Generator(device=device).manual_seed(noise_seed + rank)
        train_loader = ShardedLoader(train_data, batch_size, world_size, rank, shuffle=True,
                                     perturbation_type=perturbation_type, epsilon=epsilon, shuffle_seed=42,
                                     num_classes=num_classes, noise_generator=gen)
        val_loader = ShardedLoader(val_data, batch_size, world_size, rank, shuffle=False)
    things_loader = ShardedLoader(things_data, 8, world_size, rank, shuffle=False, with_names=True)  # MEAS:458-464
    train_loader.sampler.set_epoch(perturb_epoch)                                                     # MEAS:519

    say(f"Training perturbed epoch {perturb_epoch}...")
    train_one_epoch(trainer, train_loader, perturb_epoch, local_rank, world_size, log=say)
    scheduler.step()
    say("Evaluating...")
    val_loss, _ = validate(trainer, val_loader, local_rank, world_size)
    rsa_score, _ = compute_rsa_score(model, things_loader, things_rdm, local_rank, world_size,
                                     dataset_order=dataset_order, evaluator=evaluator)
    result = assemble_result(perturb_epoch, perturbation_type, baseline_loss, baseline_rsa, val_loss, rsa_score)
    say(f"Perturbed: loss={val_loss:.4f}, RSA={result['perturbed_rsa']:.4f}")
    say(f"Δ loss={result['delta_loss']:+.4f}, Δ RSA={result['delta_rsa']:+.4f}")
    return result


def measurement_conditions(perturb_epochs=DEFAULT_PERTURB_EPOCHS, perturbation_types=PERTURBATION_TYPES):
    """MEAS:629-635: epochs outer, perturbation types inner; epoch 0 has no prior checkpoint."""
    return [(e, t) for e in perturb_epochs if e != 0 for t in perturbation_types]


def measure_all(output_csv, perturb_epochs=DEFAULT_PERTURB_EPOCHS, perturbation_types=PERTURBATION_TYPES,
                rank=0, measure_fn=measure_perturbation_effect, **kwargs):
    """MEAS:625-662: every (epoch, type) measurement in the reference's order; rank 0 writes the result
    CSV (`RESULT_COLUMNS`).  Returns the list of result rows."""
    import pandas as pd
    results = []
    for perturb_epoch, perturbation_type in measurement_conditions(perturb_epochs, perturbation_types):
        r = measure_fn(perturb_epoch=perturb_epoch, perturbation_type=perturbation_type, rank=rank, **kwargs)
        if r is not None:
            results.append(r)
    if rank == 0:
        out_dir = os.path.dirname(os.path.abspath(output_csv))
        os.makedirs(out_dir, exist_ok=True)
        pd.DataFrame(results, columns=list(RESULT_COLUMNS) if results else None).to_csv(output_csv, index=False)
    return results


# ------------------------------------------------------------------------------------------------
# Script plumbing shared by vit_training/baseline/train_vit_sgd.py and
# vit_training/single_epoch/measure_single_epoch_perturbation_effect.py
# ------------------------------------------------------------------------------------------------
def setup_distributed():
    """VIT:13-27: (rank, world_size, local_rank) from the torchrun environment, NCCL process group."""
    import torch.distributed as dist
    if "RANK" in os.environ and "WORLD_SIZE" in os.environ:
        rank, world_size = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
        local_rank = int(os.environ["LOCAL_RANK"])
    else:
        print("Not using distributed mode")
        torch.cuda.set_device(0)
        return 0, 1, 0
    torch.cuda.set_device(local_rank)
    dist.init_process_group(backend="nccl", init_method="env://")
    return rank, world_size, local_rank


def parse_synthetic(spec):
    """'synthetic:<n_train>:<n_val>[:<num_classes>]' -> (n_train, n_val, num_classes) or None."""
    if not str(spec).startswith("synthetic"):
        return None
    parts = str(spec).split(":")[1:]
    nums = [int(p) for p in parts if p]
    n_train = nums[0] if len(nums) > 0 else 2048
    n_val = nums[1] if len(nums) > 1 else 512
    classes = nums[2] if len(nums) > 2 else 1000
    return n_train, n_val, classes


class StreamedLoader:
    """A host DataLoader (ImageFolder + the reference's transform pipeline, VIT:29-87 / MEAS:139-227) whose
    batches are moved to the device as they arrive; keeps `.sampler` for `set_epoch`."""

    def __init__(self, loader, device, with_names=False):
        self.loader, self.device, self.with_names = loader, device, with_names
        self.sampler = loader.sampler

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        for a, b in self.loader:
            if self.with_names:
                yield list(a), b.to(self.device, non_blocking=True)
            else:
                yield a.to(self.device, non_blocking=True), torch.as_tensor(b).to(self.device, non_blocking=True).long()


def imagenet_loaders(data_path, batch_size, num_workers, world_size, rank, device, perturbation_type=None,
                     epsilon=0.1, shuffle_seed=42):
    """VIT:29-87 with the perturbation hooks of MEAS:139-227 for a real ImageFolder tree (`train/`, `val/`):
    the same torchvision transforms, samplers and DataLoader settings; batches land on `device`."""
    from torch.utils.data import DataLoader
    from torchvision import datasets, transforms
    norm = transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])
    base_train = transforms.Compose([transforms.RandomResizedCrop(224), transforms.RandomHorizontalFlip(),
                                     transforms.ToTensor(), norm])
    if perturbation_type == "gaussian":
        train_transform = GaussianNoiseTransform(base_train, epsilon=epsilon)
    elif perturbation_type == "uniform_gray":
        train_transform = UniformGrayTransform(base_train)
    else:
        train_transform = base_train
    val_transform = transforms.Compose([transforms.Resize(256), transforms.CenterCrop(224), transforms.ToTensor(), norm])
    train_dataset = datasets.ImageFolder(os.path.join(data_path, "train"), transform=train_transform)
    if perturbation_type == "label_shuffle":
        train_dataset = ShuffledLabelsDataset(train_dataset, shuffle_seed=shuffle_seed)
    elif perturbation_type == "target_noise":
        train_dataset = TargetNoiseDataset(train_dataset, num_classes=1000, noise_seed=shuffle_seed)
    val_dataset = datasets.ImageFolder(os.path.join(data_path, "val"), transform=val_transform)
    train_sampler = DistributedSampler(train_dataset, num_replicas=world_size, rank=rank, shuffle=True)
    val_sampler = DistributedSampler(val_dataset, num_replicas=world_size, rank=rank, shuffle=False)
    kw = dict(prefetch_factor=2, persistent_workers=True) if num_workers > 0 else {}
    train_loader = DataLoader(train_dataset, batch_size=batch_size, sampler=train_sampler, num_workers=num_workers,
                              pin_memory=True, **kw)
    val_loader = DataLoader(val_dataset, batch_size=batch_size, sampler=val_sampler, num_workers=num_workers,
                            pin_memory=True)
    return StreamedLoader(train_loader, device), StreamedLoader(val_loader, device), train_sampler


def synthetic_things(device, n=48, dim=66, seed=2, img_size=224):
    """SURVEY 8d stand-ins: `n` random images and the human RDM `1 - corrcoef(randn(n, dim))`, diagonal 0."""
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(n, 3, img_size, img_size, generator=g)
    human = torch.randn(n, dim, generator=g).double().numpy()
    rdm = 1.0 - np.corrcoef(human)
    np.fill_diagonal(rdm, 0.0)
    return ResidentImageSet(images.to(device)), rdm


# ------------------------------------------------------------------------------------------------
# RSA of every checkpoint of a baseline run -> the `baseline_metrics_csv` the measurement reads
# ------------------------------------------------------------------------------------------------
RSA_RESULTS_COLUMNS = ("checkpoint", "epoch", "train_loss", "val_loss", "val_acc", "rsa_score")


def rsa_over_checkpoints(checkpoint_dir, things_data, things_rdm, output_csv=None, model_name="vit_base_patch16_224",
                         num_classes=1000, rank=0, world_size=1, dataset_order=True, evaluator=None, log=print):
    """The table shipped as Data/vit_results/rsa_results_final.csv (`checkpoint, epoch, train_loss, val_loss,
    val_acc, rsa_score`, one row per `checkpoint_epoch_XXX.pth` of a `train_vit_sgd.py` run): what
    `measure_perturbation_effect` reads as `baseline_metrics_csv` (MEAS:421-433 needs `epoch`, `val_loss`,
    `rsa_score`).  The reference ships the table but not the script that made it; the per-checkpoint
    computation is `compute_rsa_score` (MEAS:298-355) on the checkpoint's weights.

    Checkpoints are independent ("RSA at scale"): rank r evaluates the files r, r+W, ... on ALL
    RSA images, no data-path collective; the rows are gathered on rank 0, which writes the CSV.  Returns the
    rows (rank 0) or None."""
    import glob
    from . import vit
    files = sorted(glob.glob(os.path.join(checkpoint_dir, "checkpoint_epoch_*.pth")))
    if not files:
        raise FileNotFoundError(f"no checkpoint_epoch_*.pth under {checkpoint_dir}")
    device = things_data.images.device
    model = vit.create_model(model_name, pretrained=False, num_classes=num_classes).to(device)
    loader = ShardedLoader(things_data, 8, 1, 0, shuffle=False, with_names=True)
    if evaluator is None:
        from .rsa import RSAEvaluator
        evaluator = RSAEvaluator(load_reference_rdm(things_rdm), device)
    rows = []
    for path in files[rank::world_size]:
        ck = torch.load(path, map_location=device, weights_only=False)
        model.load_state_dict(ck["model_state_dict"])
        rho, _ = compute_rsa_score(model, loader, things_rdm, 0, 1, dataset_order=dataset_order, evaluator=evaluator)
        rows.append({"checkpoint": os.path.splitext(os.path.basename(path))[0], "epoch": int(ck["epoch"]),
                     "train_loss": float(ck.get("train_loss", float("nan"))),
                     "val_loss": float(ck.get("val_loss", float("nan"))),
                     "val_acc": float(ck.get("val_acc", float("nan"))), "rsa_score": float(rho)})
        if log is not None:
            log(f"[rank {rank}] {rows[-1]['checkpoint']}: RSA {rho:.4f}")
    if world_size > 1:
        d = _dist()
        if d is None:
            raise RuntimeError("world_size > 1 needs an initialised process group")
        gathered = [None] * world_size
        d.all_gather_object(gathered, rows)
        rows = [r for part in gathered for r in part]
    if rank != 0:
        return None
    rows.sort(key=lambda r: (r["epoch"], r["checkpoint"]))
    if output_csv:
        import pandas as pd
        os.makedirs(os.path.dirname(os.path.abspath(output_csv)), exist_ok=True)
        pd.DataFrame(rows, columns=list(RSA_RESULTS_COLUMNS)).to_csv(output_csv, index=False)
    return rows

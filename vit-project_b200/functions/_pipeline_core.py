"""Shared implementation behind the two drop-in pipeline modules

    functions/new_cvpr_train_behavior_things_pipeline.py        (reference NEW, 1226 lines)
    functions/cvpr_train_behavior_things_pipeline_baseline.py   (reference BASE, 823 lines)

Same public names, argument meaning, on-disk formats (CSV schema NEW:795 / BASE:636, DoRA checkpoint
keys NEW:665-683, random-state checkpoint NEW:709-727) and control flow semantics as the reference;
the arithmetic runs on libhba (sm_100a): the CLIP towers and DoRA merge through the plug-in ``clip``
module + ``hba.DoRALayer``, AdamW through ``hba.optim.FusedAdamW``, the RSA tail through ``hba.rsa``.

Host-synchronisation changes w.r.t. the reference (results unchanged): the per-step ``loss.item()``
calls (NEW:999/1003) and NaN checks (NEW:989-998) are evaluated on the device — the running loss is
a device scalar read once per epoch and a bad batch raises a device flag that makes the fused
optimiser skip the update, which is what the reference's ``continue`` achieves.
"""
from __future__ import annotations

import csv
import gc
import logging
import os
import random
import sys
from datetime import datetime

import numpy as np
import pandas as pd
import torch
import torch.nn as nn
from PIL import Image
from torch.utils.data import DataLoader, Dataset
from tqdm import tqdm

from hba import DoRALayer, ops, rsa
from hba.data import ResidentLoader, ResidentStore
from hba.engine import LossRequest, TrunkCache
from hba.optim import FusedAdamW
from src.models.CLIPs.clip_hba import clip

THINGS_MEAN = [0.52997664, 0.48070561, 0.41943838]
THINGS_STD = [0.27608301, 0.26593025, 0.28238822]

NEW_HEADERS = ["epoch", "train_loss", "test_loss", "behavioral_rsa_rho", "behavioral_rsa_p_value",
               "used_random_targets", "used_shuffled_targets", "used_uniform_images", "used_image_noise"]
BASE_HEADERS = NEW_HEADERS[:5]


# ------------------------------------------------------------------------------- seeding / logging
def seed_everything(seed):
    """NEW:35-48."""
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
    random.seed(seed)
    np.random.seed(seed)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False


def setup_logger(log_file_path):
    """File + stdout logger with the reference's line format (NEW:51-85)."""
    logger = logging.getLogger("training_logger")
    logger.setLevel(logging.INFO)
    logger.handlers = []
    fmt = logging.Formatter("%(asctime)s - %(levelname)s - %(message)s", datefmt="%Y-%m-%d %H:%M:%S")
    os.makedirs(os.path.dirname(log_file_path) or ".", exist_ok=True)
    for handler in (logging.FileHandler(log_file_path, mode="w"), logging.StreamHandler(sys.stdout)):
        handler.setLevel(logging.INFO)
        handler.setFormatter(fmt)
        logger.addHandler(handler)
    return logger


def _log_fn(logger):
    return logger.info if logger else print


# ------------------------------------------------------------------------------- data
class _ThingsTransform:
    """Resize((224, 224)) -> ToTensor() -> Normalize(THINGS_MEAN, THINGS_STD) of NEW:183-188 on a PIL image, restated
    without torchvision: `import torchvision` pulls in torch._dynamo (4 - 7 s per process, paid by every sweep
    worker), and for a PIL input these three transforms are: PIL's own bilinear resize, uint8 HWC -> float32 CHW
    / 255, (x - mean) / std.  Bit-identical to the torchvision Compose
    (tests/test_pipeline_cpu.py::test_things_transform_equals_torchvision)."""

    def __init__(self, size=(224, 224), mean=None, std=None):
        self.size = size
        self.mean = torch.tensor(THINGS_MEAN if mean is None else mean, dtype=torch.float32).view(-1, 1, 1)
        self.std = torch.tensor(THINGS_STD if std is None else std, dtype=torch.float32).view(-1, 1, 1)

    def __call__(self, img):
        img = img.resize(self.size[::-1], Image.BILINEAR)          # torchvision F.resize on a PIL image (size = (h, w))
        x = torch.from_numpy(np.array(img, np.uint8, copy=True))   # ToTensor: HWC uint8 ...
        if x.ndim == 2:
            x = x[:, :, None]
        x = x.permute(2, 0, 1).contiguous().to(torch.float32).div(255)   # ... -> CHW float32 in [0, 1]
        return x.sub_(self.mean).div_(self.std)                    # Normalize (on its own copy, like torchvision)


def _things_transform():
    return _ThingsTransform()


class ThingsDataset(Dataset):
    """(image_name, image[3,224,224], targets[66]) rows of the SPoSE csv (NEW:180-204)."""

    def __init__(self, csv_file, img_dir):
        self.csv_file = csv_file
        self.img_dir = img_dir
        self.transform = _things_transform()
        self.annotations = pd.read_csv(csv_file, index_col=0)

    def __len__(self):
        return len(self.annotations)

    def _image(self, index):
        name = self.annotations.iloc[index, 0]
        return name, self.transform(Image.open(os.path.join(self.img_dir, name)).convert("RGB"))

    def __getitem__(self, index):
        name, image = self._image(index)
        targets = torch.tensor(self.annotations.iloc[index, 1:].values.astype("float32"))
        return name, image, targets


class ThingsInferenceDataset(ThingsDataset):
    """(image_name, image) rows of the 48-image inference csv; carries the path of the human RDM
    (NEW:225-248)."""

    def __init__(self, inference_csv_file, img_dir, RDM48_triplet_dir):
        super().__init__(inference_csv_file, img_dir)
        self.RDM48_triplet_dir = RDM48_triplet_dir

    def __getitem__(self, index):
        return self._image(index)


class SubsetWithIndices(Dataset):
    """NEW:164-177."""

    def __init__(self, dataset, indices):
        self.dataset, self.indices = dataset, indices

    def __getitem__(self, idx):
        return self.dataset[self.indices[idx]]

    def __len__(self):
        return len(self.indices)


def load_dataset_split_indices(split_indices_path, logger=None):
    """NEW:137-161."""
    log = _log_fn(logger)
    if not os.path.exists(split_indices_path):
        log(f"Split indices file not found: {split_indices_path}")
        return None
    info = torch.load(split_indices_path)
    log(f"Loaded dataset split indices from: {split_indices_path}")
    log(f"  Train samples: {len(info['train_indices'])}")
    log(f"  Test samples: {len(info['test_indices'])}")
    log(f"  Random seed used: {info['random_seed']}")
    return info


def replace_with_gaussian_noise(image, mean, std):
    """NEW:207-221 (global torch RNG of the image's device)."""
    return torch.randn(tuple(image.size()), device=image.device) * std + mean


# ------------------------------------------------------------------------------- model
_FROZEN_MODELS = {}


def load_clip_to_cpu(backbone_name):
    """NEW:251-265 through the plug-in clip module.

    The frozen CLIP is kept per process (HBA_REUSE_MODEL=0 disables it): a sweep worker runs many
    conditions in one process (SWEEP:192-223) and every one of them starts from the same frozen
    checkpoint, so the staged weights and the frozen-trunk cache survive from condition to condition.
    On reuse the adapters of the previous condition are unwrapped back to the original out_proj."""
    reuse = os.environ.get("HBA_REUSE_MODEL", "1") != "0"
    path = clip._download(clip._MODELS[backbone_name], os.path.expanduser("~/.cache/clip"))
    key = (backbone_name, path, os.path.getmtime(path))
    if reuse and key in _FROZEN_MODELS:
        model = _FROZEN_MODELS[key]
        _warn_if_still_owned(model)
        for tower in (model.visual.transformer, model.transformer):
            for blk in tower.resblocks:
                if isinstance(blk.attn.out_proj, DoRALayer):
                    blk.attn.out_proj = blk.attn.out_proj.original_layer
        return model
    try:
        jit = torch.jit.load(path, map_location="cpu").eval()
        state_dict = jit.state_dict()
    except RuntimeError:
        state_dict = torch.load(path, map_location="cpu")
    model = clip.build_model(state_dict)
    if reuse:
        _FROZEN_MODELS[key] = model
    return model


_OWNERS = {}   # id(frozen clip model) -> weakref of the CLIPHBA built on it last


def _warn_if_still_owned(clip_model):
    """The shared frozen CLIP is about to be re-adapted for a new CLIPHBA: an earlier CLIPHBA on it that is still in
    use loses its adapters.  Sequential runs (a sweep worker) never hit this - the finished run's wrapper is garbage,
    at most held by a reference cycle - so this only speaks up for two models kept alive side by side."""
    ref = _OWNERS.get(id(clip_model))
    if ref is not None and ref() is not None:
        gc.collect()
        if ref() is not None:
            import warnings
            warnings.warn("a CLIPHBA built earlier in this process is still alive and shares its frozen CLIP with the "
                          "one being built: its adapters are removed now and it must not be used any more. Set "
                          "HBA_REUSE_MODEL=0 to give every CLIPHBA its own copy of the model.", RuntimeWarning,
                          stacklevel=4)


class CLIPHBA(nn.Module):
    """NEW:268-304: frozen CLIP scored against the fixed class prompts -> [B, n_prompts] fp32."""

    def __init__(self, classnames, backbone_name="RN50", pos_embedding=False):
        super().__init__()
        self.num_clip = len(classnames)
        self.clip_model = load_clip_to_cpu(backbone_name)
        import weakref
        _OWNERS[id(self.clip_model)] = weakref.ref(self)
        self.clip_model.float()
        self.pos_embedding = pos_embedding
        for p in self.clip_model.parameters():
            p.requires_grad = False
        self.tokenized_prompts = torch.stack([clip.tokenize(c) for c in classnames])
        self._cached_tokenized_prompts = None
        self._cached_device = None

    def forward(self, image):
        if self.clip_model.training:
            self.clip_model.eval()  # dropout never active on this path (NEW:288-289)
        if self._cached_tokenized_prompts is None or self._cached_device != image.device:
            self._cached_tokenized_prompts = self.tokenized_prompts.to(image.device)
            self._cached_device = image.device
        return self.clip_model(image, self._cached_tokenized_prompts, self.pos_embedding).float()


def _unwrap(model):
    return model.module if isinstance(model, nn.DataParallel) else model


def apply_dora_to_ViT(model, n_vision_layers=1, n_transformer_layers=1, r=8, dora_dropout=0.1, seed=123):
    """NEW:484-513: DoRA on attn.out_proj of the last vision blocks, then of the last text blocks
    (that order fixes how the global RNG is consumed)."""
    cm = _unwrap(model).clip_model
    for tower, n in ((cm.visual.transformer, n_vision_layers), (cm.transformer, n_transformer_layers)):
        for idx in range(-n, 0):
            blk = tower.resblocks[idx]
            blk.attn.out_proj = DoRALayer(blk.attn.out_proj, r=r, dora_dropout=dora_dropout)


def switch_dora_layers(model, freeze_all=True, dora_state=True):
    """NEW:516-544."""
    for p in model.parameters():
        p.requires_grad = not freeze_all
    if freeze_all:
        for mod in _unwrap(model).modules():
            if isinstance(mod, DoRALayer):
                for p in (mod.m, mod.delta_D_A, mod.delta_D_B):
                    p.requires_grad = dora_state
                if mod.bias is not None:
                    mod.bias.requires_grad = False


def count_trainable_parameters(model):
    return sum(p.numel() for p in model.parameters() if p.requires_grad)


def dora_module_paths(model, vision_layers, transformer_layers):
    cm = _unwrap(model).clip_model
    nv, nt = len(cm.visual.transformer.resblocks), len(cm.transformer.resblocks)
    return ([f"clip_model.visual.transformer.resblocks.{nv - vision_layers + i}.attn.out_proj"
             for i in range(vision_layers)] +
            [f"clip_model.transformer.resblocks.{nt - transformer_layers + i}.attn.out_proj"
             for i in range(transformer_layers)])


def _dora_state(model, paths):
    out = {}
    root = _unwrap(model)
    for path in paths:
        mod = root
        for attr in path.split("."):
            mod = getattr(mod, attr)
        for name in ("m", "delta_D_A", "delta_D_B"):
            out[f"{path}.{name}"] = getattr(mod, name).detach().cpu()
    return out


def find_dora_paths(model):
    return [n for n, m in _unwrap(model).named_modules() if isinstance(m, DoRALayer)]


# ------------------------------------------------------------------------------- checkpoints
class CheckpointWriter:
    """Per-epoch checkpoint files (NEW:657-728) written by a background thread: once the sweep
    epoch takes tens of milliseconds, the two `torch.save` calls per epoch are a measurable part of it.

    `submit(obj, path)` snapshots `obj` to host memory in the caller (device tensors are copied, host tensors
    cloned: later in-place updates of the live state do not reach the file), then the thread pickles and writes
    `path` through a temporary file + `os.replace`, so that a reader never sees a partial checkpoint.  Files and
    formats are exactly those of the synchronous path.  `flush()` waits for everything submitted and re-raises
    the first write error; it runs at the end of every `train_model`, before any checkpoint is loaded, and at
    interpreter exit.

    When it is used: inside the epoch loops of `train_model` (the `deferred()` scope, which flushes on the way out)
    unless `HBA_ASYNC_CKPT=0` - 12 grid conditions on one B200 ran at 536 instead of 503 conditions/hour with
    identical result CSVs (profiles/r02_grid12_async_ckpt.json).  A direct call of `save_dora_parameters` /
    `save_random_states` from user code stays synchronous, as in the reference (the file exists on return), unless
    `HBA_ASYNC_CKPT=1` forces the background writer everywhere."""

    def __init__(self):
        self._queue, self._thread, self._error = None, None, None
        self._scopes = 0

    def enabled(self):
        mode = os.environ.get("HBA_ASYNC_CKPT", "")
        if mode == "1":
            return True
        if mode == "0":
            return False
        return self._scopes > 0

    def deferred(self):
        """Scope of an epoch loop: checkpoints submitted inside may be written in the background; every one of
        them is on disk (or its error raised) when the scope is left."""
        import contextlib

        @contextlib.contextmanager
        def scope():
            self._scopes += 1
            try:
                yield self
            finally:
                self._scopes -= 1
                self.flush()
        return scope()

    @classmethod
    def snapshot(cls, obj):
        if torch.is_tensor(obj):
            t = obj.detach()
            return t.cpu() if t.is_cuda else t.clone()
        if isinstance(obj, dict):
            return {k: cls.snapshot(v) for k, v in obj.items()}
        if isinstance(obj, (list, tuple)):
            return type(obj)(cls.snapshot(v) for v in obj)
        if isinstance(obj, np.ndarray):
            return obj.copy()
        return obj

    def _run(self):
        while True:
            item = self._queue.get()
            try:
                if item is None:
                    return
                obj, path = item
                tmp = f"{path}.tmp{os.getpid()}"
                torch.save(obj, tmp)
                os.replace(tmp, path)
            except Exception as exc:  # noqa: BLE001  (reported by flush() in the training thread)
                if self._error is None:
                    self._error = exc
            finally:
                self._queue.task_done()

    def submit(self, obj, path):
        if not self.enabled():
            torch.save(obj, path)
            return
        import queue
        import threading
        if self._thread is None or not self._thread.is_alive():
            import atexit
            self._queue = queue.Queue()
            self._thread = threading.Thread(target=self._run, name="hba-checkpoint-writer", daemon=True)
            self._thread.start()
            atexit.register(self.flush)
        self._queue.put((self.snapshot(obj), path))

    def flush(self):
        if self._queue is not None:
            self._queue.join()
        if self._error is not None:
            err, self._error = self._error, None
            raise RuntimeError(f"background checkpoint write failed: {err}") from err


CHECKPOINTS = CheckpointWriter()


def save_dora_parameters(model, dora_parameters_path, epoch, logger=None):
    """NEW:657-693: one dict {<module path>.{m,delta_D_A,delta_D_B}: cpu tensor} per epoch.  The
    reference hard-codes blocks 22/23/11 of ViT-L/14; the adapters are located here, which yields
    the same keys for that model."""
    os.makedirs(dora_parameters_path, exist_ok=True)
    CHECKPOINTS.submit(_dora_state(model, find_dora_paths(model)),
                       os.path.join(dora_parameters_path, f"epoch{epoch + 1}_dora_params.pth"))


def save_random_states(optimizer, epoch, random_state_path, dataloader_generator, logger=None):
    """NEW:696-728."""
    ckpt = {"epoch": epoch, "optimizer_state_dict": optimizer.state_dict(),
            "torch_rng_state": torch.get_rng_state(), "numpy_rng_state": np.random.get_state(),
            "python_rng_state": random.getstate(),
            "dataloader_generator_state": dataloader_generator.get_state()}
    if torch.cuda.is_available():
        ckpt["cuda_rng_state"] = torch.cuda.get_rng_state()
        ckpt["cuda_rng_state_all"] = torch.cuda.get_rng_state_all()
    os.makedirs(random_state_path, exist_ok=True)
    path = os.path.join(random_state_path, f"epoch{epoch + 1}_random_states.pth")
    CHECKPOINTS.submit(ckpt, path)
    _log_fn(logger)(f"Random states saved: {path}")


def load_random_states(random_state_path, epoch, optimizer=None, dataloader_generator=None, logger=None):
    """NEW:88-134."""
    log = _log_fn(logger)
    CHECKPOINTS.flush()
    path = os.path.join(random_state_path, f"epoch{epoch}_random_states.pth")
    if not os.path.exists(path):
        log(f"Warning: Random state checkpoint not found: {path}")
        return False
    ckpt = torch.load(path, weights_only=False)
    torch.set_rng_state(ckpt["torch_rng_state"])
    np.random.set_state(ckpt["numpy_rng_state"])
    random.setstate(ckpt["python_rng_state"])
    if torch.cuda.is_available() and "cuda_rng_state" in ckpt:
        torch.cuda.set_rng_state(ckpt["cuda_rng_state"])
        if "cuda_rng_state_all" in ckpt:
            states = ckpt["cuda_rng_state_all"]
            if len(states) == torch.cuda.device_count():
                torch.cuda.set_rng_state_all(states)
    if optimizer is not None and "optimizer_state_dict" in ckpt:
        optimizer.load_state_dict(ckpt["optimizer_state_dict"])
        log(f"Restored optimizer state from epoch {epoch}")
    if dataloader_generator is not None and "dataloader_generator_state" in ckpt:
        dataloader_generator.set_state(ckpt["dataloader_generator_state"])
        log(f"Restored DataLoader generator state from epoch {epoch}")
    log(f"Random states loaded from: {path}")
    return True


# ------------------------------------------------------------------------------- resident data / cache
_STORES = {}


def resident_loaders(config, dataset, train_dataset, test_dataset, inference_dataset, device,
                     dataloader_generator, train_indices=None, test_indices=None):
    """HBM-resident equivalents of the three DataLoaders of NEW:1123-1126 (same batch order and
    generator consumption; images decoded once per process instead of once per epoch)."""
    # (keyed by the csv's mtime as well: a rewritten csv under the same path is decoded again)
    key = (config['csv_file'], os.path.getmtime(config['csv_file']), config['img_dir'], str(device))
    if key not in _STORES:
        _STORES[key] = ResidentStore(dataset, device)
    ikey = (config['inference_csv_file'], os.path.getmtime(config['inference_csv_file']), config['img_dir'], str(device))
    if ikey not in _STORES:
        _STORES[ikey] = ResidentStore(inference_dataset, device)
    bs = config['batch_size']
    train_loader = ResidentLoader(_STORES[key], bs, shuffle=True, generator=dataloader_generator,
                                  index_map=train_indices, dataset=train_dataset)
    test_loader = ResidentLoader(_STORES[key], bs, shuffle=False, index_map=test_indices,
                                 dataset=test_dataset)
    inference_loader = ResidentLoader(_STORES[ikey], bs, shuffle=False, dataset=inference_dataset)
    return train_loader, test_loader, inference_loader


def enable_trunk_cache(model, n_images):
    """Frozen-trunk activation cache (hba.engine.TrunkCache) sized for every distinct image.

    The engine (and with it the cache) belongs to the frozen CLIP, which a process reuses from run to run
    (`load_clip_to_cpu`): a later run on ANOTHER image set gets ids beyond the first run's capacity.  The cache is
    then replaced by a larger, empty one - and the step / forward graphs captured on this model, which replay
    gathers out of the old buffers, are dropped with it."""
    eng = _engine_of(model)
    if eng is None:
        return
    from hba.data import _NAME_IDS
    need = max(n_images, len(_NAME_IDS))
    if eng.trunk_cache is None or eng.trunk_cache.capacity < need:
        if eng.trunk_cache is not None:
            for holder in ("_hba_train_step", "_hba_forward_graphs"):
                _unwrap(model).__dict__.pop(holder, None)
        eng.trunk_cache = TrunkCache(need + 64)


def _engine_of(model):
    cm = getattr(_unwrap(model), "clip_model", None)
    return cm.hba_engine() if hasattr(cm, "hba_engine") else None


def _announce_ids(model, loader, allowed=True):
    """Tells the engine which images the next forward batch holds (ResidentLoader only), so that the
    frozen-trunk cache can be used; never for perturbed images."""
    eng = _engine_of(model)
    if eng is not None:
        eng.batch_ids = getattr(loader, "last_ids", None) if allowed else None


def fused_mse_ok(model, criterion, targets):
    """nn.MSELoss(reduction='mean') on [batch, prompts] fp32 targets (BDRV:31, NEW:994 / NEW:597) is computed
    by the head kernel itself (hba_cos_mse_fwd / _bwd); any other criterion runs as the caller wrote it."""
    return (type(criterion) is nn.MSELoss and criterion.reduction == "mean" and _engine_of(model) is not None
            and os.environ.get("HBA_FUSED_MSE", "1") != "0" and torch.is_tensor(targets) and targets.is_cuda
            and targets.dtype == torch.float32 and targets.ndim == 2
            and targets.shape[1] == _n_prompts(model))


def _n_prompts(model):
    m = _unwrap(model)
    n = getattr(m, "num_clip", None)
    if n is None and torch.is_tensor(getattr(m, "tokenized_prompts", None)):
        n = m.tokenized_prompts.shape[0]
    return -1 if n is None else int(n)


class _CachedForwardGraphs:
    """CUDA graphs of the no-grad forward on trunk-cached images (evaluation / RSA batches), one per
    batch size; see TrainStep.  __call__ returns a fresh tensor (a copy of the graph's static
    output)."""

    def __init__(self, model):
        self.model = model
        self.eng = _engine_of(model)
        self.entries = {}
        self.warm = set()

    @staticmethod
    def of(model):
        g = model.__dict__.get("_hba_forward_graphs")
        if g is None:
            g = model.__dict__["_hba_forward_graphs"] = _CachedForwardGraphs(model)
        return g

    def usable(self, loader):
        eng = self.eng
        return (os.environ.get("HBA_STEP_GRAPH", "1") != "0" and eng is not None and eng.trunk_cache is not None
                and getattr(loader, "last_ids_dev", None) is not None and not torch.is_grad_enabled()
                and eng.trunk_cache.x is not None and eng.trunk_cache.all_present(loader.last_ids, eng))

    def loss_total(self, device=None):
        """float64 device scalar that the fused-MSE evaluation batches accumulate loss * batch into.
        (`device`: where the batches live - the engine has no device before its first forward pass.)"""
        dev = torch.device(device) if device is not None else self.eng.device
        if dev is None:
            raise RuntimeError("loss_total: no device yet (pass the device of the evaluation batches)")
        if dev.type == "cuda" and dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        t = self.__dict__.get("_loss_total")
        if t is None or t.device != dev:
            t = self.__dict__["_loss_total"] = torch.zeros((), device=dev, dtype=torch.float64)
        return t

    def __call__(self, images, ids_dev, targets=None):
        """targets given: nn.MSELoss is fused into the head kernel and loss * batch is added to loss_total()."""
        if images.is_cuda and images.device.index != torch.cuda.current_device():
            with torch.cuda.device(images.device):
                return self._run(images, ids_dev, targets)
        return self._run(images, ids_dev, targets)

    def _run(self, images, ids_dev, targets):
        eng = self.eng
        key = (tuple(images.shape), eng._stamp, eng.precision, targets is not None)
        if key not in self.warm:
            self.warm.add(key)
            eng.batch_ids = ids_dev
            if targets is not None:
                eng.loss_request = LossRequest(targets.contiguous(), total=self.loss_total(images.device))
            return self.model(images)
        entry = self.entries.get(key)
        if entry is None:
            s_img, s_ids = torch.zeros_like(images), ids_dev.clone()
            s_tgt = targets.clone().contiguous() if targets is not None else None
            graph = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            c0 = ops.COUNTERS["launches"]
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                eng.batch_ids = s_ids
                if s_tgt is not None:
                    eng.loss_request = LossRequest(s_tgt, total=self.loss_total(images.device))
                out = self.model(s_img)
            entry = (graph, s_ids, s_tgt, out, ops.COUNTERS["launches"] - c0)
            ops.COUNTERS["launches"] = c0
            self.entries[key] = entry
        graph, s_ids, s_tgt, out, n_launch = entry
        s_ids.copy_(ids_dev, non_blocking=True)
        if s_tgt is not None:
            s_tgt.copy_(targets, non_blocking=True)
        graph.replay()
        ops.COUNTERS["launches"] += n_launch
        return out.clone() if targets is None else out


# ------------------------------------------------------------------------------- evaluation
def evaluate_model(model, data_loader, device, criterion):
    """NEW:584-602: sample-weighted mean loss; accumulated on the device, one read at the end."""
    model.eval()
    total = torch.zeros((), device=device, dtype=torch.float64)
    fwd = _CachedForwardGraphs.of(model)
    eng = fwd.eng
    with torch.no_grad():
        for _, images, targets in tqdm(data_loader, total=len(data_loader), desc="Evaluating",
                                       file=sys.stderr):
            images = images.to(device, non_blocking=True)
            targets = targets.to(device, non_blocking=True)
            if fused_mse_ok(model, criterion, targets):
                # MSE and the loss * batch accumulation run inside the head kernel (one launch)
                if fwd.loss_total(images.device) is not total:
                    total = fwd.loss_total(images.device)
                    total.zero_()
                if fwd.usable(data_loader):
                    fwd(images, data_loader.last_ids_dev, targets)
                else:
                    _announce_ids(model, data_loader)
                    eng.loss_request = LossRequest(targets.contiguous(), total=total)
                    model(images)
                continue
            if fwd.usable(data_loader):
                predictions = fwd(images, data_loader.last_ids_dev)
            else:
                _announce_ids(model, data_loader)
                predictions = model(images)
            total += criterion(predictions, targets).double() * images.size(0)
    return float(total) / len(data_loader.dataset)


_RSA_CACHE = {}


def _reference_rdm(path):
    import scipy.io
    return scipy.io.loadmat(path)["RDM48_triplet"]


def behavioral_RSA(model, inference_loader, device, logger=None):
    """NEW:605-654 -> (rho, p_value, model_rdm).  The embeddings stay in HBM; RDM, average-tie
    ranking and Pearson-on-ranks run in hba.rsa."""
    model.eval()
    log = _log_fn(logger)
    names, chunks = [], []
    with torch.no_grad():
        for image_name, image in inference_loader:
            image = image.to(device, non_blocking=True)
            fwd = _CachedForwardGraphs.of(model)
            if fwd.usable(inference_loader):
                chunks.append(fwd(image, inference_loader.last_ids_dev))
            else:
                _announce_ids(model, inference_loader)
                chunks.append(model(image))
            names.extend(image_name)
    emb = torch.cat(chunks, 0)
    log(f"First 10 image names: {names[:5]}")
    log(f"Embedding matrix shape: {tuple(emb.shape)}\n")
    path = inference_loader.dataset.RDM48_triplet_dir
    key = (path, os.path.getmtime(path), str(emb.device))   # (a rewritten .mat is read again)
    if key not in _RSA_CACHE:
        _RSA_CACHE[key] = rsa.RSAEvaluator(_reference_rdm(path), emb.device)
    return _RSA_CACHE[key](emb)


# ------------------------------------------------------------------------------- perturbations
def shuffle_targets(targets, perturb_seed=None, generator=None):
    """NEW:731-779: permute the batch dimension of `targets` (torch.randperm on its device)."""
    saved = None
    if generator is None and perturb_seed is not None:
        saved = (torch.get_rng_state(), np.random.get_state(), random.getstate())
        torch.manual_seed(perturb_seed)
        np.random.seed(perturb_seed)
        random.seed(perturb_seed)
    n = targets.shape[0]
    perm = (torch.randperm(n, device=targets.device, generator=generator) if generator is not None
            else torch.randperm(n, device=targets.device))
    shuffled = targets.clone()[perm]
    if saved is not None:
        torch.set_rng_state(saved[0])
        np.random.set_state(saved[1])
        random.setstate(saved[2])
    return shuffled


IMAGE_PERTURBATIONS = ("image_noise", "uniform_images")
PERTURB_FLAGS = {"random_target": "used_random_targets", "label_shuffle": "used_shuffled_targets",
                 "uniform_images": "used_uniform_images", "image_noise": "used_image_noise"}


class Perturbation:
    """The four perturbation types of NEW:843-982 with their exact RNG discipline: per-batch seed
    perturb_seed + training_run*1000 + batch_idx (independent of the epoch)."""

    def __init__(self, kind, training_run, length, seed, distribution, mean, std):
        self.kind, self.training_run, self.length = kind, training_run, length
        self.seed, self.distribution, self.mean, self.std = seed, distribution, mean, std
        self.first = training_run - 1           # 0-indexed first perturbed epoch (NEW:844)
        self.last = self.first + length - 1

    def active(self, epoch):
        return self.kind in PERTURB_FLAGS and self.first <= epoch <= self.last

    def in_window(self, epoch):
        return self.first <= epoch <= self.last

    def flags(self, epoch):
        f = dict.fromkeys(PERTURB_FLAGS.values(), False)
        if self.active(epoch):
            f[PERTURB_FLAGS[self.kind]] = True
        return f

    def apply(self, images, targets, batch_idx, device):
        s = self.seed + self.training_run * 1000 + batch_idx
        if self.kind == "image_noise":
            torch.manual_seed(s)
            if torch.cuda.is_available():
                torch.cuda.manual_seed_all(s)
            for i in range(len(images)):
                images[i] = replace_with_gaussian_noise(images[i], self.mean, self.std)
        elif self.kind == "uniform_images":
            images = torch.ones_like(images) * 0.5
        elif self.kind == "random_target":
            gen = torch.Generator(device=device)
            gen.manual_seed(s)
            noise = torch.randn(targets.shape, device=device, dtype=torch.float32, generator=gen)
            targets = noise * self.std + self.mean if self.distribution == "target" else noise
        elif self.kind == "label_shuffle":
            gen = torch.Generator(device=device)
            gen.manual_seed(s)
            targets = shuffle_targets(targets, generator=gen)
        return images, targets


# ------------------------------------------------------------------------------- training epoch
class _BadBatchFlag:
    """Device-side NaN/Inf guard (reference: NEW:932-935, 989-998)."""

    def __init__(self, device):
        self.step = torch.zeros(1, dtype=torch.int32, device=device)
        self.total = torch.zeros(1, dtype=torch.int32, device=device)

    def check(self, *tensors):
        self.step.zero_()
        for t in tensors:
            ops.nonfinite_flag(t.detach().reshape(-1).float().contiguous(), self.step)
        self.total += self.step
        return self.step


class TrainStep:
    """One optimisation step of NEW:985-1003 (zero_grad -> forward -> MSE -> NaN guard -> backward ->
    AdamW -> loss accumulation), launched either kernel by kernel or as a replayed CUDA graph.

    Three ways to run the SAME kernels with the same arguments in the same order:
      * eager          - non-fused optimiser, HBA_STEP_GRAPH=0, the first step of a shape (it allocates
                         every workspace / gradient / optimiser-state tensor the graphs reuse), and steps
                         that FILL the frozen-trunk cache (host-side bookkeeping of the filled ids);
      * "cached" graph - every image of the batch is in the trunk cache: ~90 launches / 1.5 ms of GPU
                         time behind ~4 ms of Python; static inputs are the image ids and the targets;
      * "full" graph   - no cache for this batch (image perturbations, cache disabled): the ~335
                         launches of the whole trunk + text tower step; static inputs are the images
                         and the targets.
    Target perturbations (random targets / label shuffle, Philox streams of NEW:918-959) are applied by
    the caller before the step.  Graph replay is bit-identical to the eager step
    (tests/test_gpu_pipeline.py::test_captured_step_graphs_are_bit_identical_to_eager_steps).
    `total` accumulates loss * batch over un-skipped steps, `last_loss` holds the last step's loss,
    `guard.total` counts skipped (NaN/Inf) batches - all on the device, read once per epoch."""

    def __init__(self, model, optimizer, criterion, device):
        self.model, self.opt, self.crit, self.device = model, optimizer, criterion, device
        self.total = torch.zeros((), device=device, dtype=torch.float64)
        self.last_loss = torch.zeros((), device=device, dtype=torch.float32)
        self.guard = _BadBatchFlag(device)
        self.eng = _engine_of(model)
        self.fused = isinstance(optimizer, FusedAdamW)
        self.entries = {}
        self.warm = set()

    @staticmethod
    def of(model, optimizer, criterion, device):
        st = model.__dict__.get("_hba_train_step")
        if st is None or st.opt is not optimizer or st.crit is not criterion:
            st = model.__dict__["_hba_train_step"] = TrainStep(model, optimizer, criterion, device)
        return st

    def start_epoch(self):
        if self.eng is not None and self.eng.device is not None:
            self.eng.ensure(self.eng.device)   # frozen weights changed since the graphs were captured?
        self.total.zero_()
        self.guard.total.zero_()

    def _graphs_on(self):
        return os.environ.get("HBA_STEP_GRAPH", "1") != "0" and self.eng is not None and self.fused

    def _body(self, images, ids, targets):
        self.opt.zero_grad()
        if self.eng is not None:
            self.eng.batch_ids = ids
        if self.fused and fused_mse_ok(self.model, self.crit, targets) and targets.shape[0] == images.shape[0]:
            # nn.MSELoss, the NaN / Inf guard and the loss bookkeeping are part of the head kernel; the head
            # backward reads (pred, target) directly - no eager loss kernels in the step
            eng = self.eng
            eng.loss_request = LossRequest(targets.contiguous(), self.guard.step, self.guard.total, self.total)
            self.model(images)
            # (take the loss off the engine: a tensor that outlives the step keeps its autograd graph - and the
            # AccumulateGrad nodes bound to this step's stream - alive into the next graph capture)
            loss, eng.loss_out = eng.loss_out, None
            if loss is None:
                raise RuntimeError("libhba: the forward pass did not consume the fused-loss request")
            loss.backward()
            self.opt.step(skip_flag=self.guard.step)
            self.last_loss.copy_(loss.detach())
            return
        predictions = self.model(images)
        loss = self.crit(predictions, targets)
        bad = self.guard.check(predictions, loss.reshape(1), targets)
        loss.backward()
        if self.fused:
            self.opt.step(skip_flag=bad)
        elif int(bad) == 0:
            self.opt.step()
        self.last_loss.copy_(loss.detach())
        self.total += torch.where(bad[0] == 0, loss.detach().double(), torch.zeros_like(self.total)) * images.size(0)

    def __call__(self, images, targets, ids_host=None, ids_dev=None):
        """ids_host / ids_dev: trunk-cache ids of the batch (None for perturbed images / no cache)."""
        if images.is_cuda and images.device.index != torch.cuda.current_device():
            with torch.cuda.device(images.device):   # graph capture / replay happen on the current device
                return self._run(images, targets, ids_host, ids_dev)
        return self._run(images, targets, ids_host, ids_dev)

    def _run(self, images, targets, ids_host, ids_dev):
        eng = self.eng
        cache = eng.trunk_cache if eng is not None else None
        have_ids = cache is not None and ids_host is not None
        cached = have_ids and ids_dev is not None and cache.x is not None and cache.all_present(ids_host, eng)
        if not self._graphs_on() or (have_ids and not cached):
            return self._body(images, ids_host if have_ids else None, targets)
        mode = "cached" if cached else "full"
        key = (mode, tuple(images.shape), eng._stamp, eng.precision, eng.cache_text)
        if key not in self.warm:
            self.warm.add(key)
            return self._body(images, ids_dev if cached else None, targets)
        entry = self.entries.get(key)
        if entry is None:
            # (the cached path never reads the images: a zero tensor carries the shape)
            s_img = torch.zeros_like(images) if cached else images.clone()
            s_ids = ids_dev.clone() if cached else None
            s_tgt = targets.clone()
            graph = torch.cuda.CUDAGraph()
            gc.collect()   # no autograd graph of an earlier (eager, default-stream) step may survive into the capture
            torch.cuda.synchronize()
            c0 = ops.COUNTERS["launches"]
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                self._body(s_img, s_ids, s_tgt)
            entry = (graph, s_img, s_ids, s_tgt, ops.COUNTERS["launches"] - c0)
            ops.COUNTERS["launches"] = c0
            self.entries[key] = entry
        graph, s_img, s_ids, s_tgt, n_launch = entry
        if cached:
            s_ids.copy_(ids_dev, non_blocking=True)
        else:
            s_img.copy_(images, non_blocking=True)
        s_tgt.copy_(targets, non_blocking=True)
        graph.replay()
        ops.COUNTERS["launches"] += n_launch


def train_one_epoch(model, train_loader, device, optimizer, criterion, epoch, epochs, perturb=None,
                    log=print):
    """One pass over train_loader (NEW:873-1004 / BASE:644-659): returns the sample-weighted mean
    loss, read from the device once."""
    step = TrainStep.of(model, optimizer, criterion, device)
    step.start_epoch()
    active = perturb is not None and perturb.active(epoch)
    perturbed_images = active and perturb.kind in IMAGE_PERTURBATIONS
    bar = tqdm(enumerate(train_loader), total=len(train_loader), desc=f"Epoch {epoch + 1}/{epochs}",
               file=sys.stderr)
    for batch_idx, (_, images, targets) in bar:
        images = images.to(device, non_blocking=True)
        targets = targets.to(device, non_blocking=True)
        if active:
            images, targets = perturb.apply(images, targets, batch_idx, device)
        if perturbed_images:   # never cache (or serve from the cache) a perturbed image
            step(images, targets)
        else:
            step(images, targets, getattr(train_loader, "last_ids", None),
                 getattr(train_loader, "last_ids_dev", None))
    n_bad = int(step.guard.total)
    if n_bad:
        log(f"ERROR: NaN/Inf detected in {n_bad} batch(es) of epoch {epoch + 1}; they were skipped")
    return float(step.total) / len(train_loader.dataset)


def select_device(cuda_flag):
    """config['cuda'] (NEW:1137-1144): -1 / 0 / 1 -> CUDA device; anything else asked for the CPU in
    the reference, which this build does not have."""
    if cuda_flag == -1:
        return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cuda")
    if cuda_flag in (0, 1):
        if torch.cuda.is_available():
            # CUDA-graph capture, side streams and the C-ABI launches all work on the CURRENT device: make the
            # selected GPU current for the whole run (a sweep worker pins its GPU through CUDA_VISIBLE_DEVICES
            # and maps every request onto the one device it sees)
            if cuda_flag >= torch.cuda.device_count():
                raise RuntimeError(f"config['cuda'] = {cuda_flag} but only {torch.cuda.device_count()} CUDA "
                                   "device(s) are visible")
            torch.cuda.set_device(cuda_flag)
        return torch.device(f"cuda:{cuda_flag}")
    raise RuntimeError("config['cuda'] selects the CPU, but the libhba pipeline has no CPU path; "
                       "use 0, 1 or -1")


def make_optimizer(model, lr):
    """AdamW(model.parameters(), lr) (NEW:1181) with the fused multi-tensor step."""
    return FusedAdamW(model.parameters(), lr=lr)


def open_logger(config):
    log_dir = os.path.dirname(config["checkpoint_path"])
    stamp = datetime.now().strftime("%Y%m%d_%H%M%S")
    log_file = os.path.join(log_dir, f"training_log_{stamp}.txt")
    logger = setup_logger(log_file)
    logger.info("=" * 80)
    logger.info("Starting Training Run")
    logger.info(f"Log file: {log_file}")
    logger.info("=" * 80)
    return logger


def rng_state_will_be_restored(config):
    """True when `run_behavioral_training` (NEW:1154-1199) is going to overwrite everything the global RNG decides
    before training starts: the DoRA parameters come from a checkpoint AND the random states of that epoch are
    loaded.  Only then may the draws of the model construction be skipped (see `build_model`)."""
    resume = config.get("resume_from_epoch", 0)
    if resume <= 0 or "training_run" not in config:
        return False
    if config.get("resume_dora_parameters_path"):
        dora = os.path.join(config["resume_dora_parameters_path"], f"epoch{resume}_dora_params.pth")
    else:
        dora = os.path.join(config.get("baseline_dora_directory") or "", f"epoch{config['training_run'] - 1}_dora_params.pth")
    prior = config.get("resume_random_state_path") or config.get("baseline_random_state_path")
    return bool(prior) and os.path.exists(dora) and os.path.exists(os.path.join(prior, f"epoch{resume}_random_states.pth"))


def build_model(config, device, logger):
    """CLIPHBA + DoRA surgery of NEW:1128-1152 / BASE:760-778.  The reference's `CLIPHBA(...)` constructs the
    published CLIP with its random initialisation before loading the checkpoint, which moves the global RNG; the
    plug-in `clip.build_model` draws nothing, so for a run whose RNG state is not restored from a checkpoint
    afterwards the same draws are replayed here (`clip.replay_constructor_draws`) - the DoRA matrices drawn next
    are then the reference's.  HBA_CONSTRUCTOR_RNG=1 / 0 forces / forbids the replay."""
    from functions.spose_dimensions import classnames66
    pos_embedding = config["backbone"] != "RN50"
    logger.info(f"pos_embedding is {pos_embedding}")
    model = CLIPHBA(classnames=classnames66, backbone_name=config["backbone"], pos_embedding=pos_embedding)
    mode = os.environ.get("HBA_CONSTRUCTOR_RNG", "")
    if mode == "1" or (mode != "0" and not rng_state_will_be_restored(config)):
        if hasattr(clip, "replay_constructor_draws"):
            clip.replay_constructor_draws(model.clip_model)
    apply_dora_to_ViT(model, n_vision_layers=config["vision_layers"],
                      n_transformer_layers=config["transformer_layers"], r=config["rank"],
                      dora_dropout=0.1)
    switch_dora_layers(model, freeze_all=True, dora_state=True)
    return model


def describe_run(model, config, logger):
    logger.info("\nModel Configuration:")
    logger.info("-------------------")
    for k, v in config.items():
        logger.info(f"{k}: {v}")
    logger.info("\nUpdating layers:")
    for name, p in model.named_parameters():
        if p.requires_grad:
            logger.info(name)
    logger.info(f"\nNumber of trainable parameters: {count_trainable_parameters(model)}\n")


def append_csv_row(path, row):
    with open(path, "a", newline="") as f:
        csv.writer(f).writerow(row)

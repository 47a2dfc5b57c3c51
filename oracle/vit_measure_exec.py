"""ORACLE (test infrastructure only - never imported by the product path).

Runs the ViT baseline epochs + single-epoch perturbation measurements on a tiny JPEG ImageFolder tree, CPU only,
through one of two arms, and prints the results as one JSON line:

  --arm reference   the REFERENCE'S OWN functions, executed here: `get_dataloaders`, `train_one_epoch`, `validate`,
                    `save_checkpoint` of Training/vit_training/baseline/train_vit_sgd.py and
                    `measure_perturbation_effect` of Training/vit_training/single_epoch/
                    measure_single_epoch_perturbation_effect.py, unmodified, with only their environment stubbed:
                    `timm.create_model` -> the restated ViT (oracle/vit_ref.py, tiny config), `.cuda()` -> identity,
                    `torch.load(map_location='cuda:0')` -> CPU, a world-size-1 `gloo` process group for their
                    unconditional all-reduces.  (CUDA autocast / GradScaler disable themselves without a GPU: fp32.)
                    Needs /root/reference.
  --arm product     the product's host side (hba.vit_train: `imagenet_loaders`, `train_one_epoch`, `validate`,
                    `save_checkpoint`, `measure_perturbation_effect(data_path=...)`) with the device trainer replaced
                    by a CPU stand-in of the same interface (torch autograd + torch.optim.SGD) and the libhba RSA
                    tail by NumPy / SciPy.

Each arm also runs the collective tails on a world-size-2 `gloo` group (two spawned ranks; lr 0 so that the
reference's un-wrapped model copies stay identical without DDP): the reference's `train_one_epoch`, `validate`
(loss = SUM over ranks of the rank means, VIT:196) and `compute_rsa_score` (rank-strided embeddings concatenated
rank-major, MEAS:326-334) against `hba.vit_train`'s with `reference_rank_sum=True` / `dataset_order=False`.

Both arms see the same files and the same RNG streams (the model factory re-seeds torch after building the
model; neither side draws from the global generator between that point and the first DataLoader iterator, so
the worker seeds - hence the random crops / flips / Gaussian images - coincide), so every number must be
IDENTICAL.  tests/golden/vit_measure_exec.json holds the reference arm's output (written by `--write-golden`);
tests/test_vit_measure_cpu.py compares the product arm with it, and with a live reference arm where
/root/reference is mounted.
"""
import argparse
import json
import os
import sys
import tempfile
import types
import warnings

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vit-project_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

KINDS = ("label_shuffle", "target_noise", "uniform_gray", "gaussian")
BATCH, WORKERS, EPOCHS, LR = 4, 1, 2, 0.01


def build_fixture(root):
    """Tiny ImageNet-like tree (3 classes x 4 train / 2 val JPEGs), 12 'THINGS' JPEGs + CSV + RDM .mat."""
    import pandas as pd
    import scipy.io
    from PIL import Image
    rng = np.random.RandomState(0)
    for split, n in (("train", 4), ("val", 2)):
        for c in range(3):
            d = os.path.join(root, "data", split, f"class{c}")
            os.makedirs(d)
            for i in range(n):
                Image.fromarray(rng.randint(0, 255, (40, 48, 3), dtype=np.uint8)).save(os.path.join(d, f"{i}.jpg"))
    os.makedirs(os.path.join(root, "things"))
    names = []
    for i in range(12):
        names.append(f"t{i:02d}.jpg")
        Image.fromarray(rng.randint(0, 255, (300, 280, 3), dtype=np.uint8)).save(os.path.join(root, "things", names[-1]))
    pd.DataFrame({"image_name": names}).to_csv(os.path.join(root, "things.csv"), index=False)
    rdm = 1 - np.corrcoef(rng.randn(12, 9))
    np.fill_diagonal(rdm, 0)
    scipy.io.savemat(os.path.join(root, "rdm.mat"), {"RDM48_triplet": rdm})
    return {"data": os.path.join(root, "data"), "things_csv": os.path.join(root, "things.csv"),
            "things_dir": os.path.join(root, "things"), "rdm": os.path.join(root, "rdm.mat"), "ck": os.path.join(root, "ck")}


def factory(*_a, **_k):
    from oracle import vit_ref
    m = vit_ref.VisionTransformerRef(img_size=224, patch_size=16, embed_dim=64, depth=1, num_heads=1, num_classes=1000)
    torch.manual_seed(1234)      # both arms continue from the same RNG state after building a model
    return m


def add_rsa_column(ck):
    import pandas as pd
    m = pd.read_csv(os.path.join(ck, "training_metrics.csv"))
    m["rsa_score"] = [0.1 * (i + 1) for i in range(len(m))]
    path = os.path.join(ck, "with_rsa.csv")
    m.to_csv(path, index=False)
    return path


def plain(v):
    return None if v is None else (v if isinstance(v, (str, int)) else float(v))


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(fx):
    import torch.distributed as dist
    from oracle import make_vit_measure_golden as mk
    MEAS, VIT = mk.load_reference_scripts()
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    real_load = torch.load
    torch.load = lambda f, map_location=None, **k: real_load(f, map_location="cpu", weights_only=False)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{29300 + os.getpid() % 500}", rank=0, world_size=1)
    MEAS.timm.create_model = factory
    VIT.timm.create_model = factory
    torch.manual_seed(0)
    model = factory()
    opt = torch.optim.SGD(model.parameters(), lr=LR, momentum=0.9, weight_decay=1e-4)
    sched = VIT.CosineAnnealingLRWithWarmup(opt, 5, 100)
    scaler = VIT.GradScaler()
    tl, vl, sampler = VIT.get_dataloaders(fx["data"], BATCH, WORKERS, 1, 0)
    os.makedirs(fx["ck"])
    baseline = []
    for epoch in range(EPOCHS):
        sampler.set_epoch(epoch)
        a = VIT.train_one_epoch(model, tl, opt, scaler, epoch, 0, 1)
        sched.step()
        b, c = VIT.validate(model, vl, 0, 1)
        VIT.save_checkpoint(epoch, types.SimpleNamespace(module=model), opt, sched, scaler, a, b, c, fx["ck"], 0)
        baseline.append([plain(a), plain(b), plain(c)])
    csv = add_rsa_column(fx["ck"])
    rows = {}
    for kind in KINDS:
        r = MEAS.measure_perturbation_effect(1, kind, fx["ck"], csv, fx["data"], fx["things_csv"], fx["things_dir"],
                                             fx["rdm"], 0.1, BATCH, 0.1, 0.9, 1e-4, 5, 100, WORKERS, 0, 1, 0)
        rows[kind] = {k: plain(v) for k, v in r.items()}
    missing = MEAS.measure_perturbation_effect(7, "gaussian", fx["ck"], csv, fx["data"], fx["things_csv"], fx["things_dir"],
                                               fx["rdm"], 0.1, BATCH, 0.1, 0.9, 1e-4, 5, 100, WORKERS, 0, 1, 0)
    dist.destroy_process_group()
    return {"baseline": baseline, "measure": rows, "missing_epoch": missing,
            "metrics_csv": open(os.path.join(fx["ck"], "training_metrics.csv")).read()}


# ------------------------------------------------------------------------------------------------ product arm
class TorchTrainer:
    """CPU stand-in with hba.vit.DataParallelTrainer's interface."""

    def __init__(self, model, lr=0.1, momentum=0.9, weight_decay=1e-4, process_group=None, use_graph=False):
        self.model = model
        self.opt = torch.optim.SGD(model.parameters(), lr=lr, momentum=momentum, weight_decay=weight_decay)
        self.param_groups = self.opt.param_groups

    def step(self, images, labels):
        self.model.train()
        self.opt.zero_grad()
        out = self.model(images)
        loss = F.cross_entropy(out, labels)
        loss.backward()
        self.opt.step()
        return loss.detach().reshape(1), out.max(1)[1].eq(labels).sum().to(torch.int32).reshape(1)

    def evaluate(self, images, labels):
        self.model.eval()
        with torch.no_grad():
            out = self.model(images)
            return F.cross_entropy(out, labels).reshape(1), out.max(1)[1].eq(labels).sum().to(torch.int32).reshape(1)

    def state_dict(self):
        return self.opt.state_dict()

    def load_state_dict(self, sd):
        self.opt.load_state_dict(sd)


class ScipyEvaluator:
    def __init__(self, rdm):
        self.rdm, self.N = np.asarray(rdm), len(rdm)

    def __call__(self, emb, want_rdm=True):
        from oracle import vit_measure_ref as ref
        rho, p = ref.rsa_tail_ref(emb.detach().numpy(), self.rdm)
        return rho, p, None


def run_product(fx):
    import importlib.util
    from hba import vit, vit_train as vt
    vit.create_model = factory
    vit.DataParallelTrainer = TorchTrainer
    spec = importlib.util.spec_from_file_location("_measure_script", os.path.join(
        ROOT, "vit-project_b200", "vit_training", "single_epoch", "measure_single_epoch_perturbation_effect.py"))
    script = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(script)
    cpu = torch.device("cpu")
    torch.manual_seed(0)
    model = factory()
    tr = TorchTrainer(model, lr=LR)
    sched = vit.CosineAnnealingLRWithWarmup(tr, 5, 100)
    tl, vl, sampler = vt.imagenet_loaders(fx["data"], BATCH, WORKERS, 1, 0, cpu)
    baseline = []
    for epoch in range(EPOCHS):
        sampler.set_epoch(epoch)
        a = vt.train_one_epoch(tr, tl, epoch, log=None)
        sched.step()
        b, c = vt.validate(tr, vl)
        vt.save_checkpoint(epoch, model, tr, sched, a, b, c, fx["ck"])
        baseline.append([plain(a), plain(b), plain(c)])
    csv = add_rsa_column(fx["ck"])
    things, rdm = script.load_things(fx["things_csv"], fx["things_dir"], fx["rdm"], cpu)
    kw = dict(baseline_checkpoint_dir=fx["ck"], baseline_metrics_csv=csv, train_data=None, val_data=None, things_data=things,
              things_rdm=rdm, epsilon=0.1, batch_size=BATCH, evaluator=ScipyEvaluator(rdm), log=None, data_path=fx["data"],
              num_workers=WORKERS)
    rows = {kind: {k: plain(v) for k, v in vt.measure_perturbation_effect(1, kind, **kw).items()} for kind in KINDS}
    return {"baseline": baseline, "measure": rows, "missing_epoch": vt.measure_perturbation_effect(7, "gaussian", **kw),
            "metrics_csv": open(os.path.join(fx["ck"], "training_metrics.csv")).read()}


# ------------------------------------------------------------------------------------------------ world size 2
def _collective_worker(rank, world, port, arm, fx, out_dir):
    """The collective tails on a world-size-2 `gloo` group: per-epoch training loss
    (mean over ranks), validation loss (SUM over ranks of the rank means - the reference never divides) and the
    RSA score of rank-strided, rank-major concatenated embeddings."""
    import torch.distributed as dist
    warnings.filterwarnings("ignore")
    torch.set_num_threads(1)
    sys.stdout = sys.stderr
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = factory()
    res = {}
    if arm == "reference":
        from oracle import make_vit_measure_golden as mk
        from torch.utils.data import DataLoader, DistributedSampler
        from torchvision import transforms
        MEAS, VIT = mk.load_reference_scripts()
        torch.Tensor.cuda = lambda self, *a, **k: self
        opt = torch.optim.SGD(model.parameters(), lr=0.0, momentum=0.9, weight_decay=0.0)   # lr 0: ranks stay identical without DDP
        tl, vl, sampler = VIT.get_dataloaders(fx["data"], BATCH, WORKERS, world, rank)
        sampler.set_epoch(0)
        res["train_loss"] = VIT.train_one_epoch(model, tl, opt, VIT.GradScaler(), 0, rank, world)
        res["val"] = list(VIT.validate(model, vl, rank, world))
        tf = transforms.Compose([transforms.Resize(256), transforms.CenterCrop(224), transforms.ToTensor(),
                                 transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
        ds = MEAS.THINGSInferenceDataset(fx["things_csv"], fx["things_dir"], fx["rdm"], tf)
        loader = DataLoader(ds, batch_size=8, sampler=DistributedSampler(ds, num_replicas=world, rank=rank, shuffle=False))
        res["rsa"] = list(MEAS.compute_rsa_score(model, loader, fx["rdm"], rank, world))      # MEAS:451-464, 298-355
    else:
        import importlib.util
        from hba import vit_train as vt
        spec = importlib.util.spec_from_file_location("_measure_script", os.path.join(
            ROOT, "vit-project_b200", "vit_training", "single_epoch", "measure_single_epoch_perturbation_effect.py"))
        script = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(script)
        cpu = torch.device("cpu")
        tr = TorchTrainer(model, lr=0.0, weight_decay=0.0)
        tl, vl, sampler = vt.imagenet_loaders(fx["data"], BATCH, WORKERS, world, rank, cpu)
        sampler.set_epoch(0)
        res["train_loss"] = vt.train_one_epoch(tr, tl, 0, rank, world, log=None)
        res["val"] = list(vt.validate(tr, vl, rank, world, reference_rank_sum=True))
        res["val_mean"] = list(vt.validate(tr, vl, rank, world, reference_rank_sum=False))
        things, rdm = script.load_things(fx["things_csv"], fx["things_dir"], fx["rdm"], cpu)
        ev = ScipyEvaluator(rdm)
        loader = vt.ShardedLoader(things, 8, world, rank, with_names=True)
        res["rsa"] = list(vt.compute_rsa_score(model, loader, rdm, rank, world, dataset_order=False, evaluator=ev))
        res["rsa_dataset_order"] = list(vt.compute_rsa_score(model, loader, rdm, rank, world, dataset_order=True, evaluator=ev))
        one = vt.ShardedLoader(things, 8, 1, 0, with_names=True)
        res["rsa_single_rank"] = list(vt.compute_rsa_score(model, one, rdm, 0, 1, evaluator=ev))
    with open(os.path.join(out_dir, f"rank{rank}.json"), "w") as f:
        json.dump({k: [plain(x) for x in v] if isinstance(v, list) else plain(v) for k, v in res.items()}, f)
    dist.destroy_process_group()


def run_collectives(arm, fx, root):
    import torch.multiprocessing as mp
    out_dir = os.path.join(root, f"coll_{arm}")
    os.makedirs(out_dir)
    mp.spawn(_collective_worker, args=(2, 29100 + os.getpid() % 500, arm, fx, out_dir), nprocs=2, join=True)
    return [json.load(open(os.path.join(out_dir, f"rank{r}.json"))) for r in (0, 1)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arm", choices=["reference", "product"], required=True)
    ap.add_argument("--write-golden", action="store_true")
    a = ap.parse_args()
    warnings.filterwarnings("ignore")
    torch.set_num_threads(2)
    with tempfile.TemporaryDirectory() as root:
        fx = build_fixture(root)
        collectives = run_collectives(a.arm, fx, root)   # first: the in-process stubs below must not leak into it
        real_stdout = sys.stdout
        sys.stdout = sys.stderr                     # the reference prints its progress; keep stdout for the JSON line
        try:
            out = run_reference(fx) if a.arm == "reference" else run_product(fx)
        finally:
            sys.stdout = real_stdout
    out["arm"] = a.arm
    out["world2"] = collectives
    print(json.dumps(out))
    if a.write_golden:
        if a.arm != "reference":
            raise SystemExit("--write-golden is for the reference arm")
        path = os.path.join(ROOT, "tests", "golden", "vit_measure_exec.json")
        with open(path, "w") as f:
            json.dump(out, f)
        print("wrote", path, file=sys.stderr)


if __name__ == "__main__":
    main()

"""GPU parity tests of the ViT epoch loop and the single-epoch perturbation measurement (hba.vit_train on
hba.vit / libhba; reference VIT = Training/vit_training/baseline/train_vit_sgd.py, MEAS = Training/
vit_training/single_epoch/measure_single_epoch_perturbation_effect.py) against the CPU oracle
(oracle/vit_ref.py + oracle/vit_measure_ref.py):

  * `forward_features` (MEAS:308-322) and the no-grad validation batch (VIT:178-187);
  * the optimizer state in `torch.optim.SGD.state_dict()` layout: round trip and exchange with the oracle's
    torch optimizer in both directions (VIT:98-100, 321);
  * one whole measurement - checkpoint N-1 written by the ORACLE's baseline run, perturbed epoch, validation,
    CLS-feature RSA - against `measure_ref` (fp32 mode, stated tolerances);
  * the two drop-in scripts end to end on synthetic data (bf16 mode, CUDA-graph steps, resume).
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]

DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLASSES = 10


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _pair(seed=3):
    from hba import vit
    from oracle import vit_ref
    ref = vit_ref.create_model("vit_tiny_test", num_classes=CLASSES, seed=seed)
    with torch.no_grad():
        for p in ref.parameters():
            if p.ndim == 1:
                p.add_(torch.randn_like(p) * 0.1)
    prod = vit.create_model("vit_tiny_test", num_classes=CLASSES)
    prod.load_state_dict(ref.state_dict(), strict=True)
    return ref, prod.to(DEV)


def _batches(n, seed, bs=4):
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn(bs, 3, 224, 224, generator=g), torch.randint(0, CLASSES, (bs,), generator=g))
            for _ in range(n)]


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-3), ("bf16", 5e-2)])
def test_forward_features_and_validation_batch_match_oracle(mode, tol):
    import hba
    from hba import vit
    hba.set_precision(mode)
    try:
        ref, prod = _pair()
        (x, y), = _batches(1, seed=0, bs=5)
        with torch.no_grad():
            want = ref.forward_features(x)
            want_logits = ref(x)
        got = prod.forward_features(x.to(DEV))
        assert got.shape == want.shape == (5, 197, 128) and got.dtype == torch.float32
        assert prod.global_pool == "token"
        assert rel_err(got, want) < tol, rel_err(got, want)
        with pytest.raises(RuntimeError):
            prod.forward_features(x)                       # no CPU path
        tr = vit.DataParallelTrainer(prod)
        loss, hits = tr.evaluate(x.to(DEV), y.to(DEV))
        want_loss = float(F.cross_entropy(want_logits, y))
        assert abs(float(loss) - want_loss) < tol * abs(want_loss)
        assert abs(int(hits) - int(want_logits.max(1)[1].eq(y).sum())) <= 1          # a near-tie of two logits may flip
        # evaluation leaves the parameters and the training path untouched
        before = [p.detach().clone() for p in prod.parameters()]
        tr.evaluate(x.to(DEV), y.to(DEV))
        assert all(torch.equal(a, b.detach()) for a, b in zip(before, prod.parameters()))
    finally:
        hba.set_precision("bf16")


def test_optimizer_state_round_trip_and_exchange_with_torch_sgd():
    import hba
    from hba import vit
    hba.set_precision("fp32")
    try:
        ref, prod = _pair(seed=5)
        batches = _batches(3, seed=1)
        dev_batches = [(a.to(DEV), b.to(DEV)) for a, b in batches]
        # the oracle: torch.optim.SGD, two steps
        opt = torch.optim.SGD(ref.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
        for images, labels in batches[:2]:
            opt.zero_grad()
            F.cross_entropy(ref(images), labels).backward()
            opt.step()
        torch_state = opt.state_dict()
        ref_after2 = {k: v.detach().clone() for k, v in ref.state_dict().items()}
        # the product: two steps, then its state in the same layout
        tr_a = vit.DataParallelTrainer(prod, lr=0.1, momentum=0.9, weight_decay=1e-4)
        assert tr_a.state_dict()["state"] == {}                      # no momentum before the first step
        for images, labels in dev_batches[:2]:
            tr_a.step(images, labels)
        sd = tr_a.state_dict()
        assert sorted(sd["param_groups"][0]) == sorted(torch_state["param_groups"][0])
        assert sd["param_groups"][0]["params"] == torch_state["param_groups"][0]["params"]
        assert sorted(sd["state"]) == sorted(torch_state["state"])
        for i, p in enumerate(prod.parameters()):
            got, want = sd["state"][i]["momentum_buffer"], torch_state["state"][i]["momentum_buffer"]
            assert got.shape == p.shape and rel_err(got, want) < 5e-3, (i, rel_err(got, want))
        # (i) round trip into a fresh model + trainer: the third step is bit-identical
        prod_b = vit.create_model("vit_tiny_test", num_classes=CLASSES).to(DEV)
        prod_b.load_state_dict(prod.state_dict())
        tr_b = vit.DataParallelTrainer(prod_b, lr=0.5, momentum=0.0, weight_decay=0.0)
        tr_b.load_state_dict(sd)
        assert tr_b.param_groups[0]["lr"] == 0.1 and tr_b.momentum == 0.9 and tr_b.wd == 1e-4
        back = tr_b.state_dict()                                      # still pending: returned unchanged
        assert all(torch.equal(back["state"][i]["momentum_buffer"].cpu(), sd["state"][i]["momentum_buffer"].cpu())
                   for i in sd["state"])
        la, _ = tr_a.step(*dev_batches[2])
        lb, _ = tr_b.step(*dev_batches[2])
        assert float(la) == float(lb)
        for (n, a), (_, b) in zip(prod.named_parameters(), prod_b.named_parameters()):
            assert torch.equal(a.detach(), b.detach()), n
        for i, (a, b) in enumerate(zip(prod.parameters(), prod_b.parameters())):
            assert torch.equal(tr_a.state_dict()["state"][i]["momentum_buffer"],
                               tr_b.state_dict()["state"][i]["momentum_buffer"])
        # (ii) the oracle's torch optimizer state + weights into the product: third step against torch's
        prod_c = vit.create_model("vit_tiny_test", num_classes=CLASSES).to(DEV)
        prod_c.load_state_dict(ref_after2)
        tr_c = vit.DataParallelTrainer(prod_c)
        tr_c.load_state_dict(torch_state)
        lc, _ = tr_c.step(*dev_batches[2])
        opt.zero_grad()
        lo = F.cross_entropy(ref(batches[2][0]), batches[2][1])
        lo.backward()
        opt.step()
        lo = float(lo.detach())
        assert abs(float(lc) - lo) < 1e-3 * abs(lo)
        for (n, a), (_, b) in zip(prod_c.named_parameters(), ref.named_parameters()):
            assert rel_err(a.detach(), b.detach()) < 2e-3, (n, rel_err(a.detach(), b.detach()))
        # (iii) the product's state into a torch optimizer (the reference side loading our checkpoint)
        opt2 = torch.optim.SGD(ref.parameters(), lr=1.0)
        opt2.load_state_dict({"state": {k: {"momentum_buffer": v["momentum_buffer"].cpu()} for k, v in sd["state"].items()},
                              "param_groups": sd["param_groups"]})
        assert opt2.param_groups[0]["lr"] == 0.1 and len(opt2.state_dict()["state"]) == len(sd["state"])
        # malformed states are refused
        with pytest.raises(ValueError):
            tr_c.load_state_dict({"state": {}, "param_groups": [dict(sd["param_groups"][0], params=[0, 1])]})
        with pytest.raises(NotImplementedError):
            tr_c.load_state_dict({"state": {}, "param_groups": [dict(sd["param_groups"][0], nesterov=True)]})
    finally:
        hba.set_precision("bf16")


def _oracle_baseline(tmp_path, train, val, epochs=3, batch=4):
    """Baseline run by the ORACLE (CPU, torch SGD) in the reference's checkpoint format (VIT:89-123)."""
    import pandas as pd
    from torch.utils.data import TensorDataset
    from oracle import vit_measure_ref as ref
    from oracle import vit_ref
    model = vit_ref.create_model("vit_tiny_test", num_classes=CLASSES, seed=11)
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
    sched = ref.CosineWarmupRef(opt, 5, 100, eta_min=0)
    d = os.path.join(str(tmp_path), "baseline")
    os.makedirs(d)
    rows = []
    for epoch in range(epochs):
        tl = ref.train_one_epoch_ref(model, ref.rank_loader(TensorDataset(*train), batch, 1, 0, True, epoch=epoch), opt)
        sched.step()
        vl, va = ref.reduce_validation([ref.validate_rank_ref(model, ref.rank_loader(TensorDataset(*val), batch, 1, 0, False))])
        torch.save({"epoch": epoch, "model_state_dict": model.state_dict(), "optimizer_state_dict": opt.state_dict(),
                    "scheduler_state_dict": sched.state_dict(), "scaler_state_dict": {}, "train_loss": tl,
                    "val_loss": vl, "val_acc": va}, os.path.join(d, f"checkpoint_epoch_{epoch:03d}.pth"))
        rows.append({"epoch": epoch, "train_loss": tl, "val_loss": vl, "val_acc": va, "rsa_score": 0.1 * (epoch + 1)})
    csv = os.path.join(d, "metrics_with_rsa.csv")
    pd.DataFrame(rows).to_csv(csv, index=False)
    return d, csv


def _problem(seed=0, n_train=12, n_val=8, n_things=48):
    g = torch.Generator().manual_seed(seed)
    mk = lambda n: (torch.randn(n, 3, 224, 224, generator=g), torch.randint(0, CLASSES, (n,), generator=g))
    train, val = mk(n_train), mk(n_val)
    things = torch.randn(n_things, 3, 224, 224, generator=g)
    rdm = 1 - np.corrcoef(torch.randn(n_things, 66, generator=g).double().numpy())
    np.fill_diagonal(rdm, 0)
    return train, val, things, rdm


@pytest.mark.parametrize("kind", ["uniform_gray", "label_shuffle"])
def test_measurement_matches_oracle(tmp_path, kind):
    """MEAS:403-555 on the GPU against the oracle restatement, fp32 mode: validation loss after the perturbed
    epoch within 5e-3 relative, RSA rho within 0.05 absolute of the oracle's (the rho of 1,128 ranked RDM
    entries moves with every rank swap the 1e-3 embedding differences cause); the RSA tail itself is exact:
    rho equals scipy's on the product's own features within 1e-9."""
    import hba
    from hba import vit_train as vt
    from oracle import vit_measure_ref as ref
    from oracle import vit_ref
    hba.set_precision("fp32")
    try:
        train, val, things, rdm = _problem()
        ckdir, csv = _oracle_baseline(tmp_path, train, val)
        to_dev = lambda t: t.to(DEV)
        got = vt.measure_perturbation_effect(
            2, kind, ckdir, csv, vt.ResidentImageSet(*map(to_dev, train)), vt.ResidentImageSet(*map(to_dev, val)),
            vt.ResidentImageSet(things.to(DEV)), rdm, batch_size=4, model_name="vit_tiny_test", num_classes=CLASSES,
            use_graph=False, log=None)   # host-launched steps here; the captured step runs in the script test below
        ck = torch.load(os.path.join(ckdir, "checkpoint_epoch_001.pth"), weights_only=False)
        factory = lambda: vit_ref.create_model("vit_tiny_test", num_classes=CLASSES)
        want = ref.measure_ref(ck, factory, train, val, things, rdm, 2, kind, got["baseline_loss"], got["baseline_rsa"],
                               batch_size=4, num_classes=CLASSES)
        assert list(got) == list(vt.RESULT_COLUMNS)
        assert got["baseline_rsa"] == pytest.approx(0.3) and got["perturb_epoch"] == 2 and got["perturbation_type"] == kind
        assert abs(got["perturbed_loss"] - want["perturbed_loss"]) < 5e-3 * abs(want["perturbed_loss"])
        assert abs(got["perturbed_rsa"] - want["perturbed_rsa"]) < 0.05
        assert got["delta_loss"] == got["perturbed_loss"] - got["baseline_loss"]
        assert got["delta_rsa"] == got["perturbed_rsa"] - got["baseline_rsa"]
    finally:
        hba.set_precision("bf16")


def test_rsa_of_cls_features_equals_scipy_on_the_same_features():
    """compute_rsa_score (MEAS:298-355): features -> libhba RDM / ranks / Spearman against NumPy + SciPy on the
    very same feature matrix: rho within 1e-9, p-value within 1e-6 relative."""
    from hba import vit_train as vt
    from oracle import vit_measure_ref as ref
    _, prod = _pair(seed=9)
    _, _, things, rdm = _problem(seed=2)
    data = vt.ResidentImageSet(things.to(DEV))
    loader = vt.ShardedLoader(data, 8, with_names=True)
    rho, p = vt.compute_rsa_score(prod, loader, rdm)
    feats = torch.cat([prod.forward_features(x)[:, 0] for _, x in loader]).cpu().numpy()
    want_rho, want_p = ref.rsa_tail_ref(feats, rdm)
    assert abs(rho - want_rho) < 1e-9
    assert abs(p - want_p) <= 1e-6 * abs(want_p) + 1e-300


def test_gaussian_measurement_is_reproducible_with_a_noise_seed(tmp_path):
    import hba
    from hba import vit_train as vt
    hba.set_precision("fp32")
    try:
        train, val, things, rdm = _problem(seed=3, n_things=48)
        ckdir, csv = _oracle_baseline(tmp_path, train, val, epochs=2)
        to_dev = lambda t: t.to(DEV)
        kw = dict(baseline_checkpoint_dir=ckdir, baseline_metrics_csv=csv, train_data=vt.ResidentImageSet(*map(to_dev, train)),
                  val_data=vt.ResidentImageSet(*map(to_dev, val)), things_data=vt.ResidentImageSet(things.to(DEV)),
                  things_rdm=rdm, batch_size=4, model_name="vit_tiny_test", num_classes=CLASSES, use_graph=False, log=None)
        a = vt.measure_perturbation_effect(1, "gaussian", noise_seed=5, **kw)
        b = vt.measure_perturbation_effect(1, "gaussian", noise_seed=5, **kw)
        c = vt.measure_perturbation_effect(1, "gaussian", noise_seed=6, **kw)
        assert a == b and np.isfinite(a["perturbed_loss"]) and np.isfinite(a["perturbed_rsa"])
        assert c["perturbed_loss"] != a["perturbed_loss"]
    finally:
        hba.set_precision("bf16")


def _load_script(rel):
    import importlib.util
    path = os.path.join(ROOT, "vit-project_b200", "vit_training", rel)
    spec = importlib.util.spec_from_file_location("_script_" + os.path.basename(rel)[:-3], path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_drop_in_scripts_end_to_end_on_synthetic_data(tmp_path):
    """train_vit_sgd.py (2 epochs, then resumed for a third) and measure_single_epoch_perturbation_effect.py on
    HBM-resident synthetic data, bf16 mode, captured steps: the reference's files and CSV schemas."""
    import pandas as pd
    train_script = _load_script("baseline/train_vit_sgd.py")
    measure_script = _load_script("single_epoch/measure_single_epoch_perturbation_effect.py")
    out = os.path.join(str(tmp_path), "run")
    common = ["--data_path", f"synthetic:16:8:{CLASSES}", "--output_dir", out, "--batch_size", "4", "--model", "vit_tiny_test"]
    train_script.main(common + ["--epochs", "2"])
    m = pd.read_csv(os.path.join(out, "training_metrics.csv"))
    assert list(m.columns) == ["epoch", "train_loss", "val_loss", "val_acc"] and m["epoch"].tolist() == [0, 1]
    assert np.isfinite(m[["train_loss", "val_loss"]].to_numpy()).all()
    train_script.main(common + ["--epochs", "3"])                       # resumes from checkpoint_latest.pth (VIT:314-328)
    m = pd.read_csv(os.path.join(out, "training_metrics.csv"))
    assert m["epoch"].tolist() == [0, 1, 2]
    files = sorted(os.listdir(out))
    assert files == ["checkpoint_epoch_000.pth", "checkpoint_epoch_001.pth", "checkpoint_epoch_002.pth",
                     "checkpoint_latest.pth", "training_metrics.csv"]
    ck = torch.load(os.path.join(out, "checkpoint_epoch_001.pth"), weights_only=False)
    assert ck["scheduler_state_dict"]["current_epoch"] == 2
    assert ck["optimizer_state_dict"]["param_groups"][0]["lr"] == pytest.approx(0.1 * 2 / 5)
    assert len(ck["optimizer_state_dict"]["state"]) == len(ck["model_state_dict"])
    # RSA of every checkpoint -> the measurement's baseline CSV (schema of the shipped rsa_results_final.csv)
    from hba import vit_train as vt
    things, rdm = vt.synthetic_things(DEV)
    base_csv = os.path.join(out, "rsa_results.csv")
    rows = vt.rsa_over_checkpoints(out, things, rdm, base_csv, model_name="vit_tiny_test", num_classes=CLASSES, log=None)
    b = pd.read_csv(base_csv)
    assert list(b.columns) == ["checkpoint", "epoch", "train_loss", "val_loss", "val_acc", "rsa_score"]
    assert b["epoch"].tolist() == [0, 1, 2] and np.isfinite(b["rsa_score"].to_numpy()).all()
    assert np.allclose(b["val_loss"], m["val_loss"], atol=1e-6) and len(rows) == 3
    res_csv = os.path.join(str(tmp_path), "effects.csv")
    measure_script.main(["--baseline_checkpoint_dir", out, "--baseline_metrics_csv", base_csv, "--data_path",
                         f"synthetic:16:8:{CLASSES}", "--output_csv", res_csv, "--things_csv", "synthetic",
                         "--perturbation_types", "label_shuffle", "uniform_gray", "--perturb_epochs", "0", "2", "9",
                         "--batch_size", "4", "--model", "vit_tiny_test"])
    df = pd.read_csv(res_csv)
    assert list(df.columns) == ["perturb_epoch", "perturbation_type", "baseline_loss", "baseline_rsa", "perturbed_loss",
                                "perturbed_rsa", "delta_loss", "delta_rsa"]
    assert df["perturb_epoch"].tolist() == [2, 2] and df["perturbation_type"].tolist() == ["label_shuffle", "uniform_gray"]
    assert np.isfinite(df[["perturbed_loss", "perturbed_rsa"]].to_numpy()).all()
    assert np.allclose(df["baseline_rsa"], b["rsa_score"][2], rtol=1e-12) and np.allclose(df["baseline_loss"], b["val_loss"][2], rtol=1e-12)
    assert (df["delta_loss"] - (df["perturbed_loss"] - df["baseline_loss"])).abs().max() < 1e-9

"""GPU tests of the drop-in pipeline (functions/*) against goldens produced by the REFERENCE'S OWN
train_model / CLIPHBA on the CPU (tests/golden, oracle/make_golden.py): forward, loss, gradients and a
3-epoch training trajectory with a perturbation window, CSV + checkpoint formats."""
import csv
import os

import numpy as np
import pytest
import scipy.io
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.fixture()
def tiny_checkpoint(tmp_path, monkeypatch):
    """Make the plug-in clip module serve the oracle's seeded ViT-tiny weights (what the golden used)."""
    from oracle import clip_ref
    from src.models.CLIPs.clip_hba import clip
    path = tmp_path / "ViT-tiny-14.pt"
    torch.save(clip_ref.synthetic_state_dict("ViT-tiny/14", seed=1), path)
    monkeypatch.setattr(clip, "_download", lambda url, root: str(path))
    return path


def build_model(NEW):
    from oracle.synth import PROMPTS
    model = NEW.CLIPHBA(PROMPTS, backbone_name="ViT-tiny/14", pos_embedding=True)
    torch.manual_seed(123)
    NEW.apply_dora_to_ViT(model, n_vision_layers=2, n_transformer_layers=1, r=8, dora_dropout=0.1)
    NEW.switch_dora_layers(model, freeze_all=True, dora_state=True)
    return model


def test_forward_loss_grads_match_reference_golden(tiny_checkpoint):
    import hba
    import functions.new_cvpr_train_behavior_things_pipeline as NEW
    from oracle.synth import synthetic_problem
    g = torch.load(os.path.join(GOLD, "tiny_clip_forward.pt"), weights_only=False)
    prob = synthetic_problem()
    hba.set_precision("fp32")
    try:
        model = build_model(NEW)
        assert NEW.count_trainable_parameters(model) == g["n_trainable"]
        for n, p in model.named_parameters():
            if p.requires_grad:
                assert torch.equal(p.detach(), g["dora_init"][n]), n   # same RNG consumption as the reference
        model.to(DEV)
        x, y = prob["train_images"][:3].to(DEV), prob["train_targets"][:3].to(DEV)
        pred = model(x)
        loss = torch.nn.MSELoss()(pred, y)
        loss.backward()
        assert rel_err(pred, g["pred"]) < 1e-3
        assert abs(float(loss) - g["loss"]) / abs(g["loss"]) < 1e-3
        for n, p in model.named_parameters():
            if p.requires_grad:
                assert rel_err(p.grad, g["grads"][n]) < 1e-3, n
    finally:
        hba.set_precision("bf16")


def test_training_trajectory_matches_reference_golden(tiny_checkpoint, tmp_path):
    import hba
    import functions.new_cvpr_train_behavior_things_pipeline as NEW
    from oracle.synth import ListDataset, synthetic_problem
    g = torch.load(os.path.join(GOLD, "tiny_training.pt"), weights_only=False)
    prob = synthetic_problem()
    hba.set_precision("fp32")
    try:
        mat = str(tmp_path / "RDM48_triplet.mat")
        scipy.io.savemat(mat, {"RDM48_triplet": prob["human_rdm"]})
        NEW.seed_everything(1)
        model = build_model(NEW)
        model.to(DEV)
        tr = ListDataset([(f"tr{i}", prob["train_images"][i], prob["train_targets"][i]) for i in range(16)])
        te = ListDataset([(f"te{i}", prob["test_images"][i], prob["test_targets"][i]) for i in range(8)])
        rs = ListDataset([(f"rs{i}", prob["rsa_images"][i]) for i in range(8)], RDM48_triplet_dir=mat)
        gen = torch.Generator()
        gen.manual_seed(1)
        tl = torch.utils.data.DataLoader(tr, batch_size=8, shuffle=True, generator=gen)
        el = torch.utils.data.DataLoader(te, batch_size=8, shuffle=False)
        rl = torch.utils.data.DataLoader(rs, batch_size=8, shuffle=False)
        opt = NEW.make_optimizer(model, 3e-4)
        res = str(tmp_path / "res.csv")
        NEW.train_model(model, tl, el, rl, DEV, opt, torch.nn.MSELoss(), epochs=3, training_res_path=res,
                        training_run=2, perturb_length=1, perturb_seed=42, mean=5.75, std=9.5,
                        perturb_distribution="target", perturb_type="uniform_images", logger=None,
                        early_stopping_patience=10, dora_parameters_path=str(tmp_path / "dora"),
                        random_state_path=str(tmp_path / "rand"), dataloader_generator=gen)
        rows = list(csv.reader(open(res)))
        want = g["csv_rows"]
        assert rows[0] == want[0] and len(rows) == len(want) == 4
        for got, exp in zip(rows[1:], want[1:]):
            assert got[0] == exp[0] and got[5:] == exp[5:]          # epoch, perturbation flags
            for k in (1, 2):                                       # train / test loss: 1e-3 relative
                assert abs(float(got[k]) - float(exp[k])) / abs(float(exp[k])) < 1e-3, (got, exp)
            assert abs(float(got[3]) - float(exp[3])) < 1e-4, (got, exp)   # RSA rho within 1e-4
            assert abs(float(got[4]) - float(exp[4])) < 1e-3 * max(float(exp[4]), 1e-12) + 1e-9
        assert rows[2][7] == "True"                                 # uniform images used in epoch 2
        ck = torch.load(tmp_path / "dora" / "epoch3_dora_params.pth")
        assert set(ck) == set(g["dora_epoch3"])
        for k in ck:
            assert rel_err(ck[k], g["dora_epoch3"][k]) < 1e-3, k
        rstate = torch.load(tmp_path / "rand" / "epoch3_random_states.pth", weights_only=False)
        assert sorted(rstate["optimizer_state_dict"]["state"].keys()) == g["optimizer_state_keys"]
        assert set(g["random_state_keys"]) <= set(rstate.keys()) | {"cuda_rng_state", "cuda_rng_state_all"}
    finally:
        hba.set_precision("bf16")


def test_behavioral_rsa_matches_reference_golden(tmp_path):
    import functions.new_cvpr_train_behavior_things_pipeline as NEW
    from oracle.synth import ListDataset
    g = torch.load(os.path.join(GOLD, "rsa_tail.pt"), weights_only=False)
    mat = str(tmp_path / "RDM48_triplet.mat")
    scipy.io.savemat(mat, {"RDM48_triplet": g["human_rdm"]})
    ds = ListDataset([(f"img{i}", g["emb"][i]) for i in range(48)], RDM48_triplet_dir=mat)
    loader = torch.utils.data.DataLoader(ds, batch_size=32, shuffle=False)
    rho, p, rdm = NEW.behavioral_RSA(torch.nn.Identity(), loader, DEV)
    assert abs(rho - g["rho"]) < 1e-12
    assert abs(p - g["p"]) <= 1e-9 * g["p"] + 1e-300
    assert np.abs(rdm - g["model_rdm"]).max() < 1e-12


def test_random_target_and_shuffle_are_seed_deterministic_on_device():
    from functions import _pipeline_core as core
    t = torch.arange(48, dtype=torch.float32, device=DEV).reshape(8, 6)
    for kind in ("random_target", "label_shuffle"):
        p = core.Perturbation(kind, 3, 1, 42, "target", 5.75, 9.5)
        a = p.apply(None, t, 1, DEV)[1]
        b = p.apply(None, t, 1, DEV)[1]
        c = p.apply(None, t, 2, DEV)[1]
        assert torch.equal(a, b) and not torch.equal(a, c)
    gen = torch.Generator(device=DEV)
    gen.manual_seed(42 + 3 * 1000 + 1)   # NEW:919-926
    want = torch.randn(t.shape, device=DEV, generator=gen) * 9.5 + 5.75
    got = core.Perturbation("random_target", 3, 1, 42, "target", 5.75, 9.5).apply(None, t, 1, DEV)[1]
    assert torch.equal(got, want)


def test_trunk_cache_is_bit_identical(tiny_checkpoint):
    """Frozen-trunk cache (north star item 2): the cached step equals recomputation bit for bit."""
    import hba
    import functions.new_cvpr_train_behavior_things_pipeline as NEW
    from hba.engine import TrunkCache
    from oracle.synth import synthetic_problem
    prob = synthetic_problem()
    x = prob["train_images"][:4].to(DEV)
    y66 = torch.randn(4, 6, generator=torch.Generator().manual_seed(3)).to(DEV)
    for precision in ("bf16", "fp32"):
        hba.set_precision(precision)
        model = build_model(NEW).to(DEV)
        eng = model.clip_model.hba_engine()
        crit = torch.nn.MSELoss()

        def run(ids):
            for p in model.parameters():
                p.grad = None
            eng.batch_ids = ids
            pred = model(x)
            crit(pred, y66).backward()
            return pred.detach().clone(), [p.grad.clone() for p in model.parameters() if p.requires_grad]

        base_pred, base_grads = run(None)
        eng.trunk_cache = TrunkCache(16)
        fill_pred, fill_grads = run([3, 7, 1, 9])          # miss: computes the trunk and stores it
        assert eng.trunk_cache.present == {1, 3, 7, 9}
        hit_pred, hit_grads = run([3, 7, 1, 9])            # hit: trunk skipped
        assert torch.equal(base_pred, fill_pred) and torch.equal(base_pred, hit_pred)
        for a, b, c in zip(base_grads, fill_grads, hit_grads):
            assert torch.equal(a, b) and torch.equal(a, c)
        with torch.no_grad():                              # eval path, permuted subset of cached ids
            eng.batch_ids = [9, 3]
            sub = model(x[[3, 0]])
        assert torch.equal(sub, base_pred[[3, 0]])
    hba.set_precision("bf16")


@pytest.mark.parametrize("B", [1, 3, 5])
def test_trunk_cache_hit_is_independent_of_the_batch_that_filled_it(tiny_checkpoint, B):
    """A sweep worker fills the cache once and serves later conditions (other batch compositions, a last
    batch of one image) from it: a hit on B images cached as part of an 8-image batch equals recomputation
    bit for bit - predictions and DoRA gradients."""
    import hba
    import functions.new_cvpr_train_behavior_things_pipeline as NEW
    from hba.engine import TrunkCache
    from oracle.synth import synthetic_problem
    hba.set_precision("bf16")
    prob = synthetic_problem()
    xall = prob["train_images"][:8].to(DEV)
    y = torch.randn(8, 6, generator=torch.Generator().manual_seed(3)).to(DEV)
    model = build_model(NEW).to(DEV)
    eng = model.clip_model.hba_engine()
    crit = torch.nn.MSELoss()

    def run(x, t, ids):
        for p in model.parameters():
            p.grad = None
        eng.batch_ids = ids
        pred = model(x)
        crit(pred, t).backward()
        return pred.detach().clone(), [p.grad.clone() for p in model.parameters() if p.requires_grad]

    eng.trunk_cache = None
    base = run(xall[:B], y[:B], None)
    eng.trunk_cache = TrunkCache(16)
    run(xall, y, list(range(8)))                 # fill: all 8 images in one batch
    hit = run(xall[:B], y[:B], list(range(B)))
    assert torch.equal(base[0], hit[0])
    for a, b in zip(base[1], hit[1]):
        assert torch.equal(a, b)
    eng.trunk_cache = None


def _write_things_like_dataset(root, n_train=12, n_rsa=8, seed=0):
    """A THINGS-shaped dataset on disk: PNG images, the SPoSE csv layout (index, image name, 66 target
    columns; NEW:191-202), the 48-image-style inference csv and RDM48_triplet.mat."""
    import pandas as pd
    from PIL import Image
    rng = np.random.default_rng(seed)
    img_dir = os.path.join(root, "imgs")
    os.makedirs(img_dir, exist_ok=True)

    def make(prefix, n):
        names = []
        for i in range(n):
            name = f"{prefix}{i:03d}.png"
            Image.fromarray(rng.integers(0, 255, (40, 52, 3), dtype=np.uint8)).save(os.path.join(img_dir, name))
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
    rdm = 1 - np.corrcoef(rng.standard_normal((n_rsa, 10)))
    np.fill_diagonal(rdm, 0)
    scipy.io.savemat(os.path.join(root, "RDM48_triplet.mat"), {"RDM48_triplet": rdm})
    return img_dir


def test_run_behavioral_training_drivers_end_to_end(tiny_checkpoint, tmp_path):
    """The config-dict contract of BDRV:11-33 / SWEEP:118-147 end to end on files: baseline run
    (split + per-epoch checkpoints), then a sweep condition that resumes from baseline epoch 1 with a
    random-target window — once with the HBM-resident loaders + trunk cache, once with the plain
    DataLoader path: identical CSV rows."""
    import hba
    import functions.cvpr_train_behavior_things_pipeline_baseline as BASE
    import functions.new_cvpr_train_behavior_things_pipeline as NEW
    hba.set_precision("bf16")
    root = str(tmp_path)
    img_dir = _write_things_like_dataset(root)
    common = {"csv_file": f"{root}/train.csv", "img_dir": img_dir, "inference_csv_file": f"{root}/rsa.csv",
              "RDM48_triplet_dir": f"{root}/RDM48_triplet.mat", "backbone": "ViT-tiny/14", "batch_size": 4,
              "lr": 3e-4, "random_seed": 1, "vision_layers": 2, "transformer_layers": 1, "rank": 8,
              "criterion": torch.nn.MSELoss(), "cuda": 0}
    base_cfg = dict(common, epochs=2, train_portion=0.75, early_stopping_patience=20, logger=None,
                    checkpoint_path=f"{root}/base/model.pth", training_res_path=f"{root}/base/res.csv",
                    dora_parameters_path=f"{root}/base/dora", random_state_path=f"{root}/base/rand")
    BASE.run_behavioral_training(base_cfg)
    rows = list(csv.reader(open(f"{root}/base/res.csv")))
    assert rows[0] == ["epoch", "train_loss", "test_loss", "behavioral_rsa_rho", "behavioral_rsa_p_value"]
    assert [r[0] for r in rows[1:]] == ["1", "2"]
    split = torch.load(f"{root}/base/rand/dataset_split_indices.pth")
    assert len(split["train_indices"]) == 9 and len(split["test_indices"]) == 3
    assert os.path.exists(f"{root}/base/dora/epoch2_dora_params.pth")
    assert os.path.exists(f"{root}/base/rand/epoch2_random_states.pth")

    results = {}
    for tag, resident in (("resident", True), ("plain", False)):
        cfg = dict(common, epochs=3, early_stopping_patience=10, hba_resident=resident,
                   checkpoint_path=f"{root}/{tag}/model.pth", training_res_path=f"{root}/{tag}/res.csv",
                   dora_parameters_path=f"{root}/{tag}/dora", random_state_path=f"{root}/{tag}/rand",
                   baseline_dora_directory=f"{root}/base/dora", baseline_random_state_path=f"{root}/base/rand",
                   baseline_split_indices_path=f"{root}/base/rand/dataset_split_indices.pth",
                   perturb_type="random_target", perturb_length=1, perturb_distribution="target",
                   perturb_seed=42, training_run=2, resume_from_epoch=1,
                   previous_training_res_path=f"{root}/base/res.csv")
        NEW.run_behavioral_training(cfg)
        results[tag] = list(csv.reader(open(cfg["training_res_path"])))
    r = results["resident"]
    assert r[0][5:] == ["used_random_targets", "used_shuffled_targets", "used_uniform_images", "used_image_noise"]
    assert [row[0] for row in r[1:]] == ["1", "2", "3"]          # epoch 1 pre-populated from the baseline
    assert r[1][:5] == rows[1][:5]
    assert r[2][5] == "True" and r[3][5] == "False"              # random targets in epoch 2 only
    assert r == results["plain"]                                 # resident + cached == plain, exactly


@pytest.mark.parametrize("perturb_type", ["random_target", "uniform_images"])
def test_captured_step_graphs_are_bit_identical_to_eager_steps(tiny_checkpoint, tmp_path, monkeypatch, perturb_type):
    """HBA_STEP_GRAPH (CUDA graphs of the trunk-cached training step, of the full-trunk step used under
    image perturbations, and of the cached eval / RSA forward) against the same run launched kernel by
    kernel: a sweep condition resumed from baseline epoch 1, 7 epochs with a 2-epoch perturbation window -
    identical CSV rows, DoRA checkpoints and optimizer step counts."""
    import hba
    import functions.cvpr_train_behavior_things_pipeline_baseline as BASE
    import functions.new_cvpr_train_behavior_things_pipeline as NEW
    hba.set_precision("bf16")
    root = str(tmp_path)
    img_dir = _write_things_like_dataset(root, n_train=22)
    common = {"csv_file": f"{root}/train.csv", "img_dir": img_dir, "inference_csv_file": f"{root}/rsa.csv",
              "RDM48_triplet_dir": f"{root}/RDM48_triplet.mat", "backbone": "ViT-tiny/14", "batch_size": 4,
              "lr": 3e-4, "random_seed": 1, "vision_layers": 2, "transformer_layers": 1, "rank": 8,
              "criterion": torch.nn.MSELoss(), "cuda": 0}
    base_cfg = dict(common, epochs=1, train_portion=0.8, early_stopping_patience=20, logger=None,
                    checkpoint_path=f"{root}/base/model.pth", training_res_path=f"{root}/base/res.csv",
                    dora_parameters_path=f"{root}/base/dora", random_state_path=f"{root}/base/rand")
    BASE.run_behavioral_training(base_cfg)
    results = {}
    for tag, flag in (("graph", "1"), ("eager", "0")):
        monkeypatch.setenv("HBA_STEP_GRAPH", flag)
        cfg = dict(common, epochs=7, early_stopping_patience=20, hba_resident=True,
                   checkpoint_path=f"{root}/{tag}/model.pth", training_res_path=f"{root}/{tag}/res.csv",
                   dora_parameters_path=f"{root}/{tag}/dora", random_state_path=f"{root}/{tag}/rand",
                   baseline_dora_directory=f"{root}/base/dora", baseline_random_state_path=f"{root}/base/rand",
                   baseline_split_indices_path=f"{root}/base/rand/dataset_split_indices.pth",
                   perturb_type=perturb_type, perturb_length=2, perturb_distribution="target",
                   perturb_seed=42, training_run=4, resume_from_epoch=1,
                   previous_training_res_path=f"{root}/base/res.csv")
        NEW.run_behavioral_training(cfg)
        rows = list(csv.reader(open(cfg["training_res_path"])))
        dora = torch.load(f"{root}/{tag}/dora/epoch7_dora_params.pth")
        rand = torch.load(f"{root}/{tag}/rand/epoch7_random_states.pth", weights_only=False)
        results[tag] = (rows, dora, rand)
    g, e = results["graph"], results["eager"]
    assert len(g[0]) == 8 and g[0] == e[0]
    col = {"random_target": 5, "uniform_images": 7}[perturb_type]   # used_random_targets / used_uniform_images
    assert [r[col] for r in g[0][2:]] == ["False", "False", "True", "True", "False", "False"]   # epochs 2..7
    assert g[1].keys() == e[1].keys()
    for k in g[1]:
        assert torch.equal(g[1][k], e[1][k]), k
    steps_g = [float(v["step"]) for v in g[2]["optimizer_state_dict"]["state"].values()]
    steps_e = [float(v["step"]) for v in e[2]["optimizer_state_dict"]["state"].values()]
    assert steps_g == steps_e and steps_g[0] == 7 * 5   # 17 train images / batch 4 -> 5 steps per epoch, 7 epochs


def test_fused_mse_step_equals_the_criterion_called_step(tiny_checkpoint, monkeypatch):
    """nn.MSELoss fused into the head kernel (hba_cos_mse_fwd / _bwd: loss, NaN guard, loss bookkeeping, head
    backward from (pred, target)) against the same step with the caller's criterion run by torch
    (HBA_FUSED_MSE=0): losses within 1e-6 relative (different fp32 summation order of the 5 x 3 squared errors),
    DoRA parameters within 1e-5 after 6 steps; a NaN target batch is skipped and counted in both forms;
    evaluate_model returns the same mean."""
    import hba
    import functions.new_cvpr_train_behavior_things_pipeline as NEW
    from functions import _pipeline_core as core
    from oracle.synth import synthetic_problem
    prob = synthetic_problem()
    hba.set_precision("fp32")
    try:
        x = prob["train_images"][:6].to(DEV)
        y = prob["train_targets"][:6].to(DEV)
        ybad = y[:3].clone()
        ybad[1, 2] = float("nan")
        crit = torch.nn.MSELoss()
        results = []
        for fused in ("1", "0"):
            monkeypatch.setenv("HBA_FUSED_MSE", fused)
            monkeypatch.setenv("HBA_STEP_GRAPH", "1")
            model = build_model(NEW).to(DEV)
            opt = core.make_optimizer(model, 3e-3)
            step = core.TrainStep(model, opt, crit, DEV)
            step.start_epoch()
            losses = []
            for i in range(6):          # eager warm-up, capture, replays
                step(x[3 * (i % 2):3 * (i % 2) + 3], y[3 * (i % 2):3 * (i % 2) + 3])
                losses.append(float(step.last_loss))
            before = [p.detach().clone() for p in model.parameters() if p.requires_grad]
            step(x[:3], ybad)
            assert int(step.guard.total) == 1
            for a, b in zip(before, [p for p in model.parameters() if p.requires_grad]):
                assert torch.equal(a, b.detach())            # the optimiser skipped the bad batch
            total = float(step.total)

            class _L:   # minimal loader: (names, images, targets) batches
                dataset = list(range(6))

                def __iter__(self):
                    return iter([(None, x[:3], y[:3]), (None, x[3:], y[3:])])

                def __len__(self):
                    return 2
            ev = core.evaluate_model(model, _L(), DEV, crit)
            results.append((losses, before, total, ev))
        (l1, p1, t1, e1), (l0, p0, t0, e0) = results
        assert l1 == pytest.approx(l0, rel=1e-6)
        assert t1 == pytest.approx(t0, rel=1e-6) and t1 == pytest.approx(3 * sum(l1), rel=1e-6)
        assert e1 == pytest.approx(e0, rel=1e-6)
        for a, b in zip(p1, p0):
            assert rel_err(a, b) < 1e-5
    finally:
        hba.set_precision("bf16")


def test_trunk_cache_is_never_served_after_the_engine_was_restaged(tiny_checkpoint):
    """a cache filled under one staging (precision mode) must not answer a hit after the engine was
    restaged - TrainStep decides `cached` before the forward pass, so the check has to compare the stamps, and a
    device-id lookup on a freshly reallocated cache must raise instead of returning uninitialised memory."""
    import hba
    import functions.new_cvpr_train_behavior_things_pipeline as NEW
    from functions import _pipeline_core as core
    from hba.engine import TrunkCache
    from oracle.synth import synthetic_problem
    prob = synthetic_problem()
    x = prob["train_images"][:4].to(DEV)
    y = torch.randn(4, 6, generator=torch.Generator().manual_seed(3)).to(DEV)
    ids, ids_dev = [3, 7, 1, 9], torch.tensor([3, 7, 1, 9], device=DEV)
    crit = torch.nn.MSELoss()
    hba.set_precision("bf16")
    try:
        model = build_model(NEW).to(DEV)
        eng = model.clip_model.hba_engine()
        eng.trunk_cache = TrunkCache(16)
        step = core.TrainStep(model, core.make_optimizer(model, 3e-4), crit, DEV)
        step(x, y, ids, ids_dev)                                  # fills the cache (bf16 staging)
        assert eng.trunk_cache.all_present(ids, eng)
        hba.set_precision("fp32")                                 # the next forward restages the engine
        assert not eng.trunk_cache.all_present(ids, eng)          # stale: not a hit
        eng.ensure(eng.device)
        assert not eng.trunk_cache.all_present(ids, eng)
        with pytest.raises(RuntimeError, match="restaged"):
            eng.trunk_cache.lookup_device_ids(eng, ids_dev)
        assert eng.trunk_cache.present == set()
        # the forward recomputes (and refills) instead of reading the old buffers
        with torch.no_grad():
            eng.batch_ids = None
            base = model(x)
            eng.batch_ids = ids
            fill = model(x)
            assert torch.equal(base, fill)
            assert eng.trunk_cache.all_present(ids, eng)
            eng.batch_ids = ids_dev
            assert torch.equal(base, model(x))
        assert eng.trunk_cache.all_present(ids, eng)
    finally:
        hba.set_precision("bf16")

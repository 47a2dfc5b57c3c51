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

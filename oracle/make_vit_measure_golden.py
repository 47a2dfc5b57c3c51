"""Generates tests/golden/vit_measure.json with the REFERENCE'S OWN code and data:

  * the label perturbation wrappers and the LR schedule, from the classes imported out of
    /root/reference/Training/vit_training/{single_epoch/measure_single_epoch_perturbation_effect.py,
    baseline/train_vit_sgd.py} (`timm`, which neither script needs for these classes, is stubbed);
  * the checkpoint keys and metrics-CSV text of the reference's `save_checkpoint` (VIT:89-123), run on CPU
    with a tiny stand-in module;
  * the shipped measurement rows Data/vit_results/perturbation_effects.csv (inputs: perturbed loss / RSA and
    the baseline values; outputs: the deltas the reference wrote).

    python oracle/make_vit_measure_golden.py
"""
import importlib.util
import json
import os
import re
import sys
import tempfile
import types

import pandas as pd
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def load_reference_scripts():
    sys.modules.setdefault("timm", types.ModuleType("timm"))
    mods = {}
    for name, rel in (("MEAS", "Training/vit_training/single_epoch/measure_single_epoch_perturbation_effect.py"),
                      ("VIT", "Training/vit_training/baseline/train_vit_sgd.py")):
        spec = importlib.util.spec_from_file_location(f"_ref_{name}", os.path.join(REF, rel))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        mods[name] = m
    return mods["MEAS"], mods["VIT"]


def cli_surface(mod):
    """{flag: {type, default, required, nargs}} read from the `parser.add_argument(...)` calls of the script's AST."""
    import ast
    flags = {}
    for node in ast.walk(ast.parse(open(mod.__file__).read())):
        if isinstance(node, ast.Call) and getattr(node.func, "attr", "") == "add_argument" and node.args:
            name = ast.literal_eval(node.args[0])
            kw = {k.arg: k.value for k in node.keywords}
            flags[name] = {"type": getattr(kw.get("type"), "id", None),
                           "default": ast.literal_eval(kw["default"]) if "default" in kw else None,
                           "required": bool(ast.literal_eval(kw["required"])) if "required" in kw else False,
                           "nargs": ast.literal_eval(kw["nargs"]) if "nargs" in kw else None}
    return flags


class _Labelled(torch.utils.data.Dataset):
    def __init__(self, labels):
        self.labels = labels

    def __len__(self):
        return len(self.labels)

    def __getitem__(self, i):
        return i, self.labels[i]


def main():
    MEAS, VIT = load_reference_scripts()
    out = {}
    # ---- label perturbations
    labels = [(7 * i + 3) % 10 for i in range(57)]
    base = _Labelled(labels)
    out["labels"] = labels
    for seed in (42, 7):
        sh = MEAS.ShuffledLabelsDataset(base, shuffle_seed=seed)
        tn = MEAS.TargetNoiseDataset(base, num_classes=10, noise_seed=seed)
        out[f"label_shuffle_seed{seed}"] = [int(sh[i][1]) for i in range(len(sh))]
        out[f"target_noise_seed{seed}"] = [int(tn[i][1]) for i in range(len(tn))]
        assert all(sh[i][0] == i and tn[i][0] == i for i in range(len(sh)))
    tn1000 = MEAS.TargetNoiseDataset(base, num_classes=1000, noise_seed=42)
    out["target_noise_1000_seed42"] = [int(tn1000[i][1]) for i in range(len(tn1000))]
    # ---- image transforms
    t = torch.arange(12.0).reshape(3, 2, 2)
    g = MEAS.GaussianNoiseTransform(lambda im: im, epsilon=0.25)
    torch.manual_seed(5)
    out["gaussian_eps0.25_seed5"] = g(t).flatten().tolist()
    out["uniform_gray_sum"] = float(MEAS.UniformGrayTransform(lambda im: im)(t).abs().sum())
    # ---- schedule (both scripts carry the same class)
    for tag, mod in (("meas", MEAS), ("vit", VIT)):
        opt = types.SimpleNamespace(param_groups=[{"lr": 0.1}])
        s = mod.CosineAnnealingLRWithWarmup(opt, warmup_epochs=5, max_epochs=100, eta_min=0)
        lrs = []
        for _ in range(100):
            s.step()
            lrs.append(opt.param_groups[0]["lr"])
        out[f"schedule_{tag}"] = lrs
        out[f"schedule_{tag}_state"] = s.state_dict()
    # ---- save_checkpoint on CPU (VIT:89-123)
    lin = torch.nn.Linear(3, 2)
    wrapped = types.SimpleNamespace(module=lin)
    opt = torch.optim.SGD(lin.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
    lin(torch.ones(1, 3)).sum().backward()
    opt.step()
    sched = VIT.CosineAnnealingLRWithWarmup(opt, 5, 100)
    scaler = VIT.GradScaler(enabled=False) if not torch.cuda.is_available() else VIT.GradScaler()
    with tempfile.TemporaryDirectory() as d:
        VIT.save_checkpoint(3, wrapped, opt, sched, scaler, 1.23456789, 2.5, 12.3456789, d, 0)
        VIT.save_checkpoint(4, wrapped, opt, sched, scaler, 1.0, 2.0, 50.0, d, 0)
        ck = torch.load(os.path.join(d, "checkpoint_epoch_003.pth"), weights_only=False)
        out["checkpoint_keys"] = sorted(ck.keys())
        out["checkpoint_files"] = sorted(os.listdir(d))
        out["metrics_csv"] = open(os.path.join(d, "training_metrics.csv")).read()
        out["optimizer_group_keys"] = sorted(ck["optimizer_state_dict"]["param_groups"][0].keys())
        out["optimizer_state_keys"] = sorted(ck["optimizer_state_dict"]["state"][0].keys())
    out["fresh_grad_scaler_state"] = torch.amp.GradScaler("cpu").state_dict()
    # ---- shipped measurement rows
    eff = pd.read_csv(os.path.join(REF, "Data/vit_results/perturbation_effects.csv"))
    out["effects_columns"] = list(eff.columns)
    out["effects_rows"] = eff.to_dict(orient="records")
    out["effects_order"] = [[int(r.perturb_epoch), r.perturbation_type] for r in eff.itertuples()]
    # ---- command-line defaults, read from the script's argparse source (MEAS:576-581)
    src = open(MEAS.__file__).read()
    out["default_perturb_epochs"] = json.loads(re.search(r"'--perturb_epochs'.*?default=(\[[^\]]*\])", src, re.S).group(1))
    out["default_perturbation_types"] = json.loads(
        re.search(r"'--perturbation_types'.*?default=(\[[^\]]*\])", src, re.S).group(1).replace("'", '"'))
    # ---- the two scripts' command lines: every flag with its type, default and required-ness (VIT:247-257, MEAS:562-599)
    out["cli"] = {"VIT": cli_surface(VIT), "MEAS": cli_surface(MEAS)}
    path = os.path.join(ROOT, "tests", "golden", "vit_measure.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

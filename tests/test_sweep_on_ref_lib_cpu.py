"""CPU dry runs of the two GPU sweep tests (tests/test_gpu_sweep.py, tests/test_gpu_zz_sweep_chain.py): the same
scenarios through `hba.sweep.run_sweep` and the unmodified `run_behavioral_training`, with libhba served by its CPU
restatement inside every worker process (oracle/libhba_ref.py).  What a worker keeps from condition to condition -
the frozen CLIP and its staged operands, the resident image store, the frozen-trunk cache, the text-trunk cache - and
LEN's shorter -> longer resume chain are host logic, so their "a condition's rows do not depend on what ran before it"
property is checkable without a GPU (the GPU tests assert the same on the device)."""
import csv
import os

import pytest
import torch

from test_gpu_pipeline import _write_things_like_dataset

pytestmark = pytest.mark.timeout(1200)


def _cpu_condition(cfg):
    """Runs in the spawned worker: libhba's CPU restatement for the whole worker life, the seeded tiny checkpoint,
    then the unmodified pipeline entry point (bf16 mode, like the GPU tests)."""
    import hba
    import functions.new_cvpr_train_behavior_things_pipeline as NEW
    from oracle import libhba_ref
    from src.models.CLIPs.clip_hba import clip
    if not getattr(libhba_ref, "_worker_ctx", None):
        libhba_ref._worker_ctx = libhba_ref.emulated_device()
        libhba_ref._worker_ctx.__enter__()          # (never left: the process ends with the sweep)
        hba.set_precision("bf16")
        NEW.select_device = lambda flag: torch.device("cpu")
    ckpt = cfg["_tiny_ckpt"]
    clip._download = lambda url, root: ckpt
    torch.set_num_threads(2)
    NEW.run_behavioral_training({k: v for k, v in cfg.items() if not k.startswith("_")})


def _setup(tmp_path, monkeypatch, base_epochs):
    import hba
    import functions.cvpr_train_behavior_things_pipeline_baseline as BASE
    from oracle import clip_ref
    from oracle.libhba_ref import emulated_device
    from src.models.CLIPs.clip_hba import clip
    root = str(tmp_path)
    ckpt = os.path.join(root, "ViT-tiny-14.pt")
    torch.save(clip_ref.synthetic_state_dict("ViT-tiny/14", seed=1), ckpt)
    monkeypatch.setattr(clip, "_download", lambda url, r: ckpt)
    monkeypatch.setenv("HBA_SYNTHETIC_OK", "1")
    monkeypatch.setenv("HBA_REUSE_MODEL", "0")       # (the parent's own baseline run: nothing to share with other tests)
    img_dir = _write_things_like_dataset(root, n_train=10)
    common = {"csv_file": f"{root}/train.csv", "img_dir": img_dir, "inference_csv_file": f"{root}/rsa.csv",
              "RDM48_triplet_dir": f"{root}/RDM48_triplet.mat", "backbone": "ViT-tiny/14", "batch_size": 4,
              "lr": 3e-4, "random_seed": 1, "vision_layers": 2, "transformer_layers": 1, "rank": 8,
              "criterion": torch.nn.MSELoss(), "cuda": 0}
    base_cfg = dict(common, epochs=base_epochs, train_portion=0.8, early_stopping_patience=20, logger=None,
                    checkpoint_path=f"{root}/base/model.pth", training_res_path=f"{root}/base/res.csv",
                    dora_parameters_path=f"{root}/base/dora", random_state_path=f"{root}/base/rand")
    hba.set_precision("bf16")
    monkeypatch.setattr(BASE, "select_device", lambda flag: torch.device("cpu"))
    with emulated_device():
        BASE.run_behavioral_training(base_cfg)
    monkeypatch.setenv("HBA_REUSE_MODEL", "1")       # the workers reuse their frozen CLIP: that is what is under test
    sweep_cfg = dict(common, epochs=4, early_stopping_patience=20, hba_resident=True, logger=None,
                     baseline_dora_directory=f"{root}/base/dora", baseline_random_state_path=f"{root}/base/rand",
                     baseline_split_indices_path=f"{root}/base/rand/dataset_split_indices.pth",
                     perturb_type="random_target", perturb_length=1, perturb_distribution="target",
                     perturb_seed=42, previous_training_res_path=f"{root}/base/res.csv", _tiny_ckpt=ckpt)
    return root, sweep_cfg


def test_conditions_back_to_back_in_one_worker_equal_a_fresh_worker(tmp_path, monkeypatch):
    from hba import sweep
    root, sweep_cfg = _setup(tmp_path, monkeypatch, base_epochs=3)
    conds = [{"training_run": e, "perturb_length": 1} for e in (1, 2, 3)]
    logs = []
    many = sweep.run_sweep(dict(sweep_cfg, output_base_directory=f"{root}/many"), conds, [None], run_fn=_cpu_condition,
                           log=logs.append)
    assert [r["ok"] for r in many] == [True, True, True], [r["error"] for r in many]
    alone = sweep.run_sweep(dict(sweep_cfg, output_base_directory=f"{root}/alone"), [conds[2], conds[0]], [None],
                            run_fn=_cpu_condition, log=logs.append)
    assert all(r["ok"] for r in alone), [r["error"] for r in alone]
    for run in (3, 1):        # run 3 resumes from baseline epoch 2; run 1 starts from scratch (constructor-draw replay)
        a = list(csv.reader(open(f"{root}/many/training_run{run}/training_res_run{run}.csv")))
        b = list(csv.reader(open(f"{root}/alone/training_run{run}/training_res_run{run}.csv")))
        assert len(a) == 5 and a == b, run              # header + epochs 1..4, identical whatever ran before
    a = list(csv.reader(open(f"{root}/many/training_run3/training_res_run3.csv")))
    assert [r[5] for r in a[3:]] == ["True", "False"]                     # window = epoch 3
    base = list(csv.reader(open(f"{root}/base/res.csv")))
    assert [r[:5] for r in a[1:3]] == [r[:5] for r in base[1:3]]          # epochs 1, 2 inherited from the baseline
    # a from-scratch condition starts from the baseline's own initial DoRA values (same seed, same draws): its first,
    # perturbed epoch differs from the baseline's, but it is a finite, trained trajectory
    one = list(csv.reader(open(f"{root}/many/training_run1/training_res_run1.csv")))
    assert one[1][5] == "True" and all(float(r[1]) == float(r[1]) for r in one[1:])
    assert os.path.exists(f"{root}/many/training_run1/dora_params_run1/epoch4_dora_params.pth")


def test_chained_length_conditions_equal_independent_ones_on_the_cpu(tmp_path, monkeypatch):
    from hba import sweep
    root, sweep_cfg = _setup(tmp_path, monkeypatch, base_epochs=2)
    conds = [{"training_run": 2, "perturb_length": 2}, {"training_run": 2, "perturb_length": 3}]
    logs = []
    chained = sweep.run_sweep(dict(sweep_cfg, epochs=5, output_base_directory=f"{root}/chain"), conds, [None],
                              layout="length", run_fn=_cpu_condition, log=logs.append, chain=True)
    assert [r["ok"] for r in chained] == [True, True], [r["error"] for r in chained]
    alone = sweep.run_sweep(dict(sweep_cfg, epochs=5, output_base_directory=f"{root}/alone"), conds[1:], [None],
                            layout="length", run_fn=_cpu_condition, log=logs.append)
    assert alone[0]["ok"], alone[0]["error"]
    a = list(csv.reader(open(f"{root}/chain/random_target_e2_l3/training_res.csv")))
    b = list(csv.reader(open(f"{root}/alone/random_target_e2_l3/training_res.csv")))
    assert len(a) == len(b) == 6 and [r[0] for r in a[1:]] == ["1", "2", "3", "4", "5"]
    assert [r[5] for r in a[2:]] == ["True", "True", "True", "False"]   # window = epochs 2..4
    assert a == b                                     # resumed after epoch 3 of the window-2 run == trained alone
    own = sorted(os.listdir(f"{root}/chain/random_target_e2_l3/dora_params_2"))
    assert own[0] == "epoch4_dora_params.pth" and "epoch3_dora_params.pth" not in own

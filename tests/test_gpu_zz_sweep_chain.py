"""GPU test of LEN's shorter -> longer resume chain (hba.sweep.run_sweep(chain=True); reference LEN =
Training/clip_behavioral_finetuning/length_experiments/clip_train_behavior_lengths.py:188-253) through the real
pipeline: the window-3 condition resumed from the end of the finished window-2 run of the same start epoch
must write the same result rows as the same condition trained from the baseline checkpoint on its own."""
import csv
import os

import pytest
import torch

from test_gpu_pipeline import _write_things_like_dataset
from test_gpu_sweep import _gpu_condition

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


def test_chained_length_conditions_equal_independent_ones(tmp_path, monkeypatch):
    import hba
    import functions.cvpr_train_behavior_things_pipeline_baseline as BASE
    from hba import sweep
    from oracle import clip_ref
    from src.models.CLIPs.clip_hba import clip
    hba.set_precision("bf16")
    root = str(tmp_path)
    ckpt = os.path.join(root, "ViT-tiny-14.pt")
    torch.save(clip_ref.synthetic_state_dict("ViT-tiny/14", seed=1), ckpt)
    monkeypatch.setattr(clip, "_download", lambda url, r: ckpt)
    img_dir = _write_things_like_dataset(root, n_train=22)
    common = {"csv_file": f"{root}/train.csv", "img_dir": img_dir, "inference_csv_file": f"{root}/rsa.csv",
              "RDM48_triplet_dir": f"{root}/RDM48_triplet.mat", "backbone": "ViT-tiny/14", "batch_size": 4,
              "lr": 3e-4, "random_seed": 1, "vision_layers": 2, "transformer_layers": 1, "rank": 8,
              "criterion": torch.nn.MSELoss(), "cuda": 0}
    base_cfg = dict(common, epochs=2, train_portion=0.8, early_stopping_patience=20, logger=None,
                    checkpoint_path=f"{root}/base/model.pth", training_res_path=f"{root}/base/res.csv",
                    dora_parameters_path=f"{root}/base/dora", random_state_path=f"{root}/base/rand")
    BASE.run_behavioral_training(base_cfg)          # baseline epochs 1..2: start epoch 2 resumes from epoch 1
    sweep_cfg = dict(common, epochs=6, early_stopping_patience=20, hba_resident=True, logger=None,
                     baseline_dora_directory=f"{root}/base/dora", baseline_random_state_path=f"{root}/base/rand",
                     baseline_split_indices_path=f"{root}/base/rand/dataset_split_indices.pth",
                     perturb_type="random_target", perturb_length=1, perturb_distribution="target",
                     perturb_seed=42, previous_training_res_path=f"{root}/base/res.csv", _tiny_ckpt=ckpt)
    conds = [{"training_run": 2, "perturb_length": 2}, {"training_run": 2, "perturb_length": 3}]
    logs = []
    chained = sweep.run_sweep(dict(sweep_cfg, output_base_directory=f"{root}/chain"), conds, [0], layout="length",
                              run_fn=_gpu_condition, log=logs.append, chain=True)
    assert [r["ok"] for r in chained] == [True, True], [r["error"] for r in chained]
    alone = sweep.run_sweep(dict(sweep_cfg, output_base_directory=f"{root}/alone"), conds[1:], [0], layout="length",
                            run_fn=_gpu_condition, log=logs.append)
    assert alone[0]["ok"], alone[0]["error"]
    a = list(csv.reader(open(f"{root}/chain/random_target_e2_l3/training_res.csv")))
    b = list(csv.reader(open(f"{root}/alone/random_target_e2_l3/training_res.csv")))
    assert len(a) == len(b) == 7                      # header + epochs 1..6
    assert [r[0] for r in a[1:]] == ["1", "2", "3", "4", "5", "6"]
    assert len(a[1]) == 5                             # epoch 1 is the baseline run's row (no perturbation flags)
    assert [r[5] for r in a[2:]] == ["True", "True", "True", "False", "False"]   # window = epochs 2..4
    assert a == b                                     # resumed after epoch 3 of the window-2 run == trained alone
    # the chained run did not recompute the epochs it inherited: its own checkpoints start after the resume epoch
    own = sorted(os.listdir(f"{root}/chain/random_target_e2_l3/dora_params_2"))
    assert own[0] == "epoch4_dora_params.pth" and "epoch3_dora_params.pth" not in own
    assert os.path.exists(f"{root}/alone/random_target_e2_l3/dora_params_2/epoch2_dora_params.pth")

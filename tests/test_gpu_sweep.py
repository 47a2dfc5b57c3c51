"""GPU test of the multi-condition sweep driver (hba.sweep) through the real pipeline: worker processes
pinned with CUDA_VISIBLE_DEVICES run `run_behavioral_training` for several conditions back to back (frozen
CLIP, staged weights, resident store and frozen-trunk cache are reused inside a worker), and a condition's
results do not depend on what ran before it in the same worker."""
import csv
import os

import pytest
import torch

from test_gpu_pipeline import _write_things_like_dataset

pytestmark = pytest.mark.gpu


def _gpu_condition(cfg):
    """Runs in the spawned worker: serve the seeded ViT-tiny weights (what the tests' `tiny_checkpoint`
    fixture does in-process), then the unmodified pipeline entry point."""
    from src.models.CLIPs.clip_hba import clip
    import functions.new_cvpr_train_behavior_things_pipeline as NEW
    ckpt = cfg["_tiny_ckpt"]
    clip._download = lambda url, root: ckpt
    assert torch.cuda.device_count() == 1          # pinned: only the assigned GPU is visible
    NEW.run_behavioral_training({k: v for k, v in cfg.items() if not k.startswith("_")})


def test_sweep_workers_run_conditions_back_to_back_without_cross_talk(tmp_path, monkeypatch):
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
    base_cfg = dict(common, epochs=3, train_portion=0.8, early_stopping_patience=20, logger=None,
                    checkpoint_path=f"{root}/base/model.pth", training_res_path=f"{root}/base/res.csv",
                    dora_parameters_path=f"{root}/base/dora", random_state_path=f"{root}/base/rand")
    BASE.run_behavioral_training(base_cfg)          # baseline epochs 1..3: the checkpoints conditions resume from
    sweep_cfg = dict(common, epochs=6, early_stopping_patience=20, hba_resident=True, logger=None,
                     baseline_dora_directory=f"{root}/base/dora", baseline_random_state_path=f"{root}/base/rand",
                     baseline_split_indices_path=f"{root}/base/rand/dataset_split_indices.pth",
                     perturb_type="random_target", perturb_length=1, perturb_distribution="target",
                     perturb_seed=42, previous_training_res_path=f"{root}/base/res.csv", _tiny_ckpt=ckpt)
    conds = [{"training_run": e, "perturb_length": 1} for e in (1, 2, 3)]
    logs = []
    # one worker, three conditions back to back (run 1 first: it is the longest under the LPT cost model)
    many = sweep.run_sweep(dict(sweep_cfg, output_base_directory=f"{root}/many"), conds, [0], run_fn=_gpu_condition,
                           log=logs.append)
    assert [r["ok"] for r in many] == [True, True, True], [r["error"] for r in many]
    # the last condition alone, in a fresh worker process
    alone = sweep.run_sweep(dict(sweep_cfg, output_base_directory=f"{root}/alone"), conds[2:], [0],
                            run_fn=_gpu_condition, log=logs.append)
    assert alone[0]["ok"], alone[0]["error"]
    a = list(csv.reader(open(f"{root}/many/training_run3/training_res_run3.csv")))
    b = list(csv.reader(open(f"{root}/alone/training_run3/training_res_run3.csv")))
    assert len(a) == 7 and a == b                   # header + epochs 1..6, identical
    assert [r[5] for r in a[3:]] == ["True", "False", "False", "False"]   # epochs 3..6: window = epoch 3
    assert os.path.exists(f"{root}/many/training_run1/dora_params_run1/epoch6_dora_params.pth")
    assert "3 successful, 0 failed" in logs[-2] or any("3 successful, 0 failed" in l for l in logs)

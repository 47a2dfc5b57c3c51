"""Analysis layer of the perturbation study: the result CSVs a sweep writes -> the tables
the reference's figures plot, as plain functions and a CSV-in / CSV-out CLI (tools/analyze_sweep.py).

Reference: the analysis cells of Figures/fig3 (Single Sweep Perturbation Experiments)/fig3.ipynb (cells 4-10:
deviation of test loss / behavioural alignment at the perturbed epoch from the baseline run) and
Figures/fig4 (Perturbation Recovery)/fig4.ipynb (cells 4-12: epochs until the test loss is back within 1 % of
the baseline).  Host-side pandas only - there is no GPU work here; the directory layouts are the ones
hba.sweep.condition_config writes (SWEEP:198-207, LEN:128-137)."""
from __future__ import annotations

import os
import re

import numpy as np
import pandas as pd

RECOVERY_TOLERANCE = 1.01   # "within 1 % of baseline", FIG4 cell 12


def load_baseline(csv_path):
    """Baseline run trimmed at its minimum test loss = the early-stopping point (FIG3 cell 4, FIG4 cell 4)."""
    df = pd.read_csv(csv_path)
    return df.loc[:df["test_loss"].idxmin()].copy()


def discover_single_sweep(sweep_root):
    """{run number: DataFrame} from `training_run{e}/training_res_run{e}.csv` (FIG3 cell 6, SWEEP:198-207)."""
    runs = {}
    for name in sorted(os.listdir(sweep_root)):
        m = re.fullmatch(r"training_run(\d+)", name)
        path = os.path.join(sweep_root, name, f"training_res_run{m.group(1)}.csv") if m else None
        if path and os.path.exists(path):
            runs[int(m.group(1))] = pd.read_csv(path)
    return runs


def discover_length_runs(base_dir, prefix="random_target"):
    """[(run name, start epoch, window length, DataFrame)] from `{prefix}_e{E}_l{L}/training_res.csv` (or
    `metrics.csv`), FIG4 cells 6-10 / LEN:128-137.  Frames are left untrimmed, as in FIG4 cell 10."""
    out = []
    for name in sorted(os.listdir(base_dir)):
        m = re.fullmatch(re.escape(prefix) + r"_e(\d+)_l(\d+)", name)
        if not m or not os.path.isdir(os.path.join(base_dir, name)):
            continue
        for fname in ("metrics.csv", "training_res.csv"):
            path = os.path.join(base_dir, name, fname)
            if os.path.exists(path):
                out.append((name, int(m.group(1)), int(m.group(2)), pd.read_csv(path)))
                break
    return out


def deviation_at_perturbation_epoch(baseline_df, runs, column="test_loss"):
    """FIG3 cells 8 / 10: run value at its perturbed epoch minus the baseline value at that epoch.
    -> DataFrame[run, delta] sorted by run; runs lacking the epoch on either side are dropped."""
    base = baseline_df.drop_duplicates("epoch").set_index("epoch")[column]
    rows = []
    for run, df in runs.items():
        e = int(run)
        at = df.loc[df["epoch"] == e, column]
        if len(at) and e in base.index:
            rows.append((e, float(at.iloc[0]) - float(base.loc[e])))
    return pd.DataFrame(sorted(rows), columns=["run", "delta_" + column])


def recovery_table(baseline_df, runs, tolerance=RECOVERY_TOLERANCE):
    """FIG4 cell 12 for every (name, start, length, frame): first epoch after the window whose test loss is
    <= tolerance x the baseline test loss of the same epoch (baseline minimum beyond the baseline's last
    epoch).  -> DataFrame[run_name, start_epoch, length, perturbation_end, recovery_epoch, epochs_to_recovery,
    recovered] sorted by (start_epoch, length); unrecovered runs carry NaN."""
    base = baseline_df.drop_duplicates("epoch").set_index("epoch")["test_loss"]
    floor = float(baseline_df["test_loss"].min())
    rows = []
    for name, start, length, df in runs:
        end = start + length - 1
        d = df.sort_values("epoch", kind="stable")
        epochs = d["epoch"].to_numpy().astype(np.int64)
        target = base.reindex(epochs).to_numpy(dtype=np.float64)
        target = np.where(np.isnan(target), floor, target) * tolerance
        ok = (epochs > end) & (d["test_loss"].to_numpy(dtype=np.float64) <= target)
        rec = int(epochs[np.argmax(ok)]) if ok.any() else None
        rows.append({"run_name": name, "start_epoch": start, "length": length, "perturbation_end": end,
                     "recovery_epoch": rec, "epochs_to_recovery": None if rec is None else rec - end,
                     "recovered": rec is not None})
    cols = ["run_name", "start_epoch", "length", "perturbation_end", "recovery_epoch", "epochs_to_recovery", "recovered"]
    out = pd.DataFrame(rows, columns=cols)
    return out.sort_values(["start_epoch", "length"], kind="stable").reset_index(drop=True)


def single_sweep_summary(baseline_csv, sweep_root):
    """One row per single-epoch condition: delta test loss and delta behavioural alignment (FIG3 b/c)."""
    base = load_baseline(baseline_csv)
    runs = discover_single_sweep(sweep_root)
    a = deviation_at_perturbation_epoch(base, runs, "test_loss")
    b = deviation_at_perturbation_epoch(base, runs, "behavioral_rsa_rho")
    return a.merge(b, on="run", how="outer").sort_values("run").reset_index(drop=True)


def length_grid_summary(baseline_csv, base_dir, prefix="random_target"):
    """One row per (start, length) condition of the variable-length grid: the recovery table of FIG4."""
    return recovery_table(load_baseline(baseline_csv), discover_length_runs(base_dir, prefix))


# ------------------------------------------------------------------------------------------------
# FIG2 = Figures/fig2 (Effects of Different Perturbations)/fig2.ipynb: the four perturbation types side by side
# (CLIP-HBA, cells 3-9) and the ViT measurement table (cells 12-14, Data/vit_results)
# ------------------------------------------------------------------------------------------------
FIG2_TARGET_EPOCHS = (5, 15, 25, 35, 45, 70, 98)          # FIG2 cell 7
FIG2_TYPE_DIRS = {"image_noise": "image_noise", "blank_image": "uniform_target", "label_shuffle": "label_shuffle",
                  "target_noise": "target_noise"}         # FIG2 cell 7: plotted name -> directory of the runs
VIT_SUMMARY_COLUMNS = ["perturb_epoch", "perturbation_type", "delta_loss", "delta_rsa", "baseline_loss", "baseline_rsa"]


def discover_flat_runs(root):
    """{run number: DataFrame} from the flat `training_res_run{e}.csv` files of one perturbation type
    (FIG2 cell 5 `load_all_run_data`)."""
    runs = {}
    if not os.path.isdir(root):
        return runs
    for name in sorted(os.listdir(root)):
        m = re.fullmatch(r"training_res_run(\d+)\.csv", name)
        if m:
            runs[int(m.group(1))] = pd.read_csv(os.path.join(root, name))
    return runs


def perturbation_type_comparison(baseline_df, runs_by_type, target_epochs=FIG2_TARGET_EPOCHS):
    """FIG2 cells 5-9 as one long table: for every perturbation type and target epoch e, the run
    `training_res_run{e}`'s test loss / behavioural alignment AT epoch e minus the (trimmed) baseline's; NaN
    where the run, its row, or the baseline row is missing.  Columns: perturbation, epoch, delta_test_loss,
    delta_behavioral_rsa_rho."""
    base = baseline_df.drop_duplicates("epoch").set_index("epoch")
    rows = []
    for name, runs in runs_by_type.items():
        for e in target_epochs:
            d_loss = d_ba = np.nan
            df = runs.get(e)
            if df is not None and "epoch" in df.columns and e in base.index:
                hit = df[df["epoch"] == e]
                if len(hit) > 0:
                    d_loss = float(hit.iloc[0]["test_loss"]) - float(base.at[e, "test_loss"])
                    d_ba = float(hit.iloc[0]["behavioral_rsa_rho"]) - float(base.at[e, "behavioral_rsa_rho"])
            rows.append((name, e, d_loss, d_ba))
    return pd.DataFrame(rows, columns=["perturbation", "epoch", "delta_test_loss", "delta_behavioral_rsa_rho"])


def perturbation_type_summary(baseline_csv, results_root, type_dirs=None, target_epochs=FIG2_TARGET_EPOCHS):
    """The FIG2 (A)/(B) table from a results tree laid out like the reference's Data/clip_results."""
    type_dirs = type_dirs or FIG2_TYPE_DIRS
    runs = {name: discover_flat_runs(os.path.join(results_root, d)) for name, d in type_dirs.items()}
    return perturbation_type_comparison(load_baseline(baseline_csv), runs, target_epochs)


def vit_perturbation_summary(effects):
    """Data/vit_results/perturbation_summary_table.csv from the rows `hba.vit_train.measure_all` (reference
    MEAS:653-657) writes: the six summary columns, ordered by (perturb_epoch, perturbation_type), rounded to
    4 decimals.  `effects`: CSV path or DataFrame."""
    df = pd.read_csv(effects) if isinstance(effects, (str, os.PathLike)) else effects
    out = df.sort_values(["perturb_epoch", "perturbation_type"], kind="stable")[VIT_SUMMARY_COLUMNS]
    return out.round(4).reset_index(drop=True)

"""TEST INFRASTRUCTURE ONLY (never imported by the product path): row-by-row restatement of the analysis
cells of the reference's figure notebooks, which turn the result CSVs of the perturbation sweeps into the
numbers its figures plot.

  FIG3 = Figures/fig3 (Single Sweep Perturbation Experiments)/fig3.ipynb
  FIG4 = Figures/fig4 (Perturbation Recovery)/fig4.ipynb

Pinned against the reference's own shipped result CSVs (Data/clip_results/...) by oracle/make_analysis_golden.py
-> tests/golden/analysis.json."""
import pandas as pd


def trim_at_min_test_loss(df):
    """FIG3 cell 4 / FIG4 cell 4: keep the epochs up to the minimum test loss (the early-stopping point)."""
    return df.loc[:df["test_loss"].idxmin()].copy()


def deviation_at_perturbation_epoch(baseline_df, runs, column):
    """FIG3 cells 8 (column = 'test_loss') and 10 ('behavioral_rsa_rho'): for run number e, the run's value
    at epoch e minus the (trimmed) baseline's value at epoch e; runs without that epoch on either side are
    dropped; sorted by run number.  runs: {run_number: DataFrame}."""
    out = []
    for run_num, df in runs.items():
        perturb_epoch = int(run_num)
        run_at = df[df["epoch"] == perturb_epoch]
        base_at = baseline_df[baseline_df["epoch"] == perturb_epoch]
        if len(run_at) > 0 and len(base_at) > 0:
            out.append((perturb_epoch, run_at.iloc[0][column] - base_at.iloc[0][column]))
    return sorted(out)


def recovery_table(baseline_df, runs):
    """FIG4 cell 12.  runs: list of (run_name, start_epoch, length, UNTRIMMED DataFrame) (FIG4 cell 10 keeps the
    raw frames).  Recovery epoch = first epoch after the window's last epoch (start + length - 1) whose test
    loss is <= 1.01 x the trimmed baseline's test loss at the same epoch, or <= 1.01 x the baseline minimum when
    the baseline has no such epoch."""
    rows = []
    for run_name, start_epoch, length, df in runs:
        perturbation_end = start_epoch + length - 1
        recovery_epoch = None
        for _, row in df.sort_values("epoch").iterrows():
            current_epoch = int(row["epoch"])
            if current_epoch <= perturbation_end:
                continue
            base_at = baseline_df[baseline_df["epoch"] == current_epoch]
            if len(base_at) > 0:
                target = base_at.iloc[0]["test_loss"] * 1.01
            else:
                target = baseline_df["test_loss"].min() * 1.01
            if row["test_loss"] <= target:
                recovery_epoch = current_epoch
                break
        rows.append({"run_name": run_name, "start_epoch": start_epoch, "length": length,
                     "perturbation_end": perturbation_end, "recovery_epoch": recovery_epoch,
                     "epochs_to_recovery": None if recovery_epoch is None else recovery_epoch - perturbation_end,
                     "recovered": recovery_epoch is not None})
    return pd.DataFrame(rows).sort_values(["start_epoch", "length"]).reset_index(drop=True)


# ------------------------------------------------------------------------------------------------
# FIG2 = Figures/fig2 (Effects of Different Perturbations)/fig2.ipynb
# ------------------------------------------------------------------------------------------------
FIG2_TARGET_EPOCHS = [5, 15, 25, 35, 45, 70, 98]          # FIG2 cell 7


def fig2_type_deviations(baseline_df, runs_by_type, target_epochs=FIG2_TARGET_EPOCHS):
    """FIG2 cells 5-7, row by row.  `runs_by_type`: {perturbation name: {run number: DataFrame}} (each type's
    directory holds flat `training_res_run{e}.csv` files); `baseline_df` trimmed at its minimum test loss
    (cell 3).  -> {name: {"test_loss": [...], "behavioral_rsa_rho": [...]}} with NaN where the run, the run's
    row at epoch e, or the baseline row at epoch e is missing."""
    import math
    out = {}
    for name, runs in runs_by_type.items():
        d_loss, d_ba = [], []
        for ep in target_epochs:
            df = runs.get(ep)
            base_row = baseline_df[baseline_df["epoch"] == ep]
            # cell 7 compute_deltas -> cell 5 load_run_epoch_value / baseline_epoch_loss
            run_loss = None
            if df is not None and "epoch" in df.columns:
                row = df[df["epoch"] == ep]
                if len(row) > 0:
                    run_loss = float(row.iloc[0]["test_loss"])
            base_loss = float(base_row["test_loss"].iloc[0]) if len(base_row) > 0 else None
            d_loss.append(math.nan if run_loss is None or base_loss is None else run_loss - base_loss)
            # cell 5 compute_ba_from_dict
            if df is None:
                d_ba.append(math.nan)
                continue
            row = df[df["epoch"] == ep]
            if len(row) == 0 or len(base_row) == 0:
                d_ba.append(math.nan)
                continue
            d_ba.append(float(row.iloc[0]["behavioral_rsa_rho"]) - float(base_row.iloc[0]["behavioral_rsa_rho"]))
        out[name] = {"test_loss": d_loss, "behavioral_rsa_rho": d_ba}
    return out


def vit_summary_table(effects_df):
    """Data/vit_results/perturbation_summary_table.csv from perturbation_effects.csv: the six summary columns,
    rows ordered by (perturb_epoch, perturbation_type), values rounded to 4 decimals (verified cell by cell
    against the shipped table by oracle/make_analysis_golden.py)."""
    cols = ["perturb_epoch", "perturbation_type", "delta_loss", "delta_rsa", "baseline_loss", "baseline_rsa"]
    rows = sorted(effects_df[cols].to_dict("records"), key=lambda r: (r["perturb_epoch"], r["perturbation_type"]))
    for r in rows:
        for c in cols[2:]:
            r[c] = round(float(r[c]), 4)
    return rows

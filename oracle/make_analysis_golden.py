"""Generates tests/golden/analysis.json from the REFERENCE'S OWN shipped result CSVs
(/root/reference/Data/clip_results) with the row-by-row restatement of its notebook cells
(oracle/analysis_ref.py).  The fixture carries a compact excerpt of the inputs (baseline test loss / alignment
series and six length-grid runs + eight single-epoch runs, epochs and the two metric columns only) together
with the expected outputs for those inputs, and the full-size summary (136 + 98 conditions) for checks run
where /root/reference is mounted.

    python oracle/make_analysis_golden.py
"""
import json
import os
import sys

import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import analysis_ref as ref  # noqa: E402

DATA = "/root/reference/Data/clip_results"
LEN_DIR = os.path.join(DATA, "perturb_length_experiments_baselineseed1_perturbseed0")
SWEEP_DIR = os.path.join(DATA, "single_sweep_experiments")
COLS = ["epoch", "test_loss", "behavioral_rsa_rho"]


FIG2_DIRS = {"image_noise": "image_noise", "blank_image": "uniform_target", "label_shuffle": "label_shuffle",
             "target_noise": "target_noise"}          # FIG2 cell 7: name -> directory under Data/clip_results


def fig2_golden(base):
    """FIG2: the four perturbation-type directories (flat training_res_run{e}.csv files) against the trimmed
    baseline, and the ViT summary table against the shipped Data/vit_results/perturbation_summary_table.csv."""
    runs_by_type, excerpt = {}, {}
    for name, d in FIG2_DIRS.items():
        runs_by_type[name] = {}
        for f in sorted(os.listdir(os.path.join(DATA, d))):
            if f.startswith("training_res_run") and f.endswith(".csv"):
                e = int(f[len("training_res_run"):-4])
                df = pd.read_csv(os.path.join(DATA, d, f))
                runs_by_type[name][e] = df
        # excerpt: per run, the rows around its perturbed epoch are all the cells read
        excerpt[name] = {str(e): [[int(r.epoch), float(r.test_loss), float(r.behavioral_rsa_rho)]
                                  for r in df[(df["epoch"] >= e - 1) & (df["epoch"] <= e + 1)].itertuples()]
                         for e, df in runs_by_type[name].items()}
    dev = ref.fig2_type_deviations(base, runs_by_type)
    vit_dir = "/root/reference/Data/vit_results"
    eff = pd.read_csv(os.path.join(vit_dir, "perturbation_effects.csv"))
    shipped = pd.read_csv(os.path.join(vit_dir, "perturbation_summary_table.csv"))
    table = ref.vit_summary_table(eff)
    assert list(shipped.columns) == list(table[0].keys()) and len(shipped) == len(table)
    for want, got in zip(shipped.to_dict("records"), table):      # the restatement reproduces the shipped table
        assert want == got, (want, got)
    nan_to_none = lambda xs: [None if v != v else float(v) for v in xs]
    return {"fig2_target_epochs": ref.FIG2_TARGET_EPOCHS,
            "fig2_runs": excerpt,
            "fig2_expected": {n: {k: nan_to_none(v) for k, v in d.items()} for n, d in dev.items()},
            "vit_effects_rows": eff.to_dict("records"),
            "vit_summary_csv": open(os.path.join(vit_dir, "perturbation_summary_table.csv")).read()}


def main():
    base_raw = pd.read_csv(os.path.join(DATA, "baseline_clip_results_seed1.csv"))
    base = ref.trim_at_min_test_loss(base_raw)
    # ---- length grid (FIG4)
    runs = []
    for name in sorted(os.listdir(LEN_DIR)):
        p = os.path.join(LEN_DIR, name, "metrics.csv")      # FIG4 cell 8: metrics.csv first, then training_res.csv
        if not os.path.exists(p):
            p = os.path.join(LEN_DIR, name, "training_res.csv")
        if name.startswith("random_target_e") and os.path.exists(p):
            parts = name.split("_")
            start = int([x for x in parts if x.startswith("e") and x[1:].isdigit()][0][1:])
            length = int([x for x in parts if x.startswith("l") and x[1:].isdigit()][0][1:])
            runs.append((name, start, length, pd.read_csv(p)))
    full = ref.recovery_table(base, runs)
    pick = ["random_target_e1_l2", "random_target_e10_l50", "random_target_e3_l20", "random_target_e90_l5",
            "random_target_e40_l30", "random_target_e22_l5"]
    sub = [r for r in runs if r[0] in pick]
    sub_table = ref.recovery_table(base, sub)
    # ---- single-epoch sweep (FIG3)
    sruns = {}
    for name in sorted(os.listdir(SWEEP_DIR)):
        if name.startswith("training_run"):
            n = name.split("run")[1]
            p = os.path.join(SWEEP_DIR, name, f"training_res_run{n}.csv")
            if os.path.exists(p):
                sruns[int(n)] = pd.read_csv(p)
    d_loss = ref.deviation_at_perturbation_epoch(base, sruns, "test_loss")
    d_rsa = ref.deviation_at_perturbation_epoch(base, sruns, "behavioral_rsa_rho")
    spick = [1, 2, 15, 35, 56, 70, 97, 98]

    def rec(df):
        return [[None if pd.isna(v) else (int(v) if c == "epoch" else float(v)) for c, v in zip(COLS, row)]
                for row in df[COLS].itertuples(index=False)]

    def table(t):
        return [{k: (None if pd.isna(v) else (bool(v) if k == "recovered" else (v if isinstance(v, str) else int(v))))
                 for k, v in row.items()} for row in t.to_dict("records")]

    out = {
        "source": "reference Data/clip_results + oracle/analysis_ref.py (fig3 cells 4-10, fig4 cells 4-12)",
        "baseline_raw": rec(base_raw),
        "baseline_trimmed_epochs": int(len(base)),
        "length_runs": {name: {"start": s, "length": l, "rows": rec(df)} for name, s, l, df in sub},
        "length_expected": table(sub_table),
        "single_runs": {str(n): rec(sruns[n]) for n in spick if n in sruns},
        "single_expected": {"test_loss": [[e, float(v)] for e, v in d_loss if e in spick],
                            "behavioral_rsa_rho": [[e, float(v)] for e, v in d_rsa if e in spick]},
        "full_summary": {"n_length_runs": int(len(full)), "n_recovered": int(full["recovered"].sum()),
                         "sum_epochs_to_recovery": int(full["epochs_to_recovery"].dropna().sum()),
                         "n_single_runs": len(sruns), "n_single_with_deviation": len(d_loss),
                         "sum_delta_test_loss": float(sum(v for _, v in d_loss)),
                         "sum_delta_rsa": float(sum(v for _, v in d_rsa))},
    }
    out.update(fig2_golden(base))
    path = os.path.join(ROOT, "tests", "golden", "analysis.json")
    json.dump(out, open(path, "w"))
    print(path, os.path.getsize(path), "bytes;", out["full_summary"])


if __name__ == "__main__":
    main()

"""CPU tests of the analysis layer (hba.analysis) against the row-by-row restatement of the
reference's notebook cells (oracle/analysis_ref.py) and the golden excerpt of the reference's own shipped
result CSVs (tests/golden/analysis.json, oracle/make_analysis_golden.py)."""
import json
import os
import subprocess
import sys

import numpy as np
import pandas as pd
import pytest

from hba import analysis, sweep
from oracle import analysis_ref as ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "analysis.json")))
COLS = ["epoch", "test_loss", "behavioral_rsa_rho"]


def frame(rows):
    return pd.DataFrame(rows, columns=COLS)


def trimmed_baseline():
    raw = frame(GOLD["baseline_raw"])
    base = raw.loc[:raw["test_loss"].idxmin()].copy()
    assert len(base) == GOLD["baseline_trimmed_epochs"]
    return base


def test_recovery_table_matches_reference_golden():
    runs = [(name, r["start"], r["length"], frame(r["rows"])) for name, r in GOLD["length_runs"].items()]
    got = analysis.recovery_table(trimmed_baseline(), runs)
    want = pd.DataFrame(GOLD["length_expected"])
    assert list(got["run_name"]) == list(want["run_name"])
    for col in ("start_epoch", "length", "perturbation_end"):
        assert list(got[col]) == list(want[col])
    assert list(got["recovered"]) == list(want["recovered"])
    for a, b in zip(got["recovery_epoch"], want["recovery_epoch"]):
        assert (pd.isna(a) and pd.isna(b)) or int(a) == int(b)
    for a, b in zip(got["epochs_to_recovery"], want["epochs_to_recovery"]):
        assert (pd.isna(a) and pd.isna(b)) or int(a) == int(b)


def test_single_sweep_deviations_match_reference_golden():
    runs = {int(n): frame(rows) for n, rows in GOLD["single_runs"].items()}
    base = trimmed_baseline()
    for column in ("test_loss", "behavioral_rsa_rho"):
        got = analysis.deviation_at_perturbation_epoch(base, runs, column)
        want = GOLD["single_expected"][column]
        assert list(got["run"]) == [e for e, _ in want]
        assert np.allclose(got["delta_" + column].to_numpy(), [v for _, v in want], rtol=0, atol=1e-12)


@pytest.mark.parametrize("seed", range(6))
def test_vectorised_tables_equal_the_row_by_row_restatement(seed):
    """Random sweeps incl. the awkward cases: runs that never recover, runs longer than the trimmed baseline
    (fallback to the baseline minimum), missing epochs, unsorted rows, the window reaching past the run."""
    rng = np.random.default_rng(seed)
    n_base = int(rng.integers(20, 60))
    base_raw = pd.DataFrame({"epoch": np.arange(1, n_base + 1),
                             "test_loss": 100 * np.exp(-np.arange(n_base) / 15) + rng.normal(0, 1.0, n_base) + 30,
                             "behavioral_rsa_rho": rng.uniform(0.3, 0.8, n_base)})
    base = ref.trim_at_min_test_loss(base_raw)
    assert analysis.load_baseline is not None and len(base) >= 1
    runs, sruns = [], {}
    for k in range(25):
        start, length = int(rng.integers(1, 50)), int(rng.choice([1, 2, 5, 10, 20, 50]))
        n = int(rng.integers(5, 90))
        epochs = np.arange(1, n + 1)
        keep = rng.random(n) > 0.1
        loss = 30 + 100 * np.exp(-epochs / 15) + rng.normal(0, 2.0, n) + (k % 3 == 0) * 20
        df = pd.DataFrame({"epoch": epochs, "test_loss": loss, "behavioral_rsa_rho": rng.uniform(0.2, 0.8, n)})[keep]
        df = df.sample(frac=1.0, random_state=int(rng.integers(1 << 30)))          # unsorted on disk
        runs.append((f"random_target_e{start}_l{length}_{k}", start, length, df))
        sruns[int(rng.integers(1, 70))] = df
    got, want = analysis.recovery_table(base, runs), ref.recovery_table(base, runs)
    assert list(got["run_name"]) == list(want["run_name"])
    assert list(got["recovered"]) == list(want["recovered"])
    for a, b in zip(got["recovery_epoch"], want["recovery_epoch"]):
        assert (pd.isna(a) and pd.isna(b)) or int(a) == int(b)
    for column in ("test_loss", "behavioral_rsa_rho"):
        g = analysis.deviation_at_perturbation_epoch(base, sruns, column)
        w = ref.deviation_at_perturbation_epoch(base, sruns, column)
        assert list(g["run"]) == [e for e, _ in w]
        assert np.array_equal(g["delta_" + column].to_numpy(), np.array([v for _, v in w], dtype=np.float64))


def test_discovery_reads_the_layouts_the_sweep_driver_writes_and_cli_round_trip(tmp_path):
    base_cfg = {"output_base_directory": str(tmp_path / "single"), "perturb_type": "random_target"}
    rows = "epoch,train_loss,test_loss,behavioral_rsa_rho,behavioral_rsa_p_value\n"
    baseline = tmp_path / "baseline.csv"
    baseline.write_text(rows + "".join(f"{e},{100 - e},{90 - 2 * e},{0.4 + 0.01 * e},0.0\n" for e in range(1, 11)))
    for e in (2, 5):
        cfg = sweep.condition_config(base_cfg, {"training_run": e, "perturb_length": 1}, "sweep")
        with open(cfg["training_res_path"], "w") as f:
            f.write(rows + "".join(f"{k},{100 - k},{91 - 2 * k + (3 if k == e else 0)},{0.39 + 0.01 * k},0.0\n" for k in range(1, 11)))
    len_cfg = {"output_base_directory": str(tmp_path / "length"), "perturb_type": "random_target"}
    for e, l in ((2, 2), (3, 5)):
        cfg = sweep.condition_config(len_cfg, {"training_run": e, "perturb_length": l}, "length")
        with open(cfg["training_res_path"], "w") as f:
            f.write(rows + "".join(f"{k},{100 - k},{(200 if e <= k < e + l else 90 - 2 * k + (5 if k < e + l + 2 else 0))},0.5,0.0\n"
                                   for k in range(1, 11)))
    runs = analysis.discover_single_sweep(str(tmp_path / "single"))
    assert sorted(runs) == [2, 5]
    s = analysis.single_sweep_summary(str(baseline), str(tmp_path / "single"))
    assert list(s["run"]) == [2, 5] and np.allclose(s["delta_test_loss"], [4.0, 4.0]) and np.allclose(s["delta_behavioral_rsa_rho"], -0.01)
    lr = analysis.discover_length_runs(str(tmp_path / "length"))
    assert [(n, a, b) for n, a, b, _ in lr] == [("random_target_e2_l2", 2, 2), ("random_target_e3_l5", 3, 5)]
    t = analysis.length_grid_summary(str(baseline), str(tmp_path / "length"))
    assert list(t["perturbation_end"]) == [3, 7] and list(t["recovery_epoch"]) == [6, 10] and list(t["epochs_to_recovery"]) == [3, 3]
    out = tmp_path / "recovery.csv"
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "analyze_sweep.py"), "--kind", "length", "--baseline",
                        str(baseline), "--sweep-dir", str(tmp_path / "length"), "--out", str(out)],
                       capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    back = pd.read_csv(out)
    assert list(back["recovery_epoch"]) == [6, 10] and list(back["recovered"]) == [True, True]


@pytest.mark.skipif(not os.path.isdir("/root/reference/Data/clip_results"), reason="reference data not mounted")
def test_full_reference_result_set_summary():
    """All shipped conditions through the product functions: counts and checksums of the golden generator."""
    data = "/root/reference/Data/clip_results"
    base_csv = os.path.join(data, "baseline_clip_results_seed1.csv")
    t = analysis.length_grid_summary(base_csv, os.path.join(data, "perturb_length_experiments_baselineseed1_perturbseed0"))
    s = analysis.single_sweep_summary(base_csv, os.path.join(data, "single_sweep_experiments"))
    f = GOLD["full_summary"]
    assert len(t) == f["n_length_runs"] and int(t["recovered"].sum()) == f["n_recovered"]
    assert int(t["epochs_to_recovery"].dropna().sum()) == f["sum_epochs_to_recovery"]
    assert len(s) == f["n_single_with_deviation"]
    assert abs(float(s["delta_test_loss"].sum()) - f["sum_delta_test_loss"]) < 1e-9
    assert abs(float(s["delta_behavioral_rsa_rho"].sum()) - f["sum_delta_rsa"]) < 1e-9

#!/usr/bin/env python
"""CSV-in / CSV-out analysis of a finished sweep (hba.analysis; the numbers behind the reference's fig3 / fig4).

  python tools/analyze_sweep.py --kind single --baseline base/res.csv --sweep-dir out/ --out single_summary.csv
  python tools/analyze_sweep.py --kind length --baseline base/res.csv --sweep-dir out/ --out recovery.csv
  python tools/analyze_sweep.py --kind types --baseline base/res.csv --sweep-dir Data/clip_results --out fig2.csv
  python tools/analyze_sweep.py --kind vit-summary --effects perturbation_effects.csv --out perturbation_summary_table.csv
"""
import argparse
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_analysis():
    # hba.analysis is pure pandas: load it by path so that the CLI also works where libhba.so is not built
    spec = importlib.util.spec_from_file_location(
        "hba_analysis", os.path.join(ROOT, "vit-project_b200", "hba", "analysis.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", choices=["single", "length", "types", "vit-summary"], required=True)
    ap.add_argument("--baseline", help="baseline run CSV (epoch,train_loss,test_loss,behavioral_rsa_rho,...)")
    ap.add_argument("--sweep-dir", help="sweep output directory (types: the tree holding one directory per perturbation type)")
    ap.add_argument("--effects", help="vit-summary: the result CSV of measure_single_epoch_perturbation_effect.py")
    ap.add_argument("--prefix", default="random_target", help="length grid: run directories are {prefix}_e{E}_l{L}")
    ap.add_argument("--out", default="-")
    a = ap.parse_args()
    an = _load_analysis()
    if a.kind == "vit-summary":
        if not a.effects:
            ap.error("--kind vit-summary needs --effects")
        df = an.vit_perturbation_summary(a.effects)
    else:
        if not (a.baseline and a.sweep_dir):
            ap.error(f"--kind {a.kind} needs --baseline and --sweep-dir")
        df = {"single": lambda: an.single_sweep_summary(a.baseline, a.sweep_dir),
              "length": lambda: an.length_grid_summary(a.baseline, a.sweep_dir, a.prefix),
              "types": lambda: an.perturbation_type_summary(a.baseline, a.sweep_dir)}[a.kind]()
    if a.out == "-":
        df.to_csv(sys.stdout, index=False)
    else:
        df.to_csv(a.out, index=False)
        print(f"{len(df)} rows -> {a.out}")


if __name__ == "__main__":
    main()

# round 2, call B: the whole GPU suite (no -x: every failure in one pass), kernel bandwidth table
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 > gpurun_out/r02b_tests.log 2>&1
echo "gpu tests rc=$?"; tail -25 gpurun_out/r02b_tests.log
timeout 300 python tools/bench_kernels.py --reps 10 > gpurun_out/r02b_kernels.json 2> gpurun_out/r02b_kernels.err
echo "bench_kernels rc=$?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/r02b_kernels.json"))
for k, v in d["kernels"].items():
    print(f"{k:45s} {v['ms']*1e3:8.1f} us  {v['achieved_gbs']:8.1f} GB/s  {v['frac_of_measured_peak']:.3f}")
PY

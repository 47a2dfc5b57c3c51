# round 2, call C: the previously failing tests, the new DoRA / LayerNorm kernels, RSA NaN debug, kernel table
mkdir -p gpurun_out
timeout 300 python tools/debug/rsa_nan.py > gpurun_out/r02c_rsa_nan.log 2>&1; echo "rsa nan debug rc=$?"; cat gpurun_out/r02c_rsa_nan.log | tail -8
timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_pipeline.py tests/test_gpu_sweep.py tests/test_gpu_zz_sweep_chain.py tests/test_gpu_model.py -m gpu -q -p no:cacheprovider --timeout 900 > gpurun_out/r02c_tests.log 2>&1
echo "gpu tests rc=$?"; tail -12 gpurun_out/r02c_tests.log
timeout 300 python tools/bench_kernels.py --reps 10 > gpurun_out/r02c_kernels.json 2> gpurun_out/r02c_kernels.err
echo "bench_kernels rc=$?"; tail -3 gpurun_out/r02c_kernels.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r02c_kernels.json"))
print("empty launch", d.get("empty_launch_us"))
for k, v in d["kernels"].items():
    print(f"{k:45s} {v['ms']*1e3:8.1f} us  {v['achieved_gbs']:8.1f} GB/s  {v['frac_of_measured_peak']:.3f}")
PY
HBA_DORA_CLUSTER=0 HBA_LN_STREAM=0 timeout 300 python tools/bench_kernels.py --reps 10 > gpurun_out/r02c_kernels_old.json 2>/dev/null
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02c_kernels_old.json"))
print("OLD kernels (HBA_DORA_CLUSTER=0 HBA_LN_STREAM=0), same timing method")
for k, v in d["kernels"].items():
    print(f"{k:45s} {v['ms']*1e3:8.1f} us  {v['achieved_gbs']:8.1f} GB/s  {v['frac_of_measured_peak']:.3f}")
PY

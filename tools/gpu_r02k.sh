# round 2, call K: validation of the LayerNorm (24 consumer warps), RDM (4x4 micro-tiles) and radix (2048-key tiles) changes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py tests/test_gpu_fullsize_parity.py -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/r02k_tests.log 2>&1
echo "gpu tests rc=$?"; tail -4 gpurun_out/r02k_tests.log
timeout 300 python tools/bench_kernels.py --reps 10 > gpurun_out/r02k_kernels.json 2> gpurun_out/r02k_kernels.err
echo "bench_kernels rc=$?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/r02k_kernels.json"))
print("empty launch", d.get("empty_launch_us"))
for k, v in d["kernels"].items():
    print(f"{k:45s} {v['ms']*1e3:8.1f} us  {v['achieved_gbs']:8.1f} GB/s  {v['frac_of_measured_peak']:.3f}  in-stream {v.get('ms_in_stream', 0)*1e3:6.1f} us {v.get('frac_of_measured_peak_in_stream', 0):.3f}")
PY
B="python bench.py --steps 30 --warmup 5 --no-sweep --no-cpu-baseline --no-hbm-kernels --no-fp32 --roofline-seconds 0.5"
timeout 600 $B > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err
echo "bench rc=$? $(python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02k_bench.json') if l.startswith('{')][-1])
print('ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'gemm TF', round(d['roofline']['achieved'],1), 'vit', round(d['vit_b16']['value'],1), d['roofline']['kernel_time_shares'])
")"
HBA_ASYNC_CKPT=1 timeout 600 python tools/grid_sweep_bench.py --kind grid --gpus 0 --limit 12 --max-start 10 --root /tmp/hba_grid_k --out gpurun_out/r02k_grid12_async.json > gpurun_out/r02k_grid12_async.log 2>&1
echo "grid 12 HBA_ASYNC_CKPT=1 rc=$? $(tail -1 gpurun_out/r02k_grid12_async.log | cut -c1-330)"
timeout 600 python tools/grid_sweep_bench.py --kind grid --gpus 0 --limit 12 --max-start 10 --root /tmp/hba_grid_k --out gpurun_out/r02k_grid12.json > gpurun_out/r02k_grid12.log 2>&1
echo "grid 12 rc=$? $(tail -1 gpurun_out/r02k_grid12.log | cut -c1-330)"
python - <<'PY'
import json
a, b = (json.load(open(f"gpurun_out/r02k_grid12{t}.json")) for t in ("_async", ""))
print("digests equal:", a["trajectories_digest"] == b["trajectories_digest"], a["trajectories_digest"], b["trajectories_digest"])
PY

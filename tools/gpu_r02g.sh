# round 2, call G: GEMM tail splitting (tests, A/B), then the FULL 136-condition grid through the scheduler on ONE GPU
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_fullsize.py tests/test_gpu_fullsize_parity.py tests/test_gpu_vit.py -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/r02g_tests.log 2>&1
echo "gpu tests rc=$?"; tail -6 gpurun_out/r02g_tests.log
B="python bench.py --steps 30 --warmup 5 --no-sweep --no-cpu-baseline --no-hbm-kernels --no-fp32 --roofline-seconds 0.5"
for ts in 1 0 1 0; do
  HBA_GEMM_TAIL_SPLIT=$ts timeout 600 $B > gpurun_out/r02g_bench_ts$ts.json 2> gpurun_out/r02g_bench_ts$ts.err
  echo "HBA_GEMM_TAIL_SPLIT=$ts rc=$? $(python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02g_bench_ts$ts.json') if l.startswith('{')][-1])
print('ms/step', round(d['ms_per_step'],4), 'gemm TF', round(d['roofline']['achieved'],1), 'vit', round(d['vit_b16']['value'],1), [ (s['N'],s['K'],round(s['us_per_launch'],1)) for s in d['roofline']['by_shape'][:8]])
")"
done
timeout 3000 python tools/grid_sweep_bench.py --kind grid --gpus 0 --root /tmp/hba_grid_full --out gpurun_out/r02g_grid_full_n1.json > gpurun_out/r02g_grid_full_n1.log 2>&1
echo "full grid N=1 rc=$?"; tail -1 gpurun_out/r02g_grid_full_n1.log | cut -c1-1200

# round 2, call E: skinny split-K + tensor-map cache (default), then programmatic dependent launch A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_fullsize.py -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/r02e_tests.log 2>&1
echo "gpu tests rc=$?"; tail -6 gpurun_out/r02e_tests.log
B="python bench.py --steps 30 --warmup 5 --no-sweep --no-vit --no-cpu-baseline --no-fp32 --no-hbm-kernels --roofline-seconds 0.3"
for pdl in 0 1 0 1; do
  HBA_PDL=$pdl timeout 600 $B > gpurun_out/r02e_bench_pdl$pdl.json 2> gpurun_out/r02e_bench_pdl$pdl.err
  echo "HBA_PDL=$pdl rc=$? $(python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02e_bench_pdl$pdl.json') if l.startswith('{')][-1])
print('ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'eager', round(d['roofline']['eager_ms_per_step'],3), 'gemm TF', round(d['roofline']['achieved'],1))
")"; tail -2 gpurun_out/r02e_bench_pdl$pdl.err
done
HBA_PDL=1 timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py -m gpu -q -p no:cacheprovider --timeout 600 -x > gpurun_out/r02e_tests_pdl.log 2>&1
echo "gpu tests with HBA_PDL=1 rc=$?"; tail -6 gpurun_out/r02e_tests_pdl.log
HBA_SKINNY_SPLITK=0 timeout 600 $B > gpurun_out/r02e_bench_noskinny.json 2>/dev/null
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02e_bench_noskinny.json') if l.startswith('{')][-1])
print('HBA_SKINNY_SPLITK=0 ms/step', round(d['ms_per_step'],4))"

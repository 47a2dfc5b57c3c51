mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 --no-vit --no-cpu-baseline --no-sweep > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_tmp.json')); r=d['roofline']; print(d['value'], d['ms_per_step'], d['e2e']['value'], r['frac'], r['gemm_share_of_step'], r['kernel_time_shares'], r['eager_ms_per_step'])"; tail -3 gpurun_out/bench_tmp.err

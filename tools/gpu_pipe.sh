mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_model.py -q -x -p no:cacheprovider > gpurun_out/t_pipe.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t_pipe.log
timeout 300 python bench.py --sweep-only > gpurun_out/sweep_only.json 2> gpurun_out/sweep_only.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/sweep_only.json')); print({k:v for k,v in d.items() if k not in ('config','reference_baseline')})"
timeout 300 python bench.py --steps 20 --no-vit --no-cpu-baseline --no-sweep > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_tmp.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'])"

mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_model.py -q -x -p no:cacheprovider > gpurun_out/t_pipe.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t_pipe.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-vit --no-cpu-baseline --no-sweep > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_tmp.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['frac'], d['roofline']['eager_ms_per_step'])"; tail -3 gpurun_out/bench_tmp.err
HBA_TEXT_STREAM=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-vit --no-cpu-baseline --no-sweep > gpurun_out/bench_tmp1.json 2> gpurun_out/bench_tmp1.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_tmp1.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['frac'], d['roofline']['eager_ms_per_step'])"

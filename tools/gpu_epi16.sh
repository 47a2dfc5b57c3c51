mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_vit.py tests/test_gpu_model.py -q -x -p no:cacheprovider > gpurun_out/t_vit.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t_vit.log
timeout 300 python tools/bench_vit.py --batch 256 --steps 5 --warmup 3 > gpurun_out/vit_n1.json 2> gpurun_out/vit_n1.err; cat gpurun_out/vit_n1.json; tail -3 gpurun_out/vit_n1.err
timeout 300 python bench.py --steps 20 --no-vit --no-cpu-baseline --no-sweep > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_tmp.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])
for r in d['roofline']['by_shape'][:8]: print(r)"

mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --no-vit --no-cpu-baseline --no-sweep > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_tmp.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'])
for r in d['roofline']['by_shape']: print(r)"

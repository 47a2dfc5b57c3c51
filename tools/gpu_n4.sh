mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"; python -c "
import json,sys
ls=[l for l in open('gpurun_out/bench_n$N.json')]
print(len(ls),'lines')
d=json.loads(ls[0]); print(d['n_gpus'], d['value'], d['e2e']['value'], d['sweep'].get('conditions_per_hour'), d['sweep'].get('sec_per_epoch_cached'), d['vit_b16'].get('value'), d['vit_b16'].get('ms_per_step'), d['host_cpus'])"; grep -v "^$" gpurun_out/bench_n$N.err | grep -v -i "warn" | tail -3 | cut -c1-200

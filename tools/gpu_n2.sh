mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json')); print(d['n_gpus'], d['value'], d['e2e']['value'], d['sweep']['conditions_per_hour'], d['vit_b16']['value'], d['vit_b16']['ms_per_step'])"; tail -4 gpurun_out/bench_n2.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref n2 rc=$?"; cut -c1-300 gpurun_out/bench_ref_n2.json

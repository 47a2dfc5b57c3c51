mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/bench_vit.py --batch 256 --steps 5 --warmup 3 > gpurun_out/vit_n2.json 2> gpurun_out/vit_n2.err; echo "vit n2 rc=$?"; cat gpurun_out/vit_n2.json; tail -5 gpurun_out/vit_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/bench_vit.py --batch 256 --steps 5 --warmup 3 --no-graph > gpurun_out/vit_n2_nograph.json 2> gpurun_out/vit_n2_nograph.err; echo "vit n2 nograph rc=$?"; cat gpurun_out/vit_n2_nograph.json; tail -5 gpurun_out/vit_n2_nograph.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"; cat gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json

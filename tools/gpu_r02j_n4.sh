# round 2, call J (4 GPUs): ViT-B/16 data parallel with / without SMs reserved for NCCL during the backward pass
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
for ctas in 0 132 116; do
  HBA_DP_GEMM_CTAS=$ctas timeout 600 $TR --master-port 2953$((ctas % 10)) bench.py --gpus 4 --steps 10 --warmup 3 --no-sweep --no-cpu-baseline --no-hbm-kernels --no-fp32 --roofline-seconds 0.2 > gpurun_out/r02j_bench_n4_ctas$ctas.json 2> gpurun_out/r02j_bench_n4_ctas$ctas.err
  echo "bench N=4 HBA_DP_GEMM_CTAS=$ctas rc=$? $(python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02j_bench_n4_ctas$ctas.json') if l.startswith('{')][-1])
print('clip img/s', round(d['value']), 'vit img/s', round(d['vit_b16']['value']), 'vit ms/step', round(d['vit_b16']['ms_per_step'],2))
")"; tail -2 gpurun_out/r02j_bench_n4_ctas$ctas.err | cut -c1-200
done

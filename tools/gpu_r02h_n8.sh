# round 2, call H (8 GPUs): the FULL 136-condition grid through hba.sweep.run_sweep on 8 pinned workers, CLIP RSA at scale over
# its checkpoints sharded over 8 ranks, and the 8-rank bench (ViT-B/16 data parallel with / without SMs reserved for NCCL)
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -p no:cacheprovider -k "rdm or spearman or rsa_nan" 2>&1 | tail -2
timeout 1500 python tools/grid_sweep_bench.py --kind grid --gpus 0,1,2,3,4,5,6,7 --keep --root /tmp/hba_grid_n8 --out gpurun_out/r02h_grid_full_n8.json > gpurun_out/r02h_grid_full_n8.log 2>&1
echo "full grid N=8 rc=$?"; tail -1 gpurun_out/r02h_grid_full_n8.log | cut -c1-1500
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 tools/clip_rsa_over_checkpoints.py --checkpoints /tmp/hba_grid_n8 --csv-file /tmp/hba_grid_n8/train.csv --inference-csv-file /tmp/hba_grid_n8/rsa.csv --img-dir /tmp/hba_grid_n8/imgs --limit 2400 --output-csv gpurun_out/r02h_clip_rsa_scale_n8.csv > gpurun_out/r02h_clip_rsa_scale_n8.log 2>&1
echo "clip rsa at scale N=8 rc=$?"; grep '^{"metric"' gpurun_out/r02h_clip_rsa_scale_n8.log | cut -c1-700
rm -rf /tmp/hba_grid_n8
for ctas in 0; do
  HBA_DP_GEMM_CTAS=$ctas timeout 900 $TR --master-port 2952$((ctas % 10)) bench.py --gpus 8 --steps 10 --warmup 3 --no-sweep --no-cpu-baseline --no-hbm-kernels --no-fp32 --roofline-seconds 0.3 > gpurun_out/r02h_bench_n8_ctas$ctas.json 2> gpurun_out/r02h_bench_n8_ctas$ctas.err
  echo "bench N=8 HBA_DP_GEMM_CTAS=$ctas rc=$? $(python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02h_bench_n8_ctas$ctas.json') if l.startswith('{')][-1])
print('clip img/s', round(d['value']), 'vit img/s', round(d['vit_b16']['value']), 'vit ms/step', round(d['vit_b16']['ms_per_step'],2))
")"; tail -2 gpurun_out/r02h_bench_n8_ctas$ctas.err | cut -c1-200
done

# round 2, call F: suite with PDL on by default; grid tool + CLIP RSA-at-scale tool on a small slice; quick bench;
# several sweep workers per GPU (time-sliced, then under MPS)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 > gpurun_out/r02f_tests.log 2>&1
echo "gpu tests rc=$?"; tail -6 gpurun_out/r02f_tests.log
timeout 900 python tools/grid_sweep_bench.py --kind grid --gpus 0 --limit 3 --max-start 3 --keep --root /tmp/hba_grid_f --out gpurun_out/r02f_grid_slice.json > gpurun_out/r02f_grid_slice.log 2>&1
echo "grid slice rc=$?"; tail -3 gpurun_out/r02f_grid_slice.log | cut -c1-600
timeout 900 python tools/clip_rsa_over_checkpoints.py --checkpoints /tmp/hba_grid_f --csv-file /tmp/hba_grid_f/train.csv --inference-csv-file /tmp/hba_grid_f/rsa.csv --img-dir /tmp/hba_grid_f/imgs --output-csv gpurun_out/r02f_clip_rsa_scale.csv > gpurun_out/r02f_clip_rsa_scale.log 2>&1
echo "clip rsa at scale rc=$?"; tail -2 gpurun_out/r02f_clip_rsa_scale.log | cut -c1-900; head -4 gpurun_out/r02f_clip_rsa_scale.csv
B="python bench.py --steps 30 --warmup 5 --no-sweep --no-vit --no-cpu-baseline --no-hbm-kernels --roofline-seconds 0.3"
for pdl in 1 0; do
  HBA_PDL=$pdl timeout 600 $B > gpurun_out/r02f_bench_pdl$pdl.json 2> gpurun_out/r02f_bench_pdl$pdl.err
  echo "HBA_PDL=$pdl rc=$? $(python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02f_bench_pdl$pdl.json') if l.startswith('{')][-1])
print('ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'fp32', d.get('fp32_mode'))
")"
done
# N1 probe: k worker processes on ONE GPU, 12 grid conditions (start <= 10)
for k in 1 2 4; do
  timeout 900 python tools/grid_sweep_bench.py --kind grid --gpus 0 --workers-per-gpu $k --limit 12 --max-start 10 --root /tmp/hba_grid_k --out gpurun_out/r02f_grid_k$k.json > gpurun_out/r02f_grid_k$k.log 2>&1
  echo "workers/gpu=$k rc=$? $(tail -1 gpurun_out/r02f_grid_k$k.log | cut -c1-420)"
done
which nvidia-cuda-mps-control && {
  export CUDA_MPS_PIPE_DIRECTORY=/tmp/mps_pipe CUDA_MPS_LOG_DIRECTORY=/tmp/mps_log; mkdir -p $CUDA_MPS_PIPE_DIRECTORY $CUDA_MPS_LOG_DIRECTORY
  nvidia-cuda-mps-control -d && sleep 2
  for k in 2 4; do
    timeout 900 python tools/grid_sweep_bench.py --kind grid --gpus 0 --workers-per-gpu $k --limit 12 --max-start 10 --root /tmp/hba_grid_k --out gpurun_out/r02f_grid_mps_k$k.json > gpurun_out/r02f_grid_mps_k$k.log 2>&1
    echo "MPS workers/gpu=$k rc=$? $(tail -1 gpurun_out/r02f_grid_mps_k$k.log | cut -c1-420)"
  done
  echo quit | nvidia-cuda-mps-control
}

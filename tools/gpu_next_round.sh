# First gpurun call of the next round (nothing below has run on a B200 yet; written after round 1's GPU budget
# was spent).  Usage:  gpurun --timeout 3600 -- 'bash tools/gpu_next_round.sh'   (typically ~35 min of box time; every
# step has its own timeout - comment out sections 4-6 for a quick first call)
#   1. the new GPU tests on their own (ViT measurement path, chained sweep), so that a failure there does not hide
#      behind the -x of the full suite;
#   2. compute-sanitizer memcheck + racecheck over the kernel unit tests of the HBM-bound ops (graphs off,
#      bounded: the tcgen05 kernels run ~100x slower under the tool, so they get one small case each);
#   3. timing of one measurement epoch (ViT-B/16, batch 256, resident synthetic data) next to bench.py's
#      ViT-B/16 step rate.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vit_measure.py -q -p no:cacheprovider > gpurun_out/t_vit_measure.log 2>&1
echo "vit measure tests rc=$?"; tail -5 gpurun_out/t_vit_measure.log
timeout 600 python -m pytest tests/test_gpu_zz_sweep_chain.py -q -p no:cacheprovider > gpurun_out/t_chain.log 2>&1
echo "chain test rc=$?"; tail -5 gpurun_out/t_chain.log
export HBA_STEP_GRAPH=0
for tool in memcheck racecheck; do
  timeout 420 compute-sanitizer --tool $tool --error-exitcode 9 --launch-timeout 60 \
      python -m pytest tests/test_gpu_ops.py -q -x -p no:cacheprovider \
      -k "layernorm or dora_merge or cos_head or softmax_ce or adamw or sgd_multi or rdm or rank_avg or spearman or embed" \
      > gpurun_out/sanitizer_${tool}.log 2>&1
  echo "compute-sanitizer $tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_${tool}.log | tail -3
done
unset HBA_STEP_GRAPH
timeout 300 python - > gpurun_out/vit_measure_epoch.json 2> gpurun_out/vit_measure_epoch.err <<'PY'
import json, sys, time
sys.path[:0] = [".", "vit-project_b200"]
import torch
from hba import vit, vit_train as vt
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
n, bs = 4096, 256
train = vt.synthetic_imagenet(n, 1000, seed=0, device=dev)
val = vt.synthetic_imagenet(1024, 1000, seed=1, device=dev)
things, rdm = vt.synthetic_things(dev)
model = vit.create_model("vit_base_patch16_224", num_classes=1000).to(dev)
tr = vit.DataParallelTrainer(model, use_graph=True)
out = {}
for kind in (None, "gaussian", "label_shuffle"):
    loader = vt.ShardedLoader(train, bs, shuffle=True, perturbation_type=kind)
    vt.train_one_epoch(tr, loader, 0, log=None)          # warm-up epoch (capture)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    vt.train_one_epoch(tr, loader, 1, log=None)
    torch.cuda.synchronize(); out[f"train_img_per_s_{kind}"] = n / (time.perf_counter() - t0)
vl = vt.ShardedLoader(val, bs)
vt.validate(tr, vl); torch.cuda.synchronize(); t0 = time.perf_counter()
vt.validate(tr, vl); torch.cuda.synchronize(); out["validate_img_per_s"] = 1024 / (time.perf_counter() - t0)
tl = vt.ShardedLoader(things, 8, with_names=True)
vt.compute_rsa_score(model, tl, rdm); torch.cuda.synchronize(); t0 = time.perf_counter()
vt.compute_rsa_score(model, tl, rdm); torch.cuda.synchronize(); out["rsa_48_images_ms"] = 1e3 * (time.perf_counter() - t0)
print(json.dumps(out))
PY
echo "measure epoch timing rc=$?"; cat gpurun_out/vit_measure_epoch.json
#   4. N1 probe: does a second / third sweep worker PROCESS on the same GPU raise conditions/hour (the cached sweep
#      epoch leaves the GPU idle during its host phases: CSV, checkpoints, RSA read-back)?  Sum of the per-process
#      conditions/h of k concurrent `bench.py --sweep-only` runs against k = 1.
for k in 1 2 3; do
  pids=""
  for i in $(seq 1 $k); do
    timeout 400 python bench.py --sweep-only > gpurun_out/sweep_k${k}_p${i}.json 2> gpurun_out/sweep_k${k}_p${i}.err &
    pids="$pids $!"
  done
  for p in $pids; do wait $p; done
  python - $k <<'PY'
import glob, json, sys
k = int(sys.argv[1]); tot = 0.0
for f in sorted(glob.glob(f"gpurun_out/sweep_k{k}_p*.json")):
    lines = [l for l in open(f) if l.startswith("{")]
    if lines:
        d = json.loads(lines[-1]); tot += d.get("sweep", d).get("conditions_per_hour", 0.0)
print(f"sweep workers on one GPU: k={k} total conditions/h = {tot:.1f}")
PY
done
#   5. background checkpoint writer (HBA_ASYNC_CKPT=1, opt-in): the pipeline tests under it, then the sweep epoch time
#      with and without it.
HBA_ASYNC_CKPT=1 timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_sweep.py -q -x -p no:cacheprovider \
    > gpurun_out/t_async_ckpt.log 2>&1; echo "pipeline tests with HBA_ASYNC_CKPT=1 rc=$?"; tail -3 gpurun_out/t_async_ckpt.log
for v in 0 1; do
  HBA_ASYNC_CKPT=$v timeout 300 python bench.py --sweep-only > gpurun_out/sweep_async${v}.json 2> gpurun_out/sweep_async${v}.err
  echo "HBA_ASYNC_CKPT=$v: $(grep -o '"sec_per_epoch_cached": [0-9.]*' gpurun_out/sweep_async${v}.json)"
done
#   6. where the two weakest HBM-roofline fractions come from: launch list of the kernel benchmark (radix passes of the
#      RSA-at-scale ranking, LayerNorm forward), then one full capture of the single-CTA scan and of LayerNorm.
CMD="python tools/bench_kernels.py --reps 5"
timeout 300 $CMD > gpurun_out/kernels_r02.json 2> gpurun_out/kernels_r02.err; echo "bench_kernels rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_kernels_r02.csv \
    $CMD > gpurun_out/ncu_kernels_list.log 2>&1; echo "ncu launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:radix_scan_kernel -c 2 -o gpurun_out/radix_scan_r02 -f \
    $CMD > gpurun_out/ncu_radix_scan.log 2>&1; echo "ncu radix_scan rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:layernorm_fwd_kernel -c 2 -o gpurun_out/layernorm_fwd_r02 -f \
    $CMD > gpurun_out/ncu_layernorm.log 2>&1; echo "ncu layernorm rc=$?"

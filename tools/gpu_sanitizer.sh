# compute-sanitizer memcheck + racecheck + synccheck over the kernel unit tests: the HBM-bound kernels in full,
# the tcgen05 kernels on their smallest cases (they run ~100x slower under the tool).  Summaries -> gpurun_out/.
mkdir -p gpurun_out
export HBA_STEP_GRAPH=0
K="layernorm or dora_merge or dora_layer or cos_head or cos_mse or softmax_ce or adamw or sgd_multi or rdm or rank_avg or spearman or rsa_nan or embed or patch_embed"
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --launch-timeout 120 \
      python -m pytest tests/test_gpu_ops.py -q -x -p no:cacheprovider -k "$K" > gpurun_out/sanitizer_${tool}_hbm.log 2>&1
  echo "compute-sanitizer $tool (HBM-bound kernels) rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_${tool}_hbm.log | tail -3
done
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --launch-timeout 300 \
      python -m pytest tests/test_gpu_ops.py -q -x -p no:cacheprovider \
      -k "gemm_plain and 128-256-64 or gemm_skinny and 32-768-1024-4 or attention_fwd and 1-50-1" > gpurun_out/sanitizer_${tool}_tc.log 2>&1
  echo "compute-sanitizer $tool (tcgen05 kernels, smallest cases) rc=$?"; grep -E "ERROR SUMMARY|passed|failed|deselected" gpurun_out/sanitizer_${tool}_tc.log | tail -3
done

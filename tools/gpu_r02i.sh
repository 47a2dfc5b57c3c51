# round 2, call I: PDL on LayerNorm / attention as secondaries (tests, A/B), compute-sanitizer, ncu launch list + full captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py tests/test_gpu_fullsize_parity.py tests/test_gpu_vit.py -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/r02i_tests.log 2>&1
echo "gpu tests rc=$?"; tail -4 gpurun_out/r02i_tests.log
B="python bench.py --steps 30 --warmup 5 --no-sweep --no-cpu-baseline --no-hbm-kernels --no-fp32 --roofline-seconds 0.5"
for pdl in 1 0 1 0; do
  HBA_PDL=$pdl timeout 600 $B > gpurun_out/r02i_bench_pdl$pdl.json 2> gpurun_out/r02i_bench_pdl$pdl.err
  echo "HBA_PDL=$pdl rc=$? $(python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02i_bench_pdl$pdl.json') if l.startswith('{')][-1])
print('ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'gemm TF', round(d['roofline']['achieved'],1), 'vit', round(d['vit_b16']['value'],1))
")"
done
bash tools/gpu_sanitizer.sh
CMD="python bench.py --steps 2 --warmup 3 --no-sweep --no-vit --no-cpu-baseline --no-hbm-kernels --no-fp32 --roofline-seconds 0.01"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_bench.csv $CMD > gpurun_out/ncu_list.log 2>&1; echo "ncu launch list rc=$?"
KC="python tools/bench_kernels.py --reps 2 --cpu-reps 1"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 600 -c 12 -o gpurun_out/r02_gemm_full -f $CMD > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:layernorm_fwd_stream_kernel -s 3 -c 2 -o gpurun_out/r02_ln_stream_full -f $KC > gpurun_out/ncu_ln.log 2>&1; echo "ncu ln rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dora_ -s 6 -c 6 -o gpurun_out/r02_dora_full -f $KC > gpurun_out/ncu_dora.log 2>&1; echo "ncu dora rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:radix_pass_kernel|rank_final_kernel|rdm_kernel|radix_hist_kernel" -s 30 -c 11 -o gpurun_out/r02_rsa_full -f $KC > gpurun_out/ncu_rsa.log 2>&1; echo "ncu rsa rc=$?"
ls -la gpurun_out/*.ncu-rep

mkdir -p gpurun_out
CMD="python tools/bench_gemm.py --only c_fc+quickgelu,qkv --seconds 0.005 --no-cublas"
$CMD > gpurun_out/ncu_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 8 -c 1 -o gpurun_out/gemm_cfc_r01c -f $CMD > gpurun_out/ncu_gemm.log 2>&1
tail -5 gpurun_out/ncu_gemm.log

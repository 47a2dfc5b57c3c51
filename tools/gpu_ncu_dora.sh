mkdir -p gpurun_out
CMD="python tools/bench_kernels.py --reps 2 --cpu-reps 1"
$CMD > gpurun_out/ncu_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dora_merge_bwd -s 6 -c 2 -o gpurun_out/dora_r01 -f $CMD > gpurun_out/ncu_dora.log 2>&1
tail -3 gpurun_out/ncu_dora.log

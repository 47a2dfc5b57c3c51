mkdir -p gpurun_out
CMD="python tools/bench_attn.py"
$CMD > gpurun_out/ncu_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 110 -c 1 -o gpurun_out/attn_r01d -f $CMD > gpurun_out/ncu_attn.log 2>&1
tail -3 gpurun_out/ncu_attn.log

mkdir -p gpurun_out
CMD="python tools/bench_vit.py --batch 256 --steps 1 --warmup 3 --profile"
$CMD > gpurun_out/plain.log 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_vit_r01e.csv $CMD > gpurun_out/ncu1.log 2>&1; echo "launch list rc=$?"
$CMD > gpurun_out/plain.log 2>&1 && ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attention_bwd_tc -c 1 -o gpurun_out/attn_bwd_r01e -f $CMD > gpurun_out/ncu2.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/ncu2.log

mkdir -p gpurun_out
CMD="python tools/bench_vit.py --batch 256 --steps 1 --warmup 2"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv -s 964 --log-file gpurun_out/launches_vit.csv $CMD > gpurun_out/ncu1.log 2>&1; echo "rc=$?"

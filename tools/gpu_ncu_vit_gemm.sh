mkdir -p gpurun_out
CMD="python tools/bench_vit.py --batch 256 --steps 1 --warmup 3 --profile --no-graph"
$CMD > gpurun_out/plain.log 2>&1 || exit 1
ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:Li256ELi2ELi4E -s 5 -c 1 -o /tmp/gemm_act4 -f $CMD > gpurun_out/ncu2.log 2>&1; echo "ncu act4 rc=$?"
ncu -i /tmp/gemm_act4.ncu-rep --page raw --csv > gpurun_out/gemm_act4_raw.csv 2>/dev/null
ncu -i /tmp/gemm_act4.ncu-rep --page source --csv --print-source sass > gpurun_out/gemm_act4_sass.csv 2>/dev/null
ls -la gpurun_out/gemm_act4_raw.csv gpurun_out/gemm_act4_sass.csv

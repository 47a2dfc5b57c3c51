mkdir -p gpurun_out
CMD="python tools/bench_vit.py --batch 256 --steps 1 --warmup 3 --profile"
$CMD > gpurun_out/plain.log 2>&1 || exit 1
ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:Li256ELi2ELi4E -c 1 -o gpurun_out/gemm_act4_r01e -f $CMD > gpurun_out/ncu2.log 2>&1; echo "ncu act4 rc=$?"; tail -2 gpurun_out/ncu2.log
ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:Li256ELi2ELi2E -c 1 -o gpurun_out/gemm_act2_r01e -f $CMD > gpurun_out/ncu3.log 2>&1; echo "ncu act2 rc=$?"; tail -2 gpurun_out/ncu3.log

mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench.err
python tools/profile_step.py > gpurun_out/plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu1.log 2>&1; echo "launch-list rc=$?"
python tools/profile_step.py > gpurun_out/plain2.log 2>&1 && \
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc -s 20 -c 4 \
    -o gpurun_out/gemm_r01 python tools/profile_step.py > gpurun_out/ncu2.log 2>&1; echo "full rc=$?"
ls -la gpurun_out

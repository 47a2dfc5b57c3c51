mkdir -p gpurun_out
python tools/profile_step.py > gpurun_out/plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_r01b.csv python tools/profile_step.py > gpurun_out/ncu1.log 2>&1; echo "launch-list rc=$?"
python tools/profile_step.py > gpurun_out/plain2.log 2>&1 && \
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"gemm_tc|attention_tc" -s 24 -c 6 \
    -o gpurun_out/gemm_attn_r01b python tools/profile_step.py > gpurun_out/ncu2.log 2>&1; echo "full rc=$?"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --precision fp32 > gpurun_out/bench_fp32.json 2>/dev/null; echo "fp32 bench rc=$?"

mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "all gpu tests rc=$?"; tail -3 gpurun_out/t_all.log
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_final.err
timeout 300 python bench.py --precision fp32 --steps 10 --no-sweep --no-vit --no-cpu-baseline > gpurun_out/bench_final_fp32.json 2> gpurun_out/bench_final_fp32.err; echo "fp32 rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; echo "ref rc=$?"

# round 2, call A: full GPU suite on the tree as it stands (full-size parity tests, fused MSE, single-launch radix
# passes, NaN semantics), bench line, kernel bandwidth table.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider --timeout 600 > gpurun_out/r02a_tests.log 2>&1
echo "gpu tests rc=$?"; tail -15 gpurun_out/r02a_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/r02a_bench.json
timeout 300 python tools/bench_kernels.py --reps 10 > gpurun_out/r02a_kernels.json 2> gpurun_out/r02a_kernels.err
echo "bench_kernels rc=$?"; tail -c 3000 gpurun_out/r02a_kernels.json

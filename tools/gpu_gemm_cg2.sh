mkdir -p gpurun_out
rm -f gpurun_out/rc.txt
timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -k "gemm" -p no:cacheprovider > gpurun_out/t_gemm.log 2>&1; echo "gemm rc=$?" >> gpurun_out/rc.txt
tail -5 gpurun_out/t_gemm.log
timeout 200 python tools/bench_gemm.py > gpurun_out/bench_gemm_cg2.json 2> gpurun_out/bench_gemm_cg2.err; echo "bench cg2 rc=$?" >> gpurun_out/rc.txt
HBA_GEMM_CTA_GROUP=1 timeout 200 python tools/bench_gemm.py > gpurun_out/bench_gemm_cg1.json 2> gpurun_out/bench_gemm_cg1.err; echo "bench cg1 rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/bench_gemm_cg2.json gpurun_out/bench_gemm_cg1.json
timeout 600 python -m pytest tests/test_gpu_vit.py -q -p no:cacheprovider > gpurun_out/t_vit.log 2>&1; echo "vit rc=$?" >> gpurun_out/rc.txt
tail -40 gpurun_out/t_vit.log
cat gpurun_out/rc.txt

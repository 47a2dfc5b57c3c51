mkdir -p gpurun_out
S="qkv,c_fc+quickgelu,c_proj+res,c_proj_noepi,square8k_bf16out,out_proj+res"
timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_vit.py -q -x -k "gemm" -p no:cacheprovider > gpurun_out/t_gemm.log 2>&1; echo "gemm rc=$?"
tail -3 gpurun_out/t_gemm.log
timeout 300 python tools/bench_gemm.py --seconds 1.0 > gpurun_out/g_cg2.json 2> gpurun_out/g_cg2.err
HBA_GEMM_CTA_GROUP=1 timeout 300 python tools/bench_gemm.py --seconds 0.5 --no-cublas --only $S > gpurun_out/g_cg1.json 2>> gpurun_out/g_cg2.err
tail -3 gpurun_out/g_cg2.err
timeout 600 python -m pytest tests/test_gpu_vit.py -q -k "not gemm" -p no:cacheprovider > gpurun_out/t_vit.log 2>&1; echo "vit rc=$?"
tail -30 gpurun_out/t_vit.log

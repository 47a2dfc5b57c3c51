mkdir -p gpurun_out
timeout 120 python tools/trace_attn.py clip > gpurun_out/trace_fwd.log 2>&1; echo "rc=$?"; cat gpurun_out/trace_fwd.log

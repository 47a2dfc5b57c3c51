mkdir -p gpurun_out
timeout 300 python tools/profile_sweep_phases.py prelude > gpurun_out/sweep_phases2.log 2>&1; echo "rc=$?"; grep -v "^{" gpurun_out/sweep_phases2.log | tail -12

mkdir -p gpurun_out
nproc
OMP_NUM_THREADS=1 timeout 300 python tools/profile_sweep_phases.py > gpurun_out/sweep_phases_omp1.log 2>&1; echo "rc=$?"; grep -v "^{" gpurun_out/sweep_phases_omp1.log | tail -8
timeout 300 python tools/profile_sweep_phases.py > gpurun_out/sweep_phases_ompN.log 2>&1; echo "rc=$?"; grep -v "^{" gpurun_out/sweep_phases_ompN.log | tail -8

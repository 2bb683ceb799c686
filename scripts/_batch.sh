set -x
mkdir -p gpurun_out/r2f
# one ncu --set full capture per hot kernel of the current build: C3 8 shots and C2 30 shots (skip the warm-up launches)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bwd_step_kernel" -s 6 -c 2 -o gpurun_out/r2f/c3_bwd -f python scripts/ncu_target.py 8 8 c3 > gpurun_out/r2f/ncu_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bwd_step_kernel" -s 6 -c 2 -o gpurun_out/r2f/c2_bwd -f python scripts/ncu_target.py 8 30 c2 > gpurun_out/r2f/ncu_c2.log 2>&1
tail -3 gpurun_out/r2f/ncu_c3.log gpurun_out/r2f/ncu_c2.log
ls -la gpurun_out/r2f

set -x
mkdir -p gpurun_out/r2k
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rev_image_kernel|adj_step_kernel" -s 8 -c 4 -o gpurun_out/r2k/c2_bwd -f python scripts/ncu_target.py 24 30 c2 > gpurun_out/r2k/ncu_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rev_image_kernel|adj_step_kernel" -s 8 -c 4 -o gpurun_out/r2k/c3_bwd -f python scripts/ncu_target.py 24 8 c3 > gpurun_out/r2k/ncu_c3.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 260 --csv --log-file gpurun_out/r2k/launches_c2.csv python scripts/ncu_target.py 40 30 c2 > gpurun_out/r2k/launches.log 2>&1
wc -l gpurun_out/r2k/launches_c2.csv
timeout 600 python scripts/run_config.py c3 200 1000 > gpurun_out/r2k/c3_200.json 2> gpurun_out/r2k/c3_200.err; cat gpurun_out/r2k/c3_200.json; tail -3 gpurun_out/r2k/c3_200.err
ls -la gpurun_out/r2k

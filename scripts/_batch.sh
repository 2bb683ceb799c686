set -x
mkdir -p gpurun_out/r2n
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2n/tests.log 2>&1; tail -6 gpurun_out/r2n/tests.log
timeout 300 python scripts/kernel_times.py c2 30 200 > gpurun_out/r2n/kt_c2.log 2>&1; grep -v "^obs\|^grad" gpurun_out/r2n/kt_c2.log
timeout 300 python scripts/kernel_times.py c3 8 50 > gpurun_out/r2n/kt_c3.log 2>&1; grep -v "^obs\|^grad" gpurun_out/r2n/kt_c3.log
timeout 300 python scripts/run_config.py c2 30 2000 >> gpurun_out/r2n/cfg.json 2>> gpurun_out/r2n/cfg.err
timeout 600 python scripts/run_config.py c3 25 4000 >> gpurun_out/r2n/cfg.json 2>> gpurun_out/r2n/cfg.err
cat gpurun_out/r2n/cfg.json; tail -3 gpurun_out/r2n/cfg.err

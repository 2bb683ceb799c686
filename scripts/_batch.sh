set -x
mkdir -p gpurun_out/r2i
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2i/tests.log 2>&1; tail -8 gpurun_out/r2i/tests.log
timeout 300 python scripts/kernel_times.py c2 30 200 > gpurun_out/r2i/kt_c2.log 2>&1; grep -v "^obs\|^grad" gpurun_out/r2i/kt_c2.log
timeout 300 python scripts/kernel_times.py c3 8 50 > gpurun_out/r2i/kt_c3.log 2>&1; grep -v "^obs\|^grad" gpurun_out/r2i/kt_c3.log
for r in 5 2; do FWI_RING=$r timeout 300 python scripts/run_config.py c2 30 2000 >> gpurun_out/r2i/cfg.json 2>> gpurun_out/r2i/cfg.err; done
FWI_RING=2 timeout 600 python scripts/run_config.py c3 25 4000 >> gpurun_out/r2i/cfg.json 2>> gpurun_out/r2i/cfg.err
FWI_RING=2 timeout 900 python scripts/run_config.py c5 8 8000 >> gpurun_out/r2i/cfg.json 2>> gpurun_out/r2i/cfg.err
cat gpurun_out/r2i/cfg.json; tail -5 gpurun_out/r2i/cfg.err

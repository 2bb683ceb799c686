set -x
mkdir -p gpurun_out/r2m
( time python -c "import __graft_entry__ as g; g.build(); g.smoke()" ) > gpurun_out/r2m/smoke.log 2>&1; tail -4 gpurun_out/r2m/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2m/tests.log 2>&1; tail -4 gpurun_out/r2m/tests.log
python scripts/c1_times.py > gpurun_out/r2m/c1.json 2> gpurun_out/r2m/c1.err; cat gpurun_out/r2m/c1.json
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2m/bench_ref.json 2> gpurun_out/r2m/bench_ref.err
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2m/bench.json 2> gpurun_out/r2m/bench.err
grep real gpurun_out/r2m/*.err; cut -c1-260 gpurun_out/r2m/bench.json gpurun_out/r2m/bench_ref.json

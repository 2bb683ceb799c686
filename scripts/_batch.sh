set -x
mkdir -p gpurun_out/r2h
nvidia-smi -L
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -k "multi or merged or sharded" > gpurun_out/r2h/multi_test.log 2>&1; tail -15 gpurun_out/r2h/multi_test.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 ) > gpurun_out/r2h/bench2.json 2> gpurun_out/r2h/bench2.err
tail -12 gpurun_out/r2h/bench2.err; cut -c1-300 gpurun_out/r2h/bench2.json
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 ) > gpurun_out/r2h/ref2.json 2> gpurun_out/r2h/ref2.err
tail -6 gpurun_out/r2h/ref2.err; cut -c1-300 gpurun_out/r2h/ref2.json

set -x
mkdir -p gpurun_out/r2l
nvidia-smi -L | wc -l
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 ) > gpurun_out/r2l/bench8.json 2> gpurun_out/r2l/bench8.err
tail -8 gpurun_out/r2l/bench8.err; cut -c1-300 gpurun_out/r2l/bench8.json

#!/bin/bash
# bench.py under torchrun on N GPUs of this box (N = $1), strong scaling
n=${1:-8}; TAG=${TAG:-r2m}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}_bench_${n}gpu.json 2> gpurun_out/${TAG}_bench_${n}gpu.err
tail -3 gpurun_out/${TAG}_bench_${n}gpu.err; cut -c1-250 gpurun_out/${TAG}_bench_${n}gpu.json

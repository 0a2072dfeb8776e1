#!/bin/bash
# bench.py on the visible GPUs: N=1 always, then N=2.. if more are visible
set -x
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
NG=$(python -c "import torch; print(torch.cuda.device_count())")
for n in ${GPUS_LIST:-2 4 8}; do
  if [ "$n" -le "$NG" ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; tail -3 gpurun_out/bench_n$n.err; cat gpurun_out/bench_n$n.json
  fi
done

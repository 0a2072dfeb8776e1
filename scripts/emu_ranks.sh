# Emulated ranks of the 8-way strong-scaling step on one GPU (graph replay per step) + launch list.
set -x
TAG=${TAG:-emu}
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_multi.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/${TAG}_tests.txt
cat gpurun_out/${TAG}_tests.txt
J=${STRATA:-profiles/r2m_bench_8gpu.json}
python scripts/emulate_rank.py 0 1 1e9 10 > gpurun_out/${TAG}_w1.json 2> gpurun_out/${TAG}_w1.err
rm -f gpurun_out/${TAG}_w8.json
for r in 0 1 3 4 7; do python scripts/emulate_rank.py $r 8 1e9 20 $J >> gpurun_out/${TAG}_w8.json 2>> gpurun_out/${TAG}_w8.err; done
cat gpurun_out/${TAG}_w1.json gpurun_out/${TAG}_w8.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_r3_launches.csv python scripts/emulate_rank.py 3 8 1e9 3 $J > gpurun_out/${TAG}_ncu.log 2>&1

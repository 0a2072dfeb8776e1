set -x
python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "hierarchical or guide or build or sharded" 2>&1 | tail -3
python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3
J=profiles/r2i_bench_8gpu.json
python scripts/emulate_rank.py 0 1 1e9 10 > gpurun_out/emu_w1.json 2> gpurun_out/emu_w1.err
for r in 0 1 3 4 7; do python scripts/emulate_rank.py $r 8 1e9 20 $J >> gpurun_out/emu_w8.json 2>> gpurun_out/emu_w8.err; done
cat gpurun_out/emu_w1.json gpurun_out/emu_w8.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/emu_r3_launches.csv python scripts/emulate_rank.py 3 8 1e9 3 $J > gpurun_out/emu_ncu.log 2>&1


# Round-end check on one GPU: full GPU suite, smoke, default bench, the other configs, launch list.
set -x
TAG=${TAG:-r2m}
python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/${TAG}_pytest.txt; cat gpurun_out/${TAG}_pytest.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench_1gpu.err; tail -2 gpurun_out/${TAG}_bench_1gpu.err; cut -c1-400 gpurun_out/${TAG}_bench_1gpu.json
for c in C4 C5; do python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err; cut -c1-300 gpurun_out/${TAG}_bench_$c.json; done
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-other > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-other > gpurun_out/${TAG}_ncu.log 2>&1

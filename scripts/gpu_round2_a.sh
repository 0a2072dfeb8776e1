#!/bin/bash
# first GPU check of the two-emitter cell-major path: parity tests, bench, launch list
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "cellmajor or cell_order or sample_to_host or full_size_draw" > gpurun_out/a_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/a_tests.log
tail -5 gpurun_out/a_tests.log
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-other > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
cat gpurun_out/a_bench.json
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-other > gpurun_out/a_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/a_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-other > gpurun_out/a_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/a_launches.csv') if l.startswith('"')))
h = rows[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value")
d = collections.defaultdict(list)
for r in rows[1:]:
    d[r[ki][:60]].append(float(r[vi].replace(',', '')))
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:60s} n={len(v):3d} mean={sum(v)/len(v)/1e3:9.1f} us  last={v[-1]/1e3:9.1f}")
PY

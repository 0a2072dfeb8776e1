# tuning: how throughput depends on table footprint (cells per dim)
run() { python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu "$@" 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'])"; }
for n in 32 64 128 192 256; do echo "grid_n=$n"; run --grid-n $n; done

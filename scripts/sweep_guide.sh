python -m pytest tests -m gpu -q 2>&1 | tail -5
for gb in 20 22 24; do echo "guide_bits=$gb"; python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --guide-bits $gb 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'])"; done
echo "no records gb=24"; python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-records 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'])"

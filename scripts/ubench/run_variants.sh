python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-other 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['stages_ms'], d['roofline']['step_frac_of_hbm_roofline'])"
for r in 0 3; do
  python scripts/emulate_rank.py $r 8 1e9 3 2>&1 | tail -1
  LINT_SIDE=0 python scripts/emulate_rank.py $r 8 1e9 3 2>&1 | tail -1
done

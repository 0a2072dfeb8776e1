# emitter co-residency sweep on one emulated rank and on the single-GPU step
J=${STRATA:-profiles/r2m_bench_8gpu.json}
for rb in 3 4 6 8 12; do for fb in 2 3 4; do
  echo "== REST_BPS=$rb FLAT_BPS=$fb"
  LINT_REST_BPS=$rb LINT_FLAT_BPS=$fb python scripts/emulate_rank.py 0 1 1e9 10 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  w1', round(d['graph_ms_per_step'],4))"
  LINT_REST_BPS=$rb LINT_FLAT_BPS=$fb python scripts/emulate_rank.py 3 8 1e9 20 $J 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  r3', round(d['graph_ms_per_step'],4))"
  LINT_REST_BPS=$rb LINT_FLAT_BPS=$fb python scripts/emulate_rank.py 7 8 1e9 20 $J 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  r7', round(d['graph_ms_per_step'],4))"
done; done

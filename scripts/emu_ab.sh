# step time of emulated ranks (graph replay), strata of a real 8-GPU run; A/B of LINT_PDL
J=${STRATA:-profiles/r2m_bench_8gpu.json}
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3
for pdl in 1 0; do
  echo "== LINT_PDL=$pdl"
  LINT_PDL=$pdl timeout 120 python scripts/emulate_rank.py 0 1 1e9 10 2>&1 | tail -1
  for r in 0 3 7; do LINT_PDL=$pdl timeout 120 python scripts/emulate_rank.py $r 8 1e9 20 $J 2>&1 | tail -1; done
done

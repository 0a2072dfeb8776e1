# step time of emulated ranks (graph replay), strata of a real 8-GPU run
J=${STRATA:-profiles/r2m_bench_8gpu.json}
python scripts/emulate_rank.py 0 1 1e9 10 2>/dev/null
for r in 0 3 4 7; do python scripts/emulate_rank.py $r 8 1e9 20 $J 2>/dev/null; done

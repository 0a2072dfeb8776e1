# A/B of step variants on emulated ranks (graph replay)
J=profiles/r2i_bench_8gpu.json
for cfg in "GUIDE_BITS=0" "GUIDE_BITS=16"; do
  echo "== $cfg"
  env $cfg python scripts/emulate_rank.py 0 1 1e9 10 2>/dev/null
  env $cfg python scripts/emulate_rank.py 0 8 1e9 20 $J 2>/dev/null
  env $cfg python scripts/emulate_rank.py 3 8 1e9 20 $J 2>/dev/null
  env $cfg python scripts/emulate_rank.py 7 8 1e9 20 $J 2>/dev/null
done

J=${STRATA:-profiles/r2m_bench_8gpu.json}
for lib in liblintsampler_b200.so variant_minb6.so variant_minb8.so; do
  echo "== $lib"
  export LINTSAMPLER_B200_LIB=$PWD/lintsampler_b200/_lib/$lib
  python scripts/emulate_rank.py 0 1 1e9 10 2>/dev/null
  python scripts/emulate_rank.py 3 8 1e9 20 $J 2>/dev/null
  python scripts/emulate_rank.py 7 8 1e9 20 $J 2>/dev/null
done

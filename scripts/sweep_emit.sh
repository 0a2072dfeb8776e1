# tuning sweep of the emit kernel's compile-time knobs (rebuilds on the GPU box)
run() { python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-other 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'])"; }
for cfg in "-DLINT_EMIT_U=4 -DLINT_EMIT_MINB=4" "-DLINT_EMIT_U=4 -DLINT_EMIT_MINB=3" "-DLINT_EMIT_U=4 -DLINT_EMIT_MINB=5" "-DLINT_EMIT_U=8 -DLINT_EMIT_MINB=3" "-DLINT_EMIT_U=2 -DLINT_EMIT_MINB=5" "-DLINT_EMIT_U=4 -DLINT_EMIT_MINB=8 -DLINT_EMIT_THREADS=128"; do
  echo "== $cfg"; LINT_EXTRA_NVCC="$cfg" python -m lintsampler_b200._build --force > /dev/null 2>&1 && run
done

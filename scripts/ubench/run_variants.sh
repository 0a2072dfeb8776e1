for v in "" MINB3; do
  if [ -n "$v" ]; then export LINTSAMPLER_B200_LIB=$PWD/scripts/ubench/lib_$v.so; fi
  python scripts/time_emit.py 1e9 3 > gpurun_out/l_plain_$v.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l_launches_$v.csv python scripts/time_emit.py 1e9 1 > gpurun_out/l_ncu_$v.log 2>&1
  echo "== variant [$v]"; head -1 gpurun_out/l_plain_$v.log; python scripts/launch_summary.py gpurun_out/l_launches_$v.csv > gpurun_out/l_sum_$v.txt; grep -E "emit|boxes" gpurun_out/l_sum_$v.txt
done
python -m pytest tests -m gpu -q -x -k "cellmajor_structure and g3d" 2>&1 | tail -3

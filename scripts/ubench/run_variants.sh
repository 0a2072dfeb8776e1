python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for r in 0 1 2 3 4 5 6 7; do
  python scripts/emulate_rank.py $r 8 1e9 3 2>&1 | tail -1
done
python scripts/emulate_rank.py 0 8 1e9 3 > gpurun_out/v_plain_0.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/v_launches_0.csv python scripts/emulate_rank.py 0 8 1e9 1 > gpurun_out/v_ncu_0.log 2>&1
python scripts/launch_summary.py gpurun_out/v_launches_0.csv > gpurun_out/v_sum_0.txt; head -14 gpurun_out/v_sum_0.txt

set -x
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-other --no-graph > gpurun_out/r2h_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-other --no-graph > gpurun_out/r2h_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/r2h_launches.csv > gpurun_out/r2h_launch_summary.txt; cat gpurun_out/r2h_launch_summary.txt
python scripts/time_emit.py 1e9 1 > gpurun_out/r2h_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"emit_flat|emit_kernel|poisson|group_scan|vertex_rowsum" -c 5 -o gpurun_out/r2h_prof python scripts/time_emit.py 1e9 1 > gpurun_out/r2h_ncu2.log 2>&1
tail -2 gpurun_out/r2h_ncu2.log

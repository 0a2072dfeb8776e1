python -m pytest tests -m gpu -q -x -k "(cellmajor_structure and g3d) or (flat_residual and 3-) or cell_order or sample_to_host" 2>&1 | tail -3
python scripts/time_emit.py 1e9 3 > gpurun_out/m_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/m_launches.csv python scripts/time_emit.py 1e9 1 > gpurun_out/m_ncu.log 2>&1
head -1 gpurun_out/m_plain.log; python scripts/launch_summary.py gpurun_out/m_launches.csv > gpurun_out/m_sum.txt; grep -E "emit|boxes|poisson|scan_compact|chunk_start|sample_kernel" gpurun_out/m_sum.txt

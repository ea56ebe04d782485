mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02m_launches_c5.csv python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02m_ncu_launch_c5.log 2>&1
tail -2 gpurun_out/r02m_ncu_launch_c5.log | cut -c1-200

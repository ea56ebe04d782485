mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel --launch-skip 68 --launch-count 8 -o gpurun_out/r02n_prof_c5 python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02n_ncu_full_c5.log 2>&1
tail -3 gpurun_out/r02n_ncu_full_c5.log | cut -c1-200
ls -la gpurun_out/r02n_prof_c5.ncu-rep

mkdir -p gpurun_out
python -m pytest tests/test_gpu.py -m gpu -q -k "c5_shape or wide" -s 2>&1 | grep -E "C5 shape|passed|failed" | cut -c1-330
python bench.py --workload c5 --steps 10 --no-cpu-baseline > gpurun_out/r02e_bench_c5.json 2> gpurun_out/r02e_bench_c5.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02e_bench_c5.json').read().strip().splitlines()[-1])
print('c5', round(d['value']/1e6,2),'M frames/s', round(d['ms_per_step'],3),'ms', {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})
PY
ncu --set full --clock-control none --import-source on -k regex:"pass1_kernel|pass2_kernel" --launch-skip 6 --launch-count 2 -o gpurun_out/r02e_prof_c3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02e_ncu_full_c3.log 2>&1; tail -3 gpurun_out/r02e_ncu_full_c3.log

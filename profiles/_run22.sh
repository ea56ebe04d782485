mkdir -p gpurun_out
python -m pytest tests/test_gpu.py -q -x -k "eigen" > gpurun_out/r02y_pytest.log 2>&1; tail -2 gpurun_out/r02y_pytest.log
for wl in c3 c4; do python bench.py --workload $wl --steps 10 --no-cpu-baseline > gpurun_out/r02y_bench_$wl.json 2> gpurun_out/r02y_bench_$wl.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02y_bench_$wl.json').read().strip().splitlines()[-1])
print('$wl', round(d['value']/1e6,2),'M frames/s', round(d['ms_per_step'],3),'ms', {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})
PY
done
ncu --set full --clock-control none --import-source on -k regex:"pass2_kernel" --launch-skip 3 --launch-count 1 -o gpurun_out/r02y_prof_c3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02y_ncu_full_c3.log 2>&1
python profiles/ncu_summary.py gpurun_out/r02y_prof_c3.ncu-rep > gpurun_out/r02y_ncu_c3_summary.txt 2>&1; grep -E 'duration|dram__bytes' gpurun_out/r02y_ncu_c3_summary.txt

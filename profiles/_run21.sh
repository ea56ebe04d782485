mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r02r_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02r_pytest_gpu.log
for wl in c2 c3; do python bench.py --workload $wl --steps 10 --no-cpu-baseline > gpurun_out/r02r_bench_$wl.json 2> gpurun_out/r02r_bench_$wl.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02r_bench_$wl.json').read().strip().splitlines()[-1])
print('$wl', round(d['value']/1e6,2),'M frames/s', round(d['ms_per_step'],3),'ms', {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})
PY
done

mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r02d_pytest_gpu.log 2>&1; tail -5 gpurun_out/r02d_pytest_gpu.log
python -m pytest tests/test_gpu.py -m gpu -q -k "c5_shape" -s 2>&1 | grep "C5 shape" | cut -c1-170
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02d_bench_c3.json 2> gpurun_out/r02d_bench_c3.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02d_bench_c3.json').read().strip().splitlines()[-1])
print('C3', round(d['value']/1e6,1),'M frames/s', round(d['ms_per_step'],3),'ms e2e', round(d['e2e']['value']/1e6,1), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})
PY
for wl in c4 c1 c5; do python bench.py --workload $wl --steps 10 --no-cpu-baseline > gpurun_out/r02d_bench_$wl.json 2> gpurun_out/r02d_bench_$wl.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02d_bench_$wl.json').read().strip().splitlines()[-1])
print('$wl', round(d['value']/1e6,2),'M frames/s', round(d['ms_per_step'],3),'ms', {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})
PY
done

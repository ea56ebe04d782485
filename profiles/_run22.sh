mkdir -p gpurun_out
python -m pytest tests/test_gpu.py -q -x -k "align or eigen_fast_path or full_comparison or prepass" > gpurun_out/r02w_pytest.log 2>&1; tail -3 gpurun_out/r02w_pytest.log
for wl in c3 c2p c4; do python bench.py --workload $wl --steps 10 --no-cpu-baseline > gpurun_out/r02w_bench_$wl.json 2> gpurun_out/r02w_bench_$wl.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02w_bench_$wl.json').read().strip().splitlines()[-1])
print('$wl', round(d['value']/1e6,2),'M frames/s', round(d['ms_per_step'],3),'ms', d['roofline']['frac'], {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})
PY
done

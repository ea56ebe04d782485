mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu.py -m gpu -q -x -k "ae_wide or ae_c5" > gpurun_out/r02s_pytest_ae.log 2>&1; tail -2 gpurun_out/r02s_pytest_ae.log | cut -c1-300
timeout 600 python bench.py --workload c5 --steps 10 --no-cpu-baseline > gpurun_out/r02s_bench_c5.json 2> gpurun_out/r02s_bench_c5.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02s_bench_c5.json').read().strip().splitlines()[-1])
print('c5', round(d['value']/1e6,3),'M frames/s', round(d['ms_per_step'],3),'ms', {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()}, d['roofline']['frac'])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_gemm|tile_image" --launch-skip 130 --launch-count 44 --csv --log-file gpurun_out/r02s_launches_c5.csv python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02s_ncu_launch_c5.log 2>&1
python - <<PY
import csv
rows=list(csv.reader(l for l in open('gpurun_out/r02s_launches_c5.csv') if l.startswith('"')))
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); gi=hdr.index("Grid Size")
print(' '.join(f"{r[ki][11:15]}{r[gi].split(',')[0][1:]}:{float(r[vi])/1e3:.0f}" for r in rows[1:]))
PY

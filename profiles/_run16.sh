mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu.py -m gpu -q -x -k "ae_wide or ae_c5 or ae_loss or ae_prepass" -s > gpurun_out/r02m_pytest_ae.log 2>&1; tail -12 gpurun_out/r02m_pytest_ae.log | cut -c1-300
timeout 600 python bench.py --workload c5 --steps 10 --no-cpu-baseline > gpurun_out/r02m_bench_c5.json 2> gpurun_out/r02m_bench_c5.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02m_bench_c5.json').read().strip().splitlines()[-1])
print('c5', round(d['value']/1e6,3),'M frames/s', round(d['ms_per_step'],3),'ms', {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})
PY

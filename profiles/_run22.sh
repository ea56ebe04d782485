mkdir -p gpurun_out
python -m pytest tests/test_gpu.py -q -x -k "eigen" > gpurun_out/r02u_pytest.log 2>&1; tail -4 gpurun_out/r02u_pytest.log
for wl in c3 c4 c1; do python bench.py --workload $wl --steps 10 --no-cpu-baseline > gpurun_out/r02u_bench_$wl.json 2> gpurun_out/r02u_bench_$wl.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02u_bench_$wl.json').read().strip().splitlines()[-1])
print('$wl', round(d['value']/1e6,2),'M frames/s', round(d['ms_per_step'],3),'ms', {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})
PY
done
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1; lscpu | grep -i -E 'numa|socket|model name|^CPU\(s\)' >> gpurun_out/r02_topo.txt; cat /sys/bus/pci/devices/*/numa_node 2>/dev/null | sort | uniq -c >> gpurun_out/r02_topo.txt

mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps $1 --warmup 5 --no-cpu-baseline > gpurun_out/r02f_2gpu_$2.json 2> gpurun_out/r02f_2gpu_$2.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02f_2gpu_$2.json').read().strip().splitlines()[-1])
    print('$2', round(d['value']/1e6,1),'M', round(d['ms_per_step'],3),'ms e2e', round(d['e2e']['value']/1e6,1), d['e2e'].get('host_memory'), 'strong', d.get('scaling_strong'), 'dp', d.get('dp_parity'), 'ranks', d.get('kernel_ms_per_step_by_rank'))
except Exception as e:
    print('$2 failed', e); print(open('gpurun_out/r02f_2gpu_$2.err').read()[-1500:])
PY
}
run 20 k20
CVF_BENCH_CLOCKS=0 run 20 k20_noclk
run 100 k100
python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -3

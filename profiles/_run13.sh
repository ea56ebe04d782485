mkdir -p gpurun_out
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_c3_${N}gpu.json 2> gpurun_out/r02_bench_c3_${N}gpu.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_c3_${N}gpu.json').read().strip().splitlines()[-1])
    print('N=$N', round(d['value']/1e6,1),'M', round(d['ms_per_step'],3),'ms e2e', round(d['e2e']['value']/1e6,1), d['e2e'].get('host_memory'), '\n strong', d.get('scaling_strong'), '\n dp', d.get('dp_parity'), '\n ranks', d.get('kernel_ms_per_step_by_rank'), d['clocks'])
except Exception as e:
    print('failed', e); print(open('gpurun_out/r02_bench_c3_${N}gpu.err').read()[-3000:])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | cut -c1-200

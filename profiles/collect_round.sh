#!/bin/bash
# Round artefacts in two GPU calls (the .ncu-rep files of both would exceed gpurun's 64 MiB pull limit):
#   gpurun --timeout 1800 -- 'bash profiles/collect_round.sh r01 a'     tests, benches, launch lists, full capture of C3
#   gpurun --timeout 900  -- 'bash profiles/collect_round.sh r01 b'     full captures of C5 (tcgen05 products) and C4 (feature kernels)
# Everything lands in gpurun_out/<tag>_*; numbers printed under ncu are never used as bench values.
tag=${1:-r01}
part=${2:-a}
out=gpurun_out
mkdir -p $out
if [ "$part" = a ]; then
python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1; tail -2 $out/${tag}_pytest_gpu.log
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_c3_reference_arm.json 2>/dev/null
python bench.py > $out/${tag}_bench_c3.json 2> $out/${tag}_bench_c3.err
for wl in c1 c2 c2p c4 c5; do
  python bench.py --workload $wl --steps 10 > $out/${tag}_bench_$wl.json 2> $out/${tag}_bench_$wl.err
done
python -c "
import json,glob
for f in sorted(glob.glob('$out/${tag}_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], round(d['value']/1e6,2),'M frames/s', round(d.get('ms_per_step',0),3),'ms', 'e2e', round(d['e2e']['value']/1e6,2), d.get('roofline',{}).get('frac'))
    except Exception as e: print(f, 'unreadable', e)
"
# launch lists (whole command) and full captures (dominant kernels) of the same commands
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_launch_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"pass1_kernel|pass2_kernel|dw1_kernel|prep_align_kernel" --launch-skip 12 --launch-count 4 -o $out/${tag}_prof_c3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_full_c3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_c5.csv python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_launch_c5.log 2>&1
else
ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_kernel" --launch-skip 66 --launch-count 9 -o $out/${tag}_prof_c5 python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_full_c5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"jjt_kernel|prep_feat_kernel" --launch-skip 6 --launch-count 2 -o $out/${tag}_prof_c4 python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_full_c4.log 2>&1
fi
ls -la $out | grep ${tag}_ | tail -30

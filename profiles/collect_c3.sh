#!/bin/bash
# C3-only refresh of the round artefacts after a change to the eigenfunction kernels (the full set: collect_round.sh):
#   gpurun --timeout 900 -- 'bash profiles/collect_c3.sh r02'
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1; tail -2 $out/${tag}_pytest_gpu.log
python bench.py --strong > $out/${tag}_bench_c3.json 2> $out/${tag}_bench_c3.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_launch_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"pass1_kernel|pass2_kernel|prep_align_kernel|stats_kernel" --launch-skip 12 --launch-count 4 -o $out/${tag}_prof_c3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_full_c3.log 2>&1
hash=$(python -c "import __graft_entry__ as g; print(g.library_hash())")
python profiles/ncu_traffic.py $out/${tag}_prof_c3.ncu-rep --workload c3 --frames 4194304 --source-hash $hash --out $out/ncu_traffic.json > /dev/null
python profiles/ncu_summary.py $out/${tag}_prof_c3.ncu-rep > $out/${tag}_ncu_c3_summary.txt 2>&1
cp $out/ncu_traffic.json profiles/ncu_traffic.json
python bench.py --steps 10 --no-cpu-baseline > $out/${tag}_bench_c3_with_traffic.json 2>/dev/null
python -c "
import json
for f in ('$out/${tag}_bench_c3.json','$out/${tag}_bench_c3_with_traffic.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']/1e6,2), round(d['ms_per_step'],3), d['roofline']['traffic'], d['roofline'].get('traffic_source'), d['native'])
"

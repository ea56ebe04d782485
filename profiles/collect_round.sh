#!/bin/bash
# Round artefacts in one GPU call:
#   gpurun --timeout 2400 -- 'bash profiles/collect_round.sh r02'
# tests, benches of every workload, reference arm, parity report, launch list and full ncu capture of C3 (+ the traffic table
# bench.py reads), small-batch timings, tensor-core probe.  Everything lands in gpurun_out/<tag>_*; numbers printed under ncu are
# never used as bench values.
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1; tail -2 $out/${tag}_pytest_gpu.log
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_c3_reference_arm.json 2>/dev/null
python bench.py --strong > $out/${tag}_bench_c3.json 2> $out/${tag}_bench_c3.err
for wl in c1 c2 c2p c4 c5; do
  python bench.py --workload $wl --steps 10 > $out/${tag}_bench_$wl.json 2> $out/${tag}_bench_$wl.err
done
python -c "
import json,glob
for f in sorted(glob.glob('$out/${tag}_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], round(d['value']/1e6,2),'M frames/s', round(d.get('ms_per_step',0),3),'ms', 'e2e', round(d['e2e']['value']/1e6,2), d.get('roofline',{}).get('frac'), {k:round(v['ms_per_step'],3) for k,v in d.get('kernels',{}).items()})
    except Exception as e: print(f, 'unreadable', e)
"
python profiles/parity_report.py --out $out/${tag}_parity_errors.json > $out/${tag}_parity_report.log 2>&1; tail -3 $out/${tag}_parity_report.log
python profiles/small_batch_train.py > $out/${tag}_small_batch_train.txt 2>&1; grep -E "graph|eager" $out/${tag}_small_batch_train.txt | tail -6
timeout 120 profiles/_build/tc_small_probe 2000 > $out/${tag}_tc_small_probe.json 2>&1
# launch list (whole command) and full capture (dominant kernels) of the same command
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_launch_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"pass1_kernel|pass2_kernel|prep_align_kernel|stats_kernel" --launch-skip 12 --launch-count 4 -o $out/${tag}_prof_c3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_full_c3.log 2>&1
hash=$(python -c "import __graft_entry__ as g; print(g.library_hash())")
python profiles/ncu_traffic.py $out/${tag}_prof_c3.ncu-rep --workload c3 --frames 4194304 --source-hash $hash --out $out/ncu_traffic.json > /dev/null
python profiles/ncu_summary.py $out/${tag}_prof_c3.ncu-rep > $out/${tag}_ncu_c3_summary.txt 2>&1
# C5: launch list of one step and a full capture of the forward products + the first two backward products of a chunk
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_gemm|tile_image|loss_delta|add_sums|accumulate_grad|stage_input" --launch-skip 150 --launch-count 60 --csv --log-file $out/${tag}_launches_c5.csv python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_launch_c5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel --launch-skip 68 --launch-count 8 -o $out/${tag}_prof_c5 python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_full_c5.log 2>&1
python profiles/ncu_summary.py $out/${tag}_prof_c5.ncu-rep > $out/${tag}_ncu_c5_summary.txt 2>&1
ls -la $out | grep ${tag}_ | tail -40

mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest_gpu.log 2>&1; tail -15 gpurun_out/r02a_pytest_gpu.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench_c3.json 2> gpurun_out/r02a_bench_c3.err; tail -c 1500 gpurun_out/r02a_bench_c3.json
python profiles/parity_report.py --frames 1048576 --out gpurun_out/r02_parity_errors.json > gpurun_out/r02_parity_report.log 2>&1; tail -8 gpurun_out/r02_parity_report.log
compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_memcheck.log 2>&1; tail -4 gpurun_out/r02_sanitizer_memcheck.log
compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_racecheck.log 2>&1; tail -4 gpurun_out/r02_sanitizer_racecheck.log

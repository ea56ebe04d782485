mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02b_pytest_gpu.log 2>&1; tail -40 gpurun_out/r02b_pytest_gpu.log
grep -n "C5 shape" gpurun_out/r02b_pytest_gpu.log

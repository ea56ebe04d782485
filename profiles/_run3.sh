mkdir -p gpurun_out
python -m pytest tests/test_gpu.py tests/test_regae_gpu.py -m gpu -q -k "c5_shape or regae_train" -s 2>&1 | grep -E "C5 shape|passed|failed|Error" | head -30

mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_attention.py -x -q -s 2>&1 | grep -E "passed|failed|attention B=32|attention B=8|Error|timeout" | tail -8
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_interface.py -x -q 2>&1 | tail -3

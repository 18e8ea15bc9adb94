python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k audited 2>&1 | tail -15
ls gpurun_out

timeout 25 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "full_rk3_step and float64" 2>&1 | tail -2

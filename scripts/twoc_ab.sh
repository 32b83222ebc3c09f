timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_sizes.py tests/test_gpu_batched.py -q -m gpu -k "2c2e or two_center or metric or golden or df or DF" 2>&1 | tail -3
timeout 300 python scripts/twoc_bench.py 2>&1 | tail -4
SYMPLEINTS_B200_LIB=$PWD/sympleints_b200/lib/variants/lib_prev.so timeout 300 python scripts/twoc_bench.py 2>&1 | tail -4

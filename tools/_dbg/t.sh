timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "groupnorm or gn" 2>&1 | tail -2
timeout 300 python tools/kernel_bench.py 2>&1 | grep -i "^groupnorm"

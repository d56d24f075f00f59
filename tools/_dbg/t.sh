timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention" 2>&1 | tail -3
AV2V_ATTN_SPLIT=0 AV2V_LIB=tools/_dbg/libanyv2v_b200_timers.so timeout 300 python tools/kernel_bench.py 2>&1 | grep "^nv=1"
AV2V_LIB=tools/_dbg/libanyv2v_b200_timers.so timeout 300 python tools/kernel_bench.py 2>&1 | grep "^nv=1"

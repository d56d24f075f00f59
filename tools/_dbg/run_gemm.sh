AV2V_GEMM_MC2=0 timeout 200 python tools/pair_probe.py 2>&1 | grep "mode=0" | cut -c1-80
timeout 200 python tools/gemm_epi_probe.py 2>&1 | grep dbg | cut -c1-70
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -3

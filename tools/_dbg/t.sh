AV2V_LIB=tools/_dbg/libanyv2v_b200_timers.so timeout 400 python tools/gemm_role_timers.py 2>&1 | tail -50

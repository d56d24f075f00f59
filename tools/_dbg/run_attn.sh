timeout 250 python tools/gpu_check.py attn 2>&1 | grep -E "ok=|time|SDPA|rror" | cut -c1-150
AV2V_LIB=tools/_dbg/libanyv2v_b200_timers.so timeout 200 python tools/attn_timer_probe.py 2>&1 | tail -8

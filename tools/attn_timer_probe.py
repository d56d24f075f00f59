"""Phase timers of one softmax warp pair of CTA 0 in the attention kernel (bring-up build with -DAV2V_ATTN_TIMERS).
build:  nvcc ... -DAV2V_ATTN_TIMERS -o tools/_dbg/libanyv2v_b200_timers.so   (see tools/build_dbg.sh)
run:    AV2V_LIB=tools/_dbg/libanyv2v_b200_timers.so python tools/attn_timer_probe.py"""
import ctypes, os, sys, torch
sys.path.insert(0, ".")
from anyv2v_b200 import ops, _lib
from tools.gpu_check import timeit
dev = "cuda"
lib = _lib.lib()
lib.av2v_attn_debug_timers.argtypes = [ctypes.c_void_p]
names = ["wait_s_full", "tmem_ld", "max+publish", "exps+st", "collect", "st_wait+arrive", "item_prologue+epilogue", "total"]
torch.manual_seed(0)
for (batch, heads, seq, nv) in [(16, 5, 4096, 3)]:  # nv = 1 rows mode runs on attention2q_tcgen05.cu (no timers)
    C = heads * 64
    nb = 3 if nv == 3 else 1
    qkv = torch.randn(nb * batch * seq, 3 * C, device=dev).half()
    q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    out = torch.empty(nb * batch * seq, C, device=dev, dtype=torch.float16)
    rows = batch * seq
    if nv == 1:
        fn = lambda: ops.attention(q, k, v, heads, seq, batch, out)
    else:
        fn = lambda: ops.attention(q[:rows], k[:rows], v, heads, seq, batch, out, n_v=3, v_branch_stride=rows * 3 * C, o_branch_stride=rows * C)
    t = timeit(fn, iters=5)
    buf = (ctypes.c_ulonglong * 24)()
    lib.av2v_attn_debug_timers(buf)
    items = -(-batch * heads * (seq // 128) // 148)
    tiles = items * (seq // 128)
    mnames = ["wait_p_ready_A", "wait_p_ready_B", "wait_k_full", "wait_v_full/o_empty", "wait_q_full", "-", "total", "issue(everything else)"]
    print(f"nv={nv}: {t*1e6:.1f} us; CTA 0: ~{tiles} key tiles", flush=True)
    for half in range(2):
        vals = [buf[half * 8 + i] for i in range(8)]
        print(f"   softmax half {half}: " + " ".join(f"{n}={v/tiles:.0f}" for n, v in zip(names, vals)) + "  (cycles per key tile)", flush=True)
    print("   MMA thread:    " + " ".join(f"{n}={buf[16 + i]/tiles:.0f}" for i, n in enumerate(mnames) if n != "-") + "  (cycles per key tile)", flush=True)

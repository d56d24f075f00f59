"""Role timers / what-if switches of gemm_tcgen05_kernel on the short-K shapes (bring-up build only: tools/build_dbg.sh defines
AV2V_GEMM_BRINGUP, the shipped library has neither the switches nor the environment read).
  AV2V_LIB=tools/_dbg/libanyv2v_b200_timers.so python tools/gemm_role_timers.py
AV2V_GEMM_DEBUG bits: 8 role timers of CTA 0 | 16 epilogue only frees the accumulator | 32 W tiles not re-loaded | 64 no TMA stores | 128 force CTA pairs | 256 forbid CTA pairs |
512 generic epilogue flavour only | 1024 A tiles not re-loaded"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anyv2v_b200 import _lib, ops  # noqa: E402
from tools.gpu_check import timeit  # noqa: E402

dev = "cuda"
lib = _lib.lib()
names = ["prod_wait_empty", "prod_total", "mma_wait_tempty", "mma_wait_full", "mma_total"]
shapes = [(196608, 320, 640, True, False), (49152, 640, 1280, True, False), (49152, 1920, 640, False, False), (12288, 3840, 1280, False, False), (196608, 320, 1280, True, False),
          (49152, 640, 640, True, False), (49152, 5120, 640, False, True), (12288, 1280, 1280, True, False)]
for (M, N, K, res, geglu) in shapes:
    a = torch.randn(M, K, device=dev).half()
    w = (torch.randn(N, K, device=dev) / 18).half()
    b = torch.randn(N, device=dev).half()
    if geglu:
        w, b = ops.geglu_pack(w, b)
    r = torch.randn(M, N, device=dev).half() if res else None
    out = torch.empty(M, N // 2 if geglu else N, device=dev, dtype=torch.float16)
    for dbg in (0, 256, 128):
        os.environ["AV2V_GEMM_DEBUG"] = str(dbg)
        t = timeit(lambda: ops.linear(a, w, bias=b, residual=r, geglu=geglu, out=out), iters=10)
        line = f"M={M} N={N} K={K} res={int(res)} geglu={int(geglu)} dbg={dbg:3d}: {t * 1e6:7.1f} us"
        if dbg & 8:
            buf = (ctypes.c_ulonglong * 16)()
            lib.av2v_gemm_debug_timers.argtypes = [ctypes.c_void_p]
            lib.av2v_gemm_debug_timers(buf)
            line += " | " + " ".join(f"{n}={buf[i] / 1e3:.0f}k" for i, n in enumerate(names))
        print(line, flush=True)

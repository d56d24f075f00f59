import ctypes, os, sys, torch
sys.path.insert(0, ".")
from anyv2v_b200 import ops, _lib
from tools.gpu_check import timeit
dev = "cuda"
lib = _lib.lib()
names = ["prod_wait_empty", "prod_total", "mma_wait_tempty", "mma_wait_full", "mma_total", "epi_wait_tfull", "epi_tmem_ld", "epi_fence_bar", "epi_total", "epi_first_use", "epi_bias"]
for (M, N, K, res) in [(196608, 320, 320, False), (196608, 320, 320, True), (196608, 960, 320, False), (196608, 2560, 320, False), (49152, 5120, 640, False), (12288, 2560, 5120, False)]:
    a = torch.randn(M, K, device=dev).half(); w = torch.randn(N, K, device=dev).half(); b = torch.randn(N, device=dev).half()
    r = torch.randn(M, N, device=dev).half() if res else None
    out = torch.empty(M, N, device=dev, dtype=torch.float16)
    t = timeit(lambda: ops.linear(a, w, bias=b, residual=r, out=out), iters=10)
    buf = (ctypes.c_ulonglong * 16)()
    lib.av2v_gemm_debug_timers.argtypes = [ctypes.c_void_p]
    lib.av2v_gemm_debug_timers(buf)
    tim = {n: buf[i] for i, n in enumerate(names)}
    print(f"dbg={os.environ.get('AV2V_GEMM_DEBUG','0')} M={M} N={N} K={K} res={res}: {t*1e6:.1f} us {2*M*N*K/t/1e12:.0f} TF | " +
          " ".join(f"{k}={v/1e3:.0f}k" for k, v in tim.items()), flush=True)

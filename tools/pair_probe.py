"""cta_group::2 pair mode (AV2V_GEMM_MC2=2) vs independent CTAs: timing + role timers of CTA 0 (AV2V_GEMM_DEBUG=8)."""
import ctypes, os, sys, torch
sys.path.insert(0, ".")
from anyv2v_b200 import ops, _lib
from tools.gpu_check import timeit
dev = "cuda"
lib = _lib.lib()
lib.av2v_gemm_debug_timers.argtypes = [ctypes.c_void_p]
names = ["prod_wait_empty", "prod_total", "mma_wait_tempty", "mma_wait_full", "mma_total"]
torch.manual_seed(0)

def timers():
    buf = (ctypes.c_ulonglong * 16)()
    lib.av2v_gemm_debug_timers(buf)
    return " ".join(f"{n}={buf[i]/1e3:.0f}k" for i, n in enumerate(names))

def ab(label, fn, flops):
    for mode in ("0", "2"):
        os.environ["AV2V_GEMM_MC2"] = mode
        os.environ["AV2V_GEMM_DEBUG"] = "0"
        t = timeit(fn, iters=10)
        os.environ["AV2V_GEMM_DEBUG"] = "8"
        fn(); torch.cuda.synchronize()
        print(f"{label} mode={mode}: {t*1e6:.1f} us {flops/t/1e12:.0f} TF | {timers()}", flush=True)

for (M, N, K) in [(12288, 2560, 5120), (12288, 1280, 5120), (49152, 640, 2560), (12288, 3840, 1280), (49152, 5120, 640), (196608, 960, 320)]:
    a = torch.randn(M, K, device=dev).half(); w = (torch.randn(N, K, device=dev) / K ** 0.5).half(); b = torch.randn(N, device=dev).half()
    out = torch.empty(M, N, device=dev, dtype=torch.float16)
    ab(f"linear M={M} N={N} K={K}", lambda: ops.linear(a, w, bias=b, out=out), 2 * M * N * K)
for (NF, H, W, Cin, Cout) in [(48, 64, 64, 320, 320), (48, 32, 32, 640, 640), (48, 16, 16, 1280, 1280), (48, 32, 32, 1280, 640)]:
    x = torch.randn(NF, H, W, Cin, device=dev).half(); w = (torch.randn(Cout, 9 * Cin, device=dev) / (9 * Cin) ** 0.5).half()
    ab(f"conv3x3 NF={NF} {H}x{W} {Cin}->{Cout}", lambda: ops.conv3x3(x, w), 2 * NF * H * W * Cout * 9 * Cin)

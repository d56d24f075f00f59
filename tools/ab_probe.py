"""A/B timing of kernel variants inside ONE process (same box, same clocks)."""
import os, sys, torch
sys.path.insert(0, ".")
from anyv2v_b200 import ops
from tools.gpu_check import timeit
dev = "cuda"
torch.manual_seed(0)
M, N, K = 196608, 2560, 320
a = torch.randn(M, K, device=dev).half(); w = (torch.randn(N, K, device=dev) / K ** 0.5).half(); b = torch.randn(N, device=dev).half()
wp, bp = ops.geglu_pack(w, b)
for rep in range(2):
    t = timeit(lambda: ops.linear(a, wp, bias=bp, geglu=True), iters=10)
    print(f"geglu M={M} N={N} K={K}: {t*1e6:.1f} us {2*M*N*K/t/1e12:.0f} TF", flush=True)
t = timeit(lambda: ops.linear(a, w, bias=b), iters=10)
print(f"plain M={M} N={N} K={K}: {t*1e6:.1f} us {2*M*N*K/t/1e12:.0f} TF", flush=True)
for (n, rows, C) in [(3, 65536, 320), (48, 4096, 320)]:
    x = torch.randn(n, rows, C, device=dev).half(); g = torch.randn(C, device=dev).half(); bb = torch.randn(C, device=dev).half()
    t = timeit(lambda: ops.groupnorm(x, g, bb, 32, 1e-5, True), iters=10)
    print(f"groupnorm n={n} rows={rows} C={C}: {t*1e6:.1f} us -> {4*x.numel()/t/1e9:.0f} GB/s algorithmic", flush=True)

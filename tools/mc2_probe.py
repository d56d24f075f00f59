"""A/B of the cluster-multicast GEMM mode (AV2V_GEMM_MC2=0/1) inside one process: correctness + timing."""
import os, sys, torch
sys.path.insert(0, ".")
from anyv2v_b200 import ops
from tools.gpu_check import timeit
dev = "cuda"
torch.manual_seed(0)

def run(mode):
    os.environ["AV2V_GEMM_MC2"] = mode

for (M, N, K) in [(4096, 1280, 1280), (1000, 320, 320), (196608, 320, 320), (12288, 2560, 5120), (49152, 640, 2560), (384, 640, 640)]:
    a = torch.randn(M, K, device=dev).half(); w = (torch.randn(N, K, device=dev) / K ** 0.5).half(); b = torch.randn(N, device=dev).half()
    r = torch.randn(M, N, device=dev).half()
    outs = {}
    for mode in ("0", "2"):
        run(mode)
        outs[mode] = ops.linear(a, w, bias=b, residual=r)
        torch.cuda.synchronize()
        t = timeit(lambda: ops.linear(a, w, bias=b, residual=r), iters=10)
        print(f"linear+res M={M} N={N} K={K} mc2={mode}: {t*1e6:.1f} us {2*M*N*K/t/1e12:.0f} TF", flush=True)
    print("   equal:", torch.equal(outs["0"], outs["2"]), flush=True)
for (NF, H, W, Cin, Cout) in [(48, 64, 64, 320, 320), (48, 32, 32, 640, 640), (48, 16, 16, 1280, 1280), (3, 8, 8, 128, 160)]:
    x = torch.randn(NF, H, W, Cin, device=dev).half(); w = (torch.randn(Cout, 9 * Cin, device=dev) / (9 * Cin) ** 0.5).half()
    outs = {}
    for mode in ("0", "2"):
        run(mode)
        outs[mode] = ops.conv3x3(x, w)
        torch.cuda.synchronize()
        t = timeit(lambda: ops.conv3x3(x, w), iters=10)
        print(f"conv3x3 NF={NF} {H}x{W} {Cin}->{Cout} mc2={mode}: {t*1e6:.1f} us {2*NF*H*W*Cout*9*Cin/t/1e12:.0f} TF", flush=True)
    print("   equal:", torch.equal(outs["0"], outs["2"]), flush=True)
M, N, K = 196608, 2560, 320
a = torch.randn(M, K, device=dev).half(); w = (torch.randn(N, K, device=dev) / K ** 0.5).half(); b = torch.randn(N, device=dev).half()
wp, bp = ops.geglu_pack(w, b)
for mode in ("0", "2"):
    run(mode)
    t = timeit(lambda: ops.linear(a, wp, bias=bp, geglu=True), iters=10)
    print(f"geglu M={M} N={N} K={K} mc2={mode}: {t*1e6:.1f} us {2*M*N*K/t/1e12:.0f} TF", flush=True)

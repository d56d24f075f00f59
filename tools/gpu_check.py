"""GPU bring-up checks for individual kernels (development aid; the graded parity tests live in tests/).

usage: python tools/gpu_check.py <group>      groups: ddim gn gemm conv tconv attn all
Each group runs in its own process (tools/gpu_check.sh) so a trapping kernel cannot poison the others.
"""
import sys
import time

import torch

sys.path.insert(0, ".")
from anyv2v_b200 import ops  # noqa: E402

dev = "cuda"


def report(name, got, ref, tol=2e-3):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    denom = ref.abs().clamp_min(1e-3)
    rel = (err / denom)
    ok = torch.allclose(got, ref, rtol=1e-3, atol=1e-3 * max(1.0, ref.abs().max().item()) * 0.5)
    print(f"[{name}] max_abs={err.max().item():.3e} mean_abs={err.mean().item():.3e} max_rel={rel.max().item():.3e} "
          f"ref_absmax={ref.abs().max().item():.3f} nan={torch.isnan(got).any().item()} ok={ok}", flush=True)
    return ok


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def g_ddim():
    torch.manual_seed(0)
    n = 4 * 16 * 64 * 64
    x = torch.randn(n, device=dev).half()
    vn = torch.randn(n, device=dev).half()
    ve = torch.randn(n, device=dev).half()
    ca, cb, cc, cd = 0.6, 0.8, 0.7, 0.714
    out = ops.ddim_step(x, vn, ve, 9.0, ca, cb, cc, cd)
    # torch chain with the reference's rounding points
    s = lambda v: torch.tensor(v, dtype=torch.float32)
    v = vn + 9.0 * (ve - vn)
    x0 = s(ca) * x - s(cb) * v
    ep = s(ca) * v + s(cb) * x
    ref = s(cc) * x0 + s(cd) * ep
    print("ddim bit-exact:", torch.equal(out, ref), "max diff", (out.float() - ref.float()).abs().max().item())
    out2 = ops.ddim_step(x, vn, None, 1.0, ca, cb, cc, cd, inverse=True)
    x0 = s(ca) * x - s(cb) * vn
    ep = s(ca) * vn + s(cb) * x
    ref2 = s(cc) * x0 + s(cd) * ep
    print("ddim-inv bit-exact:", torch.equal(out2, ref2))


def g_gn():
    torch.manual_seed(0)
    for (n, rows, C, silu, eps) in [(6, 256, 2560, True, 1e-5), (6, 4096, 320, True, 1e-5), (2, 16 * 1024, 640, False, 1e-6),
                                    (48, 4096, 960, True, 1e-5), (3, 7, 64, True, 1e-5)]:
        x = (torch.randn(n, rows, C, device=dev) * 2 + 0.5).half()
        g = torch.randn(C, device=dev).half()
        b = torch.randn(C, device=dev).half()
        y = ops.groupnorm(x, g, b, 32, eps, silu)
        xr = x.float().permute(0, 2, 1)  # [n, C, rows]
        ref = torch.nn.functional.group_norm(xr, 32, g.float(), b.float(), eps)
        if silu:
            ref = torch.nn.functional.silu(ref.half().float())
        ref = ref.permute(0, 2, 1)
        report(f"gn n={n} rows={rows} C={C} silu={silu}", y, ref)
        if rows >= 4096:
            t = timeit(lambda: ops.groupnorm(x, g, b, 32, eps, silu))
            print(f"   time {t*1e6:.1f} us  -> {4*x.numel()/t/1e9:.0f} GB/s algorithmic", flush=True)


def g_gemm():
    torch.manual_seed(0)
    for (M, N, K, use_bias, use_res) in [(256, 64, 64, False, False), (128, 160, 128, True, False), (1000, 320, 320, True, True),
                                         (4096, 1280, 1280, True, True), (12288, 960, 320, False, False),
                                         (196608, 320, 320, True, True), (49152, 5120, 640, True, False), (12288, 2560, 10240 // 2, True, False)]:
        a = torch.randn(M, K, device=dev).half()
        w = (torch.randn(N, K, device=dev) / K ** 0.5).half()
        bias = torch.randn(N, device=dev).half() if use_bias else None
        res = torch.randn(M, N, device=dev).half() if use_res else None
        out = ops.linear(a, w, bias=bias, residual=res)
        if M * N <= 4096 * 1280 * 4:
            ref = a.float() @ w.float().t()
            if use_bias:
                ref += bias.float()
            if use_res:
                ref += res.float()
            report(f"gemm M={M} N={N} K={K}", out, ref)
        else:
            idx = torch.randint(0, M, (2048,), device=dev)
            ref = a[idx].float() @ w.float().t()
            if use_bias:
                ref += bias.float()
            if use_res:
                ref += res[idx].float()
            report(f"gemm(sampled rows) M={M} N={N} K={K}", out[idx], ref)
        if M >= 4096:
            t = timeit(lambda: ops.linear(a, w, bias=bias, residual=res, out=out))
            print(f"   time {t*1e6:.1f} us -> {2*M*N*K/t/1e12:.1f} TFLOP/s", flush=True)
            tt = timeit(lambda: torch.nn.functional.linear(a, w, bias))
            print(f"   torch(cuBLAS) {tt*1e6:.1f} us -> {2*M*N*K/tt/1e12:.1f} TFLOP/s", flush=True)


def g_conv():
    torch.manual_seed(0)
    for (NF, H, W, Cin, Cout, slots) in [(2, 16, 16, 64, 64, 1), (3, 8, 8, 128, 160, 1), (4, 16, 16, 2560, 1280, 1), (2, 32, 32, 640, 640, 1),
                                         (2, 64, 64, 320, 320, 1), (2, 20, 24, 64, 64, 1), (16, 16, 16, 1280, 1280, 3), (48, 64, 64, 320, 320, 1)]:
        x = torch.randn(NF, H, W, Cin, device=dev).half()
        w = (torch.randn(Cout, Cin, 3, 3, device=dev) / (9 * Cin) ** 0.5).half()
        bias = torch.randn(Cout, device=dev).half()
        temb = torch.randn(NF, Cout, device=dev).half()
        wp = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
        ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).float(), w.float(), bias.float(), padding=1)
        ref = ref + temb.float()[:, :, None, None]
        ref = ref.permute(0, 2, 3, 1)
        if slots == 1:
            out = ops.conv3x3(x, wp, bias=bias, rowbias=temb, rows_per_rowbias=H * W)
            report(f"conv3x3 NF={NF} {H}x{W} {Cin}->{Cout}", out, ref)
        else:
            res = torch.randn(slots, NF, H, W, Cout, device=dev).half()
            out = torch.empty_like(res)
            ops.conv3x3(x, wp, bias=bias, rowbias=temb, rows_per_rowbias=H * W, residual=res, out=out, n_slots=slots,
                        slot_stride=NF * H * W * Cout)
            report(f"conv3x3+inject slots={slots} NF={NF} {H}x{W} {Cin}->{Cout}", out, ref[None] + res.float())
        if NF * H * W >= 4096:
            t = timeit(lambda: ops.conv3x3(x, wp, bias=bias), iters=10)
            fl = 2 * NF * H * W * Cout * 9 * Cin
            print(f"   time {t*1e6:.1f} us -> {fl/t/1e12:.1f} TFLOP/s", flush=True)
            xc = x.permute(0, 3, 1, 2)
            wc = w.contiguous(memory_format=torch.channels_last)
            tt = timeit(lambda: torch.nn.functional.conv2d(xc, wc, bias, padding=1), iters=10)
            print(f"   torch(cuDNN, channels_last) {tt*1e6:.1f} us -> {fl/tt/1e12:.1f} TFLOP/s", flush=True)


def g_tconv():
    torch.manual_seed(0)
    for (B, F, HW, C) in [(1, 4, 64, 64), (2, 8, 256, 320), (3, 16, 1024, 640)]:
        x = torch.randn(B, F * HW, C, device=dev).half()
        w = (torch.randn(C, C, 3, 1, 1, device=dev) / (3 * C) ** 0.5).half()
        bias = torch.randn(C, device=dev).half()
        res = torch.randn(B, F * HW, C, device=dev).half()
        wp = w[:, :, :, 0, 0].permute(0, 2, 1).reshape(C, 3 * C).contiguous()
        out = ops.tconv3(x, wp, F, HW, bias=bias, residual=res)
        x5 = x.view(B, F, HW, 1, C).permute(0, 4, 1, 2, 3).float()  # [B,C,F,HW,1]
        ref = torch.nn.functional.conv3d(x5, w.float(), bias.float(), padding=(1, 0, 0))
        ref = ref.permute(0, 2, 3, 4, 1).reshape(B, F * HW, C) + res.float()
        report(f"tconv3 B={B} F={F} HW={HW} C={C}", out, ref)


GROUPS = {"ddim": g_ddim, "gn": g_gn, "gemm": g_gemm, "conv": g_conv, "tconv": g_tconv}

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    t0 = time.time()
    print("device:", torch.cuda.get_device_name(0), flush=True)
    if which == "attn":
        from tools import gpu_check_attn
        gpu_check_attn.main(sys.argv[2:])
    else:
        for name, fn in GROUPS.items():
            if which in ("all", name):
                print(f"=== {name}", flush=True)
                fn()
                torch.cuda.synchronize()
    print(f"done in {time.time()-t0:.1f}s", flush=True)

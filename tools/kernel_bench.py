"""Kernel-level timings of the hot ops at the shapes of the 16 f x 512^2 step, against their rooflines (development aid; the
numbers quoted in DESIGN.md / profiles/ come from here).  CUDA events, warm, back-to-back launches, inputs > L2 unless noted.

  python tools/kernel_bench.py [gn] [ln] [attn] [tattn] [gemm]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from anyv2v_b200 import ops  # noqa: E402

dev = "cuda"
PK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
HBM, TF_BURST, TF_SUST = PK["hbm_gbs"], PK["bf16_tflops"], PK.get("bf16_tflops_sustained", PK["bf16_tflops"])


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
    return e0.elapsed_time(e1) / iters * 1e3


def gn():
    print(f"--- GroupNorm(+SiLU), us per call; GB/s = 4*n*rows*C / t (algorithmic: one read + one write); HBM peak {HBM:.0f} GB/s")
    for n, rows, C, silu in ((3, 65536, 320, True), (1, 65536, 320, True), (48, 4096, 320, True), (48, 4096, 320, False), (3, 16384, 640, True),
                             (48, 1024, 640, True), (1, 1024, 1280, True), (1, 4096, 1280, True), (16, 4096, 320, True), (48, 4096, 960, True),
                             (16, 4096, 960, True), (3, 4096, 1280, True), (48, 256, 2560, True), (3, 128 * 4096, 320, True)):
        x = torch.randn(n, rows, C, device=dev).half()
        g, b, o = torch.randn(C, device=dev).half(), torch.randn(C, device=dev).half(), torch.empty_like(x)
        t = timeit(lambda: ops.groupnorm(x, g, b, 32, 1e-5, silu, out=o))
        gb = 4.0 * n * rows * C
        print(f"groupnorm n={n:2d} rows={rows:6d} C={C:4d} silu={int(silu)}: {t:8.1f} us  {gb / t / 1e3:6.0f} GB/s  frac {gb / t / 1e3 / HBM:5.2f}  ({gb / 1e6:6.1f} MB)")


def ln():
    print("--- LayerNorm, us per launch, GB/s = 4*rows*C / t")
    for rows, C in ((196608, 320), (65536, 320), (49152, 640), (12288, 1280)):
        x = torch.randn(rows, C, device=dev).half()
        g, b, o = torch.randn(C, device=dev).half(), torch.randn(C, device=dev).half(), torch.empty_like(x)
        t = timeit(lambda: ops.layernorm(x, g, b, 1e-5, out=o))
        print(f"layernorm rows={rows:6d} C={C:4d}: {t:7.1f} us  {4.0 * rows * C / t / 1e3:6.0f} GB/s  frac {4.0 * rows * C / t / 1e3 / HBM:5.2f}")


def attn():
    print(f"--- attention (TF = algorithmic FLOPs / t; burst peak {TF_BURST:.0f})")
    for name, batch, heads, seq, seq_kv, div in (("edit L0 self 48x5x4096", 48, 5, 4096, 0, 0), ("inv L0 self 16x5x4096", 16, 5, 4096, 0, 0),
                                                ("edit L1 self 48x10x1024", 48, 10, 1024, 0, 0), ("edit L2 self 48x20x256", 48, 20, 256, 0, 0),
                                                ("edit L0 cross 48x5x4096 kv145", 48, 5, 4096, 145, 16)):
        C = heads * 64
        q = torch.randn(batch * seq, C, device=dev).half()
        nk = seq_kv or seq
        kvb = batch // div if div else batch
        kv = torch.randn(kvb * nk, 2 * C, device=dev).half()
        out = torch.empty(batch * seq, C, device=dev, dtype=torch.float16)
        t = timeit(lambda: ops.attention(q, kv[:, :C], kv[:, C:], heads, seq, batch, out, seq_kv=seq_kv, kv_batch_div=div))
        fl = 4.0 * batch * heads * seq * nk * 64
        qq = q.view(batch, seq, heads, 64).transpose(1, 2)
        kk = kv[:, :C].reshape(kvb, nk, heads, 64).transpose(1, 2)
        vv = kv[:, C:].reshape(kvb, nk, heads, 64).transpose(1, 2)
        if div:
            kk, vv = kk.repeat_interleave(div, 0), vv.repeat_interleave(div, 0)
        ts = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(qq, kk, vv))
        print(f"nv=1 {name:32s}: {t:8.1f} us {fl / t / 1e6:7.1f} TF frac {fl / t / 1e6 / TF_BURST:4.2f} | torch SDPA {ts:8.1f} us")
    for name, batch, heads, seq in (("L0 16x5x4096", 16, 5, 4096), ("L1 16x10x1024", 16, 10, 1024), ("L2 16x20x256", 16, 20, 256)):
        C = heads * 64
        rows = batch * seq
        qk = torch.randn(rows, 2 * C, device=dev).half()
        v = torch.randn(3 * rows, C, device=dev).half()
        out = torch.empty(3 * rows, C, device=dev, dtype=torch.float16)
        t = timeit(lambda: ops.attention(qk[:, :C], qk[:, C:], v, heads, seq, batch, out, n_v=3, v_branch_stride=rows * C, o_branch_stride=rows * C))
        fl = 2.0 * batch * heads * seq * seq * 64 * 4
        print(f"nv=3 {name:32s}: {t:8.1f} us {fl / t / 1e6:7.1f} TF frac {fl / t / 1e6 / TF_BURST:4.2f}")


def tattn():
    print("--- temporal self-attention, Q/K/V projection fused (one kernel): us; GB/s = (x read + o written) / t; vs QKV GEMM + frames-mode attention")
    for clips, heads, F, HW, Cx in ((3, 5, 16, 4096, 320), (1, 5, 16, 4096, 320), (3, 8, 16, 4096, 320), (3, 10, 16, 1024, 640), (3, 20, 16, 256, 1280),
                                    (1, 5, 128, 4096, 320)):
        C = heads * 64
        rows = clips * F * HW
        x = torch.randn(rows, Cx, device=dev).half()
        w = (torch.randn(3 * C, Cx, device=dev) / Cx ** 0.5).half()
        o = torch.empty(rows, C, device=dev, dtype=torch.float16)
        qkv = torch.empty(rows, 3 * C, device=dev, dtype=torch.float16)

        def two():
            ops.linear(x, w, out=qkv)
            ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], heads, F, clips * HW, o, frames_mode=True, HW=HW)

        t2 = timeit(two)
        t1 = timeit(lambda: ops.temporal_attention_fused(x, w, heads, F, HW, clips, o))
        t3 = timeit(lambda: ops.temporal_attention_fused(x, w, heads, F, HW, clips, o, n_v=3)) if clips % 3 == 0 else float("nan")
        nb = rows * (Cx + C) * 2.0
        print(f"clips={clips} heads={heads:2d} F={F:3d} HW={HW:4d} Cx={Cx:4d}: two kernels {t2:7.1f} us | fused nv=1 {t1:7.1f} us ({nb / t1 / 1e3:5.0f} GB/s, frac {nb / t1 / 1e3 / HBM:4.2f})"
              f" | fused nv=3 (injected) {t3:7.1f} us")


def gemm():
    print(f"--- GEMM shapes of the step (TF vs sustained peak {TF_SUST:.0f})")
    for M, N, K, kind in ((196608, 960, 320, ""), (196608, 320, 320, "res"), (196608, 2560, 320, "geglu"), (196608, 320, 1280, "res"), (65536, 2560, 320, "geglu"),
                          (49152, 640, 640, "res"), (49152, 5120, 640, "geglu"), (49152, 640, 2560, "res"), (12288, 1280, 1280, "res"), (12288, 10240, 1280, "geglu"),
                          (12288, 1280, 5120, "res")):
        a = torch.randn(M, K, device=dev).half()
        w = (torch.randn(N, K, device=dev) / K ** 0.5).half()
        b = torch.randn(N, device=dev).half()
        if kind == "geglu":
            wp, bp = ops.geglu_pack(w, b)
            o = torch.empty(M, N // 2, device=dev, dtype=torch.float16)
            fn = lambda: ops.linear(a, wp, bias=bp, geglu=True, out=o)
        else:
            r = torch.randn(M, N, device=dev).half() if kind == "res" else None
            o = torch.empty(M, N, device=dev, dtype=torch.float16)
            fn = lambda: ops.linear(a, w, bias=b, residual=r, out=o)
        t = timeit(fn)
        print(f"linear{'+' + kind if kind else '':7s} M={M:6d} N={N:5d} K={K:4d}: {t:7.1f} us {2.0 * M * N * K / t / 1e6:7.1f} TF frac {2.0 * M * N * K / t / 1e6 / TF_SUST:4.2f}")
    for NF, HWs, Cin, Cout in ((48, 64, 320, 320), (48, 32, 640, 640), (48, 16, 1280, 1280), (16, 16, 2560, 1280), (48, 64, 960, 320)):
        x = torch.randn(NF, HWs, HWs, Cin, device=dev).half()
        w = (torch.randn(Cout, 9 * Cin, device=dev) / (9 * Cin) ** 0.5).half()
        t = timeit(lambda: ops.conv3x3(x, w))
        fl = 2.0 * NF * HWs * HWs * Cout * 9 * Cin
        print(f"conv3x3 NF={NF} {HWs}x{HWs} {Cin}->{Cout}: {t:7.1f} us {fl / t / 1e6:7.1f} TF frac {fl / t / 1e6 / TF_SUST:4.2f}")


if __name__ == "__main__":
    which = sys.argv[1:] or ["gn", "ln", "attn", "tattn", "gemm"]
    for name in which:
        globals()[name]()

"""Round-2 opening move: ONE gpurun call that tells which of the default-off candidates to switch on.

  /usr/local/graft/bin/gpurun --timeout 2700 -- 'python tools/r2_probe.py > gpurun_out/r2_probe.txt 2>&1'   (~30 GPU-minutes)
  (or in two calls: `python tools/r2_probe.py parity kernels`, then `python tools/r2_probe.py bench`)

1. parity: tests/test_gpu_experimental.py with AV2V_EXPERIMENTAL=1 (each candidate against the fp32 restatement and,
   where the arithmetic is unchanged, bit-for-bit against the shipped kernels);
2. kernel A/B (CUDA events, warm, back-to-back launches, inputs > L2): attention shapes of the step under
   AV2V_ATTN_2Q = 0/1/2/3 (+ torch SDPA for scale), "+ residual" GEMM shapes under AV2V_GEMM_RESBUFS = 2/4;
3. whole job: bench.py in sub-processes under each switch combination (the switches are read at CUDA-graph capture).

Every mbarrier wait in the kernels is bounded (ptx.cuh: trap after 4e9 cycles), so a protocol bug shows up as a CUDA
error in that sub-test, not as a hung box; each stage also runs under its own `timeout`.
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def sh(cmd, env=None, timeout=900):
    e = dict(os.environ)
    e.update(env or {})
    t0 = time.time()
    try:
        r = subprocess.run(cmd, shell=True, cwd=ROOT, env=e, capture_output=True, text=True, timeout=timeout)
        return r.returncode, r.stdout + r.stderr, time.time() - t0
    except subprocess.TimeoutExpired as ex:
        return 124, (ex.stdout or "") + (ex.stderr or "") if isinstance(ex.stdout, str) else "timeout", time.time() - t0


def stage_parity():
    print("=" * 100 + "\n[1] parity of the candidates (tests/test_gpu_experimental.py)", flush=True)
    rc, out, dt = sh("python -m pytest tests/test_gpu_experimental.py -q -m gpu -x --timeout 600 2>&1 | tail -25",
                     {"AV2V_EXPERIMENTAL": "1"}, timeout=1200)
    print(out.strip(), f"\n[parity rc={rc} {dt:.0f}s]", flush=True)
    # per candidate, so that one broken candidate does not hide the others
    for name, k in (("attention 2q", "attention_2q"), ("fused temporal attention", "temporal_attention_fused"), ("attention v10", "attention_v10"), ("pdl", "pdl"), ("ping-pong traversal", "pingpong"), ("deep residual prefetch", "deep_residual"), ("packed geglu", "geglu_packed"), ("W-stationary gemm", "w_stationary"), ("layernorm v2", "layernorm_v2"), ("groupnorm v2", "groupnorm_v2"), ("groupnorm cluster", "groupnorm_cluster"), ("shared prefix", "shared_uncond"), ("all together", "all_candidates")):
        rc, out, dt = sh(f"python -m pytest tests/test_gpu_experimental.py -q -m gpu -k {k} --timeout 600 2>&1 | tail -4",
                         {"AV2V_EXPERIMENTAL": "1"}, timeout=1200)
        print(f"  {name:28s} rc={rc} {dt:5.0f}s  {out.strip().splitlines()[-1] if out.strip() else ''}", flush=True)


KERNEL_AB = r'''
import os, sys, torch
sys.path.insert(0, %r)
from anyv2v_b200 import ops
dev = "cuda"

def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

def setenv(**kv):
    for k, v in kv.items():
        if v is None: os.environ.pop(k, None)
        else: os.environ[k] = str(v)

import ctypes
from anyv2v_b200 import _lib
_lib.lib().av2v_gemm_debug_timers.argtypes = [ctypes.c_void_p]
def role_timers():
    buf = (ctypes.c_ulonglong * 16)(); _lib.lib().av2v_gemm_debug_timers(buf)
    return " ".join(f"{n}={buf[i] / 1e3:.0f}k" for i, n in enumerate(("prod_wait_empty", "prod_total", "mma_wait_tempty", "mma_wait_full", "mma_total")))

print("--- attention, n_v = 1 (us per launch; TF = 4*B*H*N*L*64 / t)")
for name, batch, heads, seq, seq_kv, div in (("edit  L0 self  48x5x4096", 48, 5, 4096, 0, 0), ("inv   L0 self  16x5x4096", 16, 5, 4096, 0, 0),
                                            ("edit  L1 self  48x10x1024", 48, 10, 1024, 0, 0), ("edit  L2 self  48x20x256", 48, 20, 256, 0, 0),
                                            ("edit  L0 cross 48x5x4096 kv145", 48, 5, 4096, 145, 16), ("edit  L1 cross 48x10x1024 kv145", 48, 10, 1024, 145, 16)):
    C = heads * 64
    q = torch.randn(batch * seq, C, device=dev).half()
    nk = seq_kv or seq
    kvb = batch // div if div else batch
    kv = torch.randn(kvb * nk, 2 * C, device=dev).half()
    out = torch.empty(batch * seq, C, device=dev, dtype=torch.float16)
    fn = lambda: ops.attention(q, kv[:, :C], kv[:, C:], heads, seq, batch, out, seq_kv=seq_kv, kv_batch_div=div)
    flops = 4.0 * batch * heads * seq * nk * 64
    row = []
    ref = None
    for mode in (0, 1, 2, 3, 4):
        setenv(AV2V_ATTN_2Q=mode if mode else None)
        try:
            t = timeit(fn)
            o = out.float().clone()
            if ref is None: ref = o
            row.append(f"2q={mode}: {t:8.1f} us {flops / t / 1e6:7.1f} TF maxdiff {float((o - ref).abs().max()):.1e}")
        except Exception as ex:
            row.append(f"2q={mode}: FAILED {str(ex)[:80]}")
    setenv(AV2V_ATTN_2Q=None)
    qq = q.view(batch, seq, heads, 64).transpose(1, 2)
    kk = kv[:, :C].reshape(kvb, nk, heads, 64).transpose(1, 2)
    vv = kv[:, C:].reshape(kvb, nk, heads, 64).transpose(1, 2)
    if div: kk, vv = kk.repeat_interleave(div, 0), vv.repeat_interleave(div, 0)
    t = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(qq, kk, vv))
    print(f"{name:34s} " + " | ".join(row) + f" | torch SDPA {t:8.1f} us")

print("--- injected attention, n_v = 3 (the roofline kernel): v9 vs v10 (AV2V_ATTN_V10), TF = 2*B*H*N*N*64*(1+3) / t")
for name, batch, heads, seq in (("L0 16x5x4096", 16, 5, 4096), ("L1 16x10x1024", 16, 10, 1024), ("L2 16x20x256", 16, 20, 256)):
    C = heads * 64; rows = batch * seq
    qk = torch.randn(rows, 2 * C, device=dev).half(); v = torch.randn(3 * rows, C, device=dev).half()
    out = torch.empty(3 * rows, C, device=dev, dtype=torch.float16)
    fn = lambda: ops.attention(qk[:, :C], qk[:, C:], v, heads, seq, batch, out, n_v=3, v_branch_stride=rows * C, o_branch_stride=rows * C)
    flops = 2.0 * batch * heads * seq * seq * 64 * 4
    setenv(AV2V_ATTN_V10=None); t9 = timeit(fn); o9 = out.float().clone()
    try:
        setenv(AV2V_ATTN_V10=1); t10 = timeit(fn); d = float((out.float() - o9).abs().max())
        setenv(AV2V_ATTN_V10=2); t10p = timeit(fn); dp = float((out.float() - o9).abs().max())
        setenv(AV2V_ATTN_V10=3); t10q = timeit(fn); dq = float((out.float() - o9).abs().max())
        print(f"nv=3 {name:14s}: v9 {t9:8.1f} us {flops / t9 / 1e6:7.1f} TF -> v10 {t10:8.1f} us {flops / t10 / 1e6:7.1f} TF maxdiff {d:.1e}"
              f" -> v10+poly {t10p:8.1f} us {flops / t10p / 1e6:7.1f} TF maxdiff {dp:.1e} -> v10+packed {t10q:8.1f} us {flops / t10q / 1e6:7.1f} TF maxdiff {dq:.1e}")
    except Exception as ex:
        print(f"nv=3 {name:14s}: v9 {t9:8.1f} us; v10 FAILED {str(ex)[:100]}")
    setenv(AV2V_ATTN_V10=None)
for name, clips, heads, F, HW in (("temporal L0 3x5 F16 HW4096", 3, 5, 16, 4096), ("temporal L0 1x5 F128 HW4096", 1, 5, 128, 4096)):
    C = heads * 64; rows = clips * F * HW
    x = torch.randn(rows, 3 * C, device=dev).half(); out = torch.empty(rows, C, device=dev, dtype=torch.float16)
    fn = lambda: ops.attention(x[:, :C], x[:, C:2 * C], x[:, 2 * C:], heads, F, clips * HW, out, frames_mode=True, HW=HW)
    setenv(AV2V_ATTN_V10=None); t9 = timeit(fn); o9 = out.float().clone()
    try:
        setenv(AV2V_ATTN_V10=1); t10 = timeit(fn); d = float((out.float() - o9).abs().max())
        print(f"{name:30s}: v9 {t9:8.1f} us -> v10 {t10:8.1f} us  maxdiff {d:.1e}")
    except Exception as ex:
        print(f"{name:30s}: v9 {t9:8.1f} us; v10 FAILED {str(ex)[:100]}")
    setenv(AV2V_ATTN_V10=None)

print("--- temporal self-attention: QKV GEMM + attention (two kernels) vs the fused kernel (AV2V_TATTN_FUSED), us")
for clips, heads, F, HW, Cx in ((3, 5, 16, 4096, 320), (1, 5, 16, 4096, 320), (3, 8, 16, 4096, 512), (3, 10, 16, 1024, 640), (3, 20, 16, 256, 1280)):
    C = heads * 64; rows = clips * F * HW
    x = torch.randn(rows, Cx, device=dev).half(); w = (torch.randn(3 * C, Cx, device=dev) / Cx ** 0.5).half()
    o1 = torch.empty(rows, C, device=dev, dtype=torch.float16); o2 = torch.empty_like(o1); qkv = torch.empty(rows, 3 * C, device=dev, dtype=torch.float16)
    def two():
        ops.linear(x, w, out=qkv)
        ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], heads, F, clips * HW, o1, frames_mode=True, HW=HW)
    t2 = timeit(two)
    try:
        t1 = timeit(lambda: ops.temporal_attention_fused(x, w, heads, F, HW, clips, o2))
        print(f"temporal attn clips={clips} heads={heads:2d} F={F} HW={HW:4d} Cx={Cx:4d}: two kernels {t2:7.1f} us -> fused {t1:7.1f} us ({t2 / t1:4.2f}x) "
              f"maxdiff {float((o1.float() - o2.float()).abs().max()):.1e} bit-identical={torch.equal(o1, o2)}")
    except Exception as ex:
        print(f"temporal attn clips={clips} heads={heads} HW={HW}: two kernels {t2:7.1f} us; fused FAILED {str(ex)[:100]}")

print("--- GEMM + residual, AV2V_GEMM_RESBUFS 2 (shipped) vs 4 (us per launch)")
for M, N, K in ((196608, 320, 320), (65536, 320, 320), (49152, 640, 640), (16384, 640, 640), (12288, 1280, 1280), (196608, 320, 1280), (49152, 640, 2560)):
    a = torch.randn(M, K, device=dev).half(); w = (torch.randn(N, K, device=dev) / K ** 0.5).half()
    b = torch.randn(N, device=dev).half(); r = torch.randn(M, N, device=dev).half(); o = torch.empty(M, N, device=dev, dtype=torch.float16)
    fn = lambda: ops.linear(a, w, bias=b, residual=r, out=o)
    setenv(AV2V_GEMM_RESBUFS=None); t2 = timeit(fn); o2 = o.clone()
    setenv(AV2V_GEMM_RESBUFS=4); t4 = timeit(fn); same = torch.equal(o, o2)
    setenv(AV2V_GEMM_RESBUFS=None)
    print(f"linear+res M={M:6d} N={N:4d} K={K:4d}: {t2:7.1f} -> {t4:7.1f} us ({t2 / t4:4.2f}x) bit-identical={same}")
print("--- short-K GEMMs of the 64x64 level: shipped schedule vs W-stationary (AV2V_GEMM_WRES), us per launch + role timers of CTA 0")
for M, N, K, kind in ((196608, 960, 320, ""), (196608, 320, 320, ""), (196608, 320, 320, "res"), (196608, 2560, 320, "geglu"), (65536, 960, 320, ""), (65536, 2560, 320, "geglu")):
    a = torch.randn(M, K, device=dev).half(); w = (torch.randn(N, K, device=dev) / K ** 0.5).half(); b = torch.randn(N, device=dev).half()
    if kind == "geglu":
        wp, bp = ops.geglu_pack(w, b); o = torch.empty(M, N // 2, device=dev, dtype=torch.float16)
        fn = lambda: ops.linear(a, wp, bias=bp, geglu=True, out=o)
    else:
        r = torch.randn(M, N, device=dev).half() if kind == "res" else None; o = torch.empty(M, N, device=dev, dtype=torch.float16)
        fn = lambda: ops.linear(a, w, bias=b, residual=r, out=o)
    line = []
    ref = None
    for wres in (None, 1):
        setenv(AV2V_GEMM_WRES=wres, AV2V_GEMM_DEBUG=None)
        try:
            t = timeit(fn); cur = o.clone()
            setenv(AV2V_GEMM_DEBUG=8); fn(); torch.cuda.synchronize(); tm = role_timers(); setenv(AV2V_GEMM_DEBUG=None)
            if ref is None: ref = cur
            line.append(f"wres={wres or 0}: {t:7.1f} us {2.0 * M * N * K / t / 1e6:7.1f} TF same={torch.equal(cur, ref)} [{tm}]")
        except Exception as ex:
            line.append(f"wres={wres}: FAILED {str(ex)[:80]}")
    setenv(AV2V_GEMM_WRES=None)
    print(f"linear{'+' + kind if kind else '':7s} M={M:6d} N={N:5d} K={K:4d}: " + " | ".join(line))

print("--- fused GEGLU GEMM, scalar vs packed fp32x2 epilogue (AV2V_GEGLU_PACKED), us per launch")
for M, N, K in ((196608, 2560, 320), (65536, 2560, 320), (49152, 5120, 640), (12288, 10240, 1280)):
    a = torch.randn(M, K, device=dev).half(); w = (torch.randn(N, K, device=dev) / K ** 0.5).half(); b = torch.randn(N, device=dev).half()
    wp, bp = ops.geglu_pack(w, b); o = torch.empty(M, N // 2, device=dev, dtype=torch.float16)
    fn = lambda: ops.linear(a, wp, bias=bp, geglu=True, out=o)
    setenv(AV2V_GEGLU_PACKED=None); t1 = timeit(fn); o1 = o.clone()
    setenv(AV2V_GEGLU_PACKED=1); t2 = timeit(fn); same = torch.equal(o, o1)
    setenv(AV2V_GEGLU_PACKED=None)
    print(f"geglu M={M:6d} N={N:5d} K={K:4d}: {t1:7.1f} us {2.0 * M * N * K / t1 / 1e6:7.1f} TF -> {t2:7.1f} us {2.0 * M * N * K / t2 / 1e6:7.1f} TF bit-identical={same}")

print("--- what the residual costs at long K (step profile: +res GEMMs run at 600-800 TF where the plain ones reach 1.1-1.4 PF); role timers of CTA 0")
for M, N, K in ((12288, 1280, 5120), (49152, 640, 2560), (196608, 320, 1280), (12288, 1280, 1280)):
    a = torch.randn(M, K, device=dev).half(); w = (torch.randn(N, K, device=dev) / K ** 0.5).half()
    b = torch.randn(N, device=dev).half(); r = torch.randn(M, N, device=dev).half(); o = torch.empty(M, N, device=dev, dtype=torch.float16)
    for res, rb in ((None, None), (r, None), (r, 4)):
        setenv(AV2V_GEMM_RESBUFS=rb, AV2V_GEMM_DEBUG=None)
        fn = lambda: ops.linear(a, w, bias=b, residual=res, out=o)
        t = timeit(fn)
        setenv(AV2V_GEMM_DEBUG=8); fn(); torch.cuda.synchronize(); tm = role_timers(); setenv(AV2V_GEMM_DEBUG=None)
        print(f"linear M={M:6d} N={N:4d} K={K:4d} res={'yes' if res is not None else 'no ':3s} resbufs={rb or 2}: {t:7.1f} us {2.0 * M * N * K / t / 1e6:7.1f} TF | {tm}")
setenv(AV2V_GEMM_RESBUFS=None)

for NF, HW, C in ((48, 64, 320), (48, 32, 640)):
    x = torch.randn(NF, HW, HW, C, device=dev).half(); w = (torch.randn(C, 9 * C, device=dev) / (9 * C) ** 0.5).half()
    r = torch.randn(NF, HW, HW, C, device=dev).half()
    fn = lambda: ops.conv3x3(x, w, residual=r)
    setenv(AV2V_GEMM_RESBUFS=None); t2 = timeit(fn); o2 = fn()
    setenv(AV2V_GEMM_RESBUFS=4); t4 = timeit(fn); same = torch.equal(fn(), o2)
    setenv(AV2V_GEMM_RESBUFS=None)
    print(f"conv3x3+res NF={NF} {HW}x{HW} C={C}: {t2:7.1f} -> {t4:7.1f} us ({t2 / t4:4.2f}x) bit-identical={same}")

print("--- GroupNorm(+SiLU) v1 vs v2 (AV2V_GN_V2), us per call (2 kernels), GB/s = 6*n*rows*C / t (two reads + one write)")
for n, rows, C, silu in ((3, 65536, 320, True), (1, 65536, 320, True), (48, 4096, 320, True), (48, 4096, 320, False), (3, 16384, 640, True), (48, 1024, 640, True), (1, 1024, 1280, True), (1, 4096, 1280, True), (16, 4096, 320, True), (48, 4096, 960, True)):
    x = torch.randn(n, rows, C, device=dev).half(); g = torch.randn(C, device=dev).half(); b = torch.randn(C, device=dev).half(); o = torch.empty_like(x)
    fn = lambda: ops.groupnorm(x, g, b, 32, 1e-5, silu, out=o)
    setenv(AV2V_GN_V2=None); t1 = timeit(fn); o1 = o.clone()
    setenv(AV2V_GN_V2=1); t2 = timeit(fn); d = float((o.float() - o1.float()).abs().max())
    setenv(AV2V_GN_V2=None, AV2V_GN_CLUSTER=1)
    try:
        t3 = timeit(fn); d3 = float((o.float() - o1.float()).abs().max())
    except Exception as ex:
        t3, d3 = float("nan"), str(ex)[:60]
    setenv(AV2V_GN_CLUSTER=None)
    gb = 6.0 * n * rows * C
    print(f"groupnorm n={n:2d} rows={rows:6d} C={C:4d} silu={int(silu)}: v1 {t1:7.1f} us ({gb / t1 / 1e3:6.0f} GB/s) -> v2 {t2:7.1f} us ({gb / t2 / 1e3:6.0f} GB/s) maxdiff {d:.1e}"
          f" -> cluster (per-frame shapes only, else fallback) {t3:7.1f} us maxdiff {d3}")

print("--- LayerNorm v1 vs v2 (AV2V_LN_V2), us per launch, GB/s = 4*rows*C / t")
for rows, C in ((196608, 320), (65536, 320), (49152, 640), (12288, 1280)):
    x = torch.randn(rows, C, device=dev).half(); g = torch.randn(C, device=dev).half(); b = torch.randn(C, device=dev).half()
    o = torch.empty_like(x)
    fn = lambda: ops.layernorm(x, g, b, 1e-5, out=o)
    setenv(AV2V_LN_V2=None); t1 = timeit(fn); o1 = o.clone()
    setenv(AV2V_LN_V2=1); t2 = timeit(fn); d = float((o.float() - o1.float()).abs().max())
    setenv(AV2V_LN_V2=None)
    print(f"layernorm rows={rows:6d} C={C:4d}: {t1:7.1f} us ({4.0 * rows * C / t1 / 1e3:6.0f} GB/s) -> {t2:7.1f} us ({4.0 * rows * C / t2 / 1e3:6.0f} GB/s) maxdiff {d:.1e}")

print("--- PDL on a chain of short kernels (GroupNorm -> conv -> linear+res -> LayerNorm -> qkv -> attention), CUDA graph replay, us per chain")
from tests.test_gpu_experimental import _chain, _chain_inputs
args = _chain_inputs(C=320, NF=16, H=32, W=32)
for pdl in (None, 1):
    setenv(AV2V_PDL=pdl)
    _chain(ops, *args); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s): _chain(ops, *args)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g): _chain(ops, *args)
    print(f"AV2V_PDL={pdl}: {timeit(g.replay, iters=50):8.1f} us")
setenv(AV2V_PDL=None)
''' % ROOT


def stage_kernels():
    """every `print("--- ...")` section of KERNEL_AB runs in its own process (shared preamble): a kernel that traps poisons
    only its own CUDA context"""
    print("=" * 100 + "\n[2] kernel A/B", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    parts = KERNEL_AB.split('\nprint("--- ')
    preamble, sections = parts[0], ['print("--- ' + sec for sec in parts[1:]]
    for i, sec in enumerate(sections):
        path = os.path.join(ROOT, "gpurun_out", f"_r2_kernel_ab_{i}.py")
        with open(path, "w") as fh:
            fh.write(preamble + "\n" + sec)
        rc, out, dt = sh(f"python {path}", {"AV2V_EXPERIMENTAL": "1", "PYTHONPATH": ROOT + os.pathsep + os.path.join(ROOT, "tests")}, timeout=600)
        print(out.strip()[-6000:], f"\n[section {i} rc={rc} {dt:.0f}s]", flush=True)


def stage_bench(steps=10):
    print("=" * 100 + "\n[3] whole job (bench.py --no-cpu-baseline) per switch combination", flush=True)
    ALL = {"AV2V_PDL": "1", "AV2V_PINGPONG": "1", "AV2V_GEMM_RESBUFS": "4", "AV2V_GEMM_WRES": "1", "AV2V_GEGLU_PACKED": "1", "AV2V_LN_V2": "1", "AV2V_GN_V2": "1", "AV2V_GN_CLUSTER": "1", "AV2V_TATTN_FUSED": "1", "AV2V_SHARED_PREFIX": "1", "AV2V_PRUNE_SOURCE": "1"}
    combos = [("shipped", {}),
              ("PDL", {"AV2V_PDL": "1"}),
              ("PINGPONG", {"AV2V_PINGPONG": "1"}),
              ("RESBUFS=4", {"AV2V_GEMM_RESBUFS": "4"}),
              ("GEMM_WRES", {"AV2V_GEMM_WRES": "1"}),
              ("GN_V2+CLUSTER", {"AV2V_GN_V2": "1", "AV2V_GN_CLUSTER": "1"}),
              ("TATTN_FUSED", {"AV2V_TATTN_FUSED": "1"}),
              ("SHARED_PREFIX", {"AV2V_SHARED_PREFIX": "1"}),
              ("PRUNE_SOURCE", {"AV2V_PRUNE_SOURCE": "1"}),
              ("all w/o attention", ALL),
              ("all + 2Q=1 + V10=1", dict(ALL, AV2V_ATTN_2Q="1", AV2V_ATTN_V10="1")),
              ("all + 2Q=4 + V10=3", dict(ALL, AV2V_ATTN_2Q="4", AV2V_ATTN_V10="3"))]
    for name, env in combos:
        rc, out, dt = sh(f"python bench.py --steps {steps} --warmup 4 --no-cpu-baseline", env, timeout=600)
        line = next((l for l in out.splitlines()[::-1] if l.startswith("{")), None)
        if rc == 0 and line:
            d = json.loads(line)
            c = d["config"]
            print(f"  {name:20s} {d['value']:7.3f} steps/s  inv {c['ms_per_inversion_step']:6.2f} ms  edit {c['ms_per_edit_step']:6.2f} ms  "
                  f"e2e {d['e2e']['value']:7.3f}  finite={c['outputs_finite']}  clocks={d['clocks']}", flush=True)
        else:
            print(f"  {name:20s} FAILED rc={rc} {dt:.0f}s: {out.strip()[-400:]}", flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["parity", "kernels", "bench"]
    if "parity" in which:
        stage_parity()
    if "kernels" in which:
        stage_kernels()
    if "bench" in which:
        stage_bench()

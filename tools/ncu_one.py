"""One launch-set of a single op for an `ncu --set full` capture (B200_PROFILING.md recipe):
  ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 1 -o gpurun_out/<name> python tools/ncu_one.py <op>
ops: gn (GroupNorm+SiLU [3,65536,320]) | gn1 ([1,65536,320]) | tattn (fused temporal attention nv=1, 3 clips) | tattn3 (injected) |
     attn3 (spatial PnP attention nv=3, 16x5x4096) | attn1 (plain attention 48x5x4096) | geglu (M=196608 N=2560 K=320) | lin960 / linres (K = 320 GEMMs) | ln"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anyv2v_b200 import ops  # noqa: E402

dev = "cuda"
op = sys.argv[1]
torch.manual_seed(0)
if op in ("gn", "gn1"):
    n = 3 if op == "gn" else 1
    x = torch.randn(n, 65536, 320, device=dev).half()
    g, b, o = torch.randn(320, device=dev).half(), torch.randn(320, device=dev).half(), torch.empty_like(x)
    fn = lambda: ops.groupnorm(x, g, b, 32, 1e-5, True, out=o)
elif op in ("tattn", "tattn3"):
    heads, F, HW, clips = 5, 16, 4096, 3
    rows = clips * F * HW
    x = torch.randn(rows, 320, device=dev).half()
    w = (torch.randn(960, 320, device=dev) / 18).half()
    o = torch.empty(rows, 320, device=dev, dtype=torch.float16)
    fn = lambda: ops.temporal_attention_fused(x, w, heads, F, HW, clips, o, n_v=3 if op == "tattn3" else 1)
elif op == "attn3":
    heads, seq, batch = 5, 4096, 16
    rows = batch * seq
    qk = torch.randn(rows, 640, device=dev).half()
    v = torch.randn(3 * rows, 320, device=dev).half()
    o = torch.empty(3 * rows, 320, device=dev, dtype=torch.float16)
    fn = lambda: ops.attention(qk[:, :320], qk[:, 320:], v, heads, seq, batch, o, n_v=3, v_branch_stride=rows * 320, o_branch_stride=rows * 320)
elif op == "attn1":
    heads, seq, batch = 5, 4096, 48
    qkv = torch.randn(batch * seq, 960, device=dev).half()
    o = torch.empty(batch * seq, 320, device=dev, dtype=torch.float16)
    fn = lambda: ops.attention(qkv[:, :320], qkv[:, 320:640], qkv[:, 640:], heads, seq, batch, o)
elif op == "geglu":
    a = torch.randn(196608, 320, device=dev).half()
    w = (torch.randn(2560, 320, device=dev) / 18).half()
    bb = torch.randn(2560, device=dev).half()
    wp, bp = ops.geglu_pack(w, bb)
    o = torch.empty(196608, 1280, device=dev, dtype=torch.float16)
    fn = lambda: ops.linear(a, wp, bias=bp, geglu=True, out=o)
elif op in ("lin960", "linres"):
    M, K = 196608, 320
    N = 960 if op == "lin960" else 320
    a = torch.randn(M, K, device=dev).half()
    w = (torch.randn(N, K, device=dev) / 18).half()
    bb = torch.randn(N, device=dev).half()
    r = torch.randn(M, N, device=dev).half() if op == "linres" else None
    o = torch.empty(M, N, device=dev, dtype=torch.float16)
    fn = lambda: ops.linear(a, w, bias=bb, residual=r, out=o)
elif op == "ln":
    x = torch.randn(196608, 320, device=dev).half()
    g, b, o = torch.randn(320, device=dev).half(), torch.randn(320, device=dev).half(), torch.empty_like(x)
    fn = lambda: ops.layernorm(x, g, b, 1e-5, out=o)
else:
    raise SystemExit(f"unknown op {op}")
for _ in range(3):
    fn()
torch.cuda.synchronize()

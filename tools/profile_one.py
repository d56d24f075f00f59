"""Run one kernel shape a few times (for ncu captures). usage: python tools/profile_one.py <what>"""
import sys

import torch

sys.path.insert(0, ".")
from anyv2v_b200 import ops  # noqa: E402

what = sys.argv[1]
dev = "cuda"
torch.manual_seed(0)
if what in ("attn1", "attn3"):
    heads, seq = 5, 4096
    batch = 48 if what == "attn1" else 16
    C = heads * 64
    nb = 1 if what == "attn1" else 3
    qkv = torch.randn(nb * batch * seq, 3 * C, device=dev).half()
    q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    out = torch.empty(nb * batch * seq, C, device=dev, dtype=torch.float16)
    rows = batch * seq
    for _ in range(4):
        if what == "attn1":
            ops.attention(q, k, v, heads, seq, batch, out)
        else:
            ops.attention(q[:rows], k[:rows], v, heads, seq, batch, out, n_v=3, v_branch_stride=rows * 3 * C,
                          o_branch_stride=rows * C)
elif what == "gemm2":
    M, N, K = 196608, 2560, 320
    a = torch.randn(M, K, device=dev).half()
    w = torch.randn(N, K, device=dev).half()
    b = torch.randn(N, device=dev).half()
    for _ in range(4):
        ops.linear(a, w, bias=b)
elif what == "gemm":
    M, N, K = 196608, 320, 320
    a = torch.randn(M, K, device=dev).half()
    w = torch.randn(N, K, device=dev).half()
    b = torch.randn(N, device=dev).half()
    for _ in range(4):
        ops.linear(a, w, bias=b)
elif what == "conv":
    x = torch.randn(48, 64, 64, 320, device=dev).half()
    w = torch.randn(320, 9 * 320, device=dev).half()
    for _ in range(4):
        ops.conv3x3(x, w)
torch.cuda.synchronize()
print("ok")

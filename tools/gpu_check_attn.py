"""Bring-up checks for the tcgen05 attention kernel (development aid)."""
import sys

import torch

sys.path.insert(0, ".")
from anyv2v_b200 import ops  # noqa: E402
from tools.gpu_check import report, timeit  # noqa: E402

dev = "cuda"


def ref_attn(q, k, v, heads):
    # q,k,v: [batch, seq, heads*64] fp16 -> fp32 reference
    B, N, C = q.shape
    qh = q.float().view(B, N, heads, 64).transpose(1, 2)
    kh = k.float().view(B, N, heads, 64).transpose(1, 2)
    vh = v.float().view(B, N, heads, 64).transpose(1, 2)
    s = (qh @ kh.transpose(-1, -2)) * 0.125
    p = torch.softmax(s, dim=-1)
    return (p @ vh).transpose(1, 2).reshape(B, N, C)


def rows_case(batch, heads, seq, nv, scale_in=1.0, timing=False):
    torch.manual_seed(1)
    C = heads * 64
    nb = 3 if nv == 3 else 1
    # fused qkv buffer [nb*batch*seq, 3C]
    qkv = (torch.randn(nb * batch * seq, 3 * C, device=dev) * scale_in).half()
    q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    out = torch.zeros(nb * batch * seq, C, device=dev, dtype=torch.float16)
    if nv == 1:
        ops.attention(q, k, v, heads, seq, nb * batch, out)
        ref = ref_attn(q.reshape(batch, seq, C), k.reshape(batch, seq, C), v.reshape(batch, seq, C), heads)
        ok = report(f"attn rows nv=1 b={batch} h={heads} seq={seq} x{scale_in}", out.view(batch, seq, C), ref)
    else:
        rows = batch * seq
        ops.attention(q[:rows], k[:rows], v, heads, seq, batch, out, n_v=3, v_branch_stride=rows * 3 * C,
                      o_branch_stride=rows * C)
        qs = q[:rows].reshape(batch, seq, C)
        ks = k[:rows].reshape(batch, seq, C)
        refs = [ref_attn(qs, ks, v[i * rows:(i + 1) * rows].reshape(batch, seq, C), heads) for i in range(3)]
        ref = torch.stack(refs).reshape(3 * batch, seq, C)
        ok = report(f"attn rows nv=3 b={batch} h={heads} seq={seq} x{scale_in}", out.view(3 * batch, seq, C), ref)
    if timing:
        if nv == 1:
            fn = lambda: ops.attention(q, k, v, heads, seq, nb * batch, out)
            flops = 4.0 * batch * heads * seq * seq * 64
        else:
            rows = batch * seq
            fn = lambda: ops.attention(q[:rows], k[:rows], v, heads, seq, batch, out, n_v=3,
                                       v_branch_stride=rows * 3 * C, o_branch_stride=rows * C)
            flops = 2.0 * batch * heads * seq * seq * 64 * (1 + 3)
        t = timeit(fn, iters=10)
        print(f"   time {t*1e6:.1f} us -> {flops/t/1e12:.1f} TFLOP/s (executed flops)", flush=True)
        if nv == 1:
            qh = q.reshape(batch, seq, heads, 64).transpose(1, 2)
            kh = k.reshape(batch, seq, heads, 64).transpose(1, 2)
            vh = v.reshape(batch, seq, heads, 64).transpose(1, 2)
            tt = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(qh, kh, vh), iters=10)
            print(f"   torch SDPA {tt*1e6:.1f} us -> {flops/tt/1e12:.1f} TFLOP/s", flush=True)
    return ok


def frames_case(clips, heads, F, HW, nv, timing=False):
    torch.manual_seed(2)
    C = heads * 64
    nb = 3 if nv == 3 else 1
    x = torch.randn(nb * clips * F * HW, 3 * C, device=dev).half()
    q, k, v = x[:, :C], x[:, C:2 * C], x[:, 2 * C:]
    out = torch.zeros(nb * clips * F * HW, C, device=dev, dtype=torch.float16)

    def to_seq(t, n):  # [n*F*HW, C] frame-major -> [n*HW, F, C]
        return t.reshape(n, F, HW, C).permute(0, 2, 1, 3).reshape(n * HW, F, C)

    def from_seq(t, n):
        return t.reshape(n, HW, F, C).permute(0, 2, 1, 3).reshape(n * F * HW, C)

    if nv == 1:
        ops.attention(q, k, v, heads, F, clips * HW, out, frames_mode=True, HW=HW)
        ref = from_seq(ref_attn(to_seq(q, clips), to_seq(k, clips), to_seq(v, clips), heads), clips)
        ok = report(f"attn frames nv=1 clips={clips} h={heads} F={F} HW={HW}", out, ref)
    else:
        rows = clips * F * HW
        ops.attention(q[:rows], k[:rows], v, heads, F, clips * HW, out, n_v=3, v_branch_stride=rows * 3 * C,
                      o_branch_stride=rows * C, frames_mode=True, HW=HW)
        refs = [from_seq(ref_attn(to_seq(q[:rows], clips), to_seq(k[:rows], clips),
                                  to_seq(v[i * rows:(i + 1) * rows], clips), heads), clips) for i in range(3)]
        ok = report(f"attn frames nv=3 clips={clips} h={heads} F={F} HW={HW}", out, torch.cat(refs))
    if timing:
        if nv == 1:
            fn = lambda: ops.attention(q, k, v, heads, F, clips * HW, out, frames_mode=True, HW=HW)
        else:
            rows = clips * F * HW
            fn = lambda: ops.attention(q[:rows], k[:rows], v, heads, F, clips * HW, out, n_v=3,
                                       v_branch_stride=rows * 3 * C, o_branch_stride=rows * C, frames_mode=True, HW=HW)
        t = timeit(fn, iters=10)
        byts = (q.numel() * (2 if nv == 1 else 2 / 3) + v.numel() + out.numel()) * 2
        print(f"   time {t*1e6:.1f} us -> {byts/t/1e9:.0f} GB/s (q,k,v,o bytes)", flush=True)
    return ok


def main(argv):
    stage = argv[0] if argv else "all"
    if stage in ("all", "small"):
        rows_case(1, 1, 128, 1)
        rows_case(2, 2, 256, 1)
        rows_case(1, 2, 256, 3)
        rows_case(2, 2, 1024, 1, scale_in=3.0)
        rows_case(1, 1, 200, 1)
        rows_case(1, 2, 880, 3)
        frames_case(1, 1, 16, 64, 1)
        frames_case(2, 2, 16, 64, 3)
        frames_case(1, 2, 8, 256, 1)
        frames_case(1, 1, 128, 16, 1)
        frames_case(1, 1, 256, 8, 1)
    if stage in ("all", "big"):
        rows_case(48, 5, 4096, 1, timing=True)
        rows_case(16, 5, 4096, 3, timing=True)
        rows_case(48, 10, 1024, 1, timing=True)
        rows_case(16, 20, 256, 3, timing=True)
        frames_case(3, 5, 16, 4096, 1, timing=True)
        frames_case(1, 5, 16, 4096, 3, timing=True)

"""TEST INFRASTRUCTURE — CPU restatement of the CONTRACT of every C-ABI op in anyv2v_b200.ops (include/anyv2v_b200.h).

The product has no CPU path (ops.* raise on CPU tensors).  For the `-m "not gpu"` tests this module states what each
kernel is specified to compute — same arguments, same layouts (channels-last activations, packed weights, branch slots,
row-strided views), exact (float64) arithmetic on the fp16 inputs, one rounding to fp16 at the store — in plain PyTorch
(float64 rather than the kernels' fp32 accumulation makes the emulation independent of how a batch is split, so the
de-duplication / pruning switches can be checked for bit-identical results), and the
``emulated_ops`` fixture (tests/conftest.py) swaps it in for the duration of ONE test.  That lets the host logic that sits
on top of the kernels (the channels-last UNet wiring, the PnP de-duplication, hooks, loops, latent store) run on CPU and
be compared with the oracle.  The kernels themselves are checked against the same fp32 formulas on the GPU
(tests/test_gpu_kernels.py); nothing outside tests/ imports this file.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

_launches = 0


def _count(n=1):
    global _launches
    _launches += n


def launch_count():
    return _launches


def _f16(t, name):
    assert t.dtype == torch.float16, f"{name}: the C ABI takes fp16 tensors, got {t.dtype}"


def _store(out, y32, shape=None):
    y16 = y32.to(torch.float16).contiguous()  # the kernels write freshly allocated, contiguous outputs
    if out is None:
        return y16 if shape is None else y16.view(shape)
    out.copy_(y16.view(out.shape))
    return out


# ------------------------------------------------------------------------------------------------------------- K7
def ddim_step(x, v_neg, v_edit, guidance, ca, cb, cc, cd, out=None, inverse=False, coef_dev=None):
    """csrc/elementwise.cu ddim_one: every product / sum rounded to fp16 separately, fp32 scalars."""
    _f16(x, "ddim.x")
    if coef_dev is not None:
        ca, cb, cc, cd, guidance = (float(v) for v in coef_dev[:5].tolist())
    f32 = lambda s: torch.tensor(s, dtype=torch.float32)
    r16 = lambda t: t.to(torch.float16).to(torch.float32)
    xf, vn = x.float(), v_neg.float()
    v = vn
    if v_edit is not None:
        d0 = r16(v_edit.float() - vn)
        d1 = r16(f32(guidance) * d0)
        v = r16(vn + d1)
    x0 = r16(r16(f32(ca) * xf) - r16(f32(cb) * v))
    ep = r16(r16(f32(ca) * v) + r16(f32(cb) * xf))
    direction = r16(f32(cd) * ep)
    res = r16(r16(f32(cc) * x0) + direction)
    _count()
    return _store(out, res, x.shape)


# ------------------------------------------------------------------------------------------------------------- K6
def groupnorm(x, gamma, beta, groups, eps, silu, out=None, x2=None):
    _f16(x, "groupnorm.x")
    assert x.dim() == 3 and x.is_contiguous()
    if x2 is not None:  # two-source input: the logical tensor is [x | x2] along the channels
        _f16(x2, "groupnorm.x2")
        assert x2.is_contiguous() and x2.shape[:2] == x.shape[:2] and x.shape[2] % 8 == 0
        x = torch.cat([x, x2], dim=2)
    n, rows, C = x.shape
    xf = x.double().view(n, rows, groups, C // groups)
    mean = xf.mean(dim=(1, 3), keepdim=True)
    var = xf.var(dim=(1, 3), unbiased=False, keepdim=True)
    y = ((xf - mean) * torch.rsqrt(var + eps)).view(n, rows, C) * gamma.double() + beta.double()
    if silu:
        y = y.to(torch.float16).double()  # the reference rounds the GroupNorm output before SiLU (two ops)
        y = y * torch.sigmoid(y)
    _count(1)
    return _store(out, y, x.shape)


def layernorm(x, gamma, beta, eps=1e-5, out=None):
    _f16(x, "layernorm.x")
    assert x.is_contiguous()
    y = F.layer_norm(x.double(), (x.shape[-1],), gamma.double(), beta.double(), eps)
    _count()
    return _store(out, y, x.shape)


# ------------------------------------------------------------------------------------------------------------- GEMM
def geglu_pack(w, bias):
    n2, k = w.shape
    inner = n2 // 2
    assert inner % 32 == 0
    wp = torch.stack([w[:inner].view(inner // 32, 32, k), w[inner:].view(inner // 32, 32, k)], dim=1).reshape(n2, k)
    bp = torch.stack([bias[:inner].view(inner // 32, 32), bias[inner:].view(inner // 32, 32)], dim=1).reshape(n2)
    return wp.contiguous(), bp.contiguous()


def _epilogue(y, bias, rowbias, rows_per_rowbias, residual2d):
    if bias is not None:
        y = y + bias.double()
    if rowbias is not None:
        idx = torch.arange(y.shape[0]) // rows_per_rowbias
        y = y + rowbias.double()[idx]
    if residual2d is not None:
        y = y + residual2d.double()
    return y


def linear(a, w, bias=None, residual=None, out=None, rowbias=None, rows_per_rowbias=0, geglu=False, a2=None):
    _f16(a, "linear.a")
    if a2 is not None:  # two-source K loop: the logical A is [a | a2]
        _f16(a2, "linear.a2")
        assert a.shape[1] % 64 == 0 and a2.shape[0] == a.shape[0]
        a = torch.cat([a, a2], dim=1)
    assert a.dim() == 2 and a.stride(1) == 1 and w.is_contiguous() and w.shape[1] == a.shape[1]
    M, N = a.shape[0], w.shape[0]
    y = a.double() @ w.double().t()
    if geglu:
        assert residual is None and rowbias is None and N % 64 == 0
        y = _epilogue(y, bias, None, 0, None).view(M, N // 64, 2, 32)
        y = (y[:, :, 0] * F.gelu(y[:, :, 1])).reshape(M, N // 2)  # exact (erf) GELU, one rounding at the store
    else:
        y = _epilogue(y, bias, rowbias, rows_per_rowbias, residual)
    _count()
    if out is None:
        return y.to(torch.float16).contiguous()
    assert out.stride(1) == 1
    out.copy_(y.to(torch.float16))
    return out


def conv3x3(x, w_packed, bias=None, rowbias=None, rows_per_rowbias=0, residual=None, out=None, n_slots=1, slot_stride=0, stride=1):
    _f16(x, "conv3x3.x")
    assert x.dim() == 4 and x.is_contiguous()
    NF, H, W, C = x.shape
    Cout = w_packed.shape[0]
    Cin = w_packed.shape[1] // 9
    assert w_packed.shape[1] == 9 * Cin and C <= Cin and H % stride == 0 and W % stride == 0
    w = w_packed.double().view(Cout, 3, 3, Cin).permute(0, 3, 1, 2)[:, :C]  # [Cout][ky][kx][Cin] -> OIHW; padded channels read zeros
    H, W = H // stride, W // stride
    y = F.conv2d(x.double().permute(0, 3, 1, 2), w, None, stride=stride, padding=1).permute(0, 2, 3, 1).reshape(NF * H * W, Cout)
    y = _epilogue(y, bias, rowbias, rows_per_rowbias, None)
    M = NF * H * W
    _count()
    if out is None:
        assert n_slots == 1
        if residual is not None:
            y = y + residual.double().reshape(M, Cout)
        return y.to(torch.float16).contiguous().view(NF, H, W, Cout)
    flat = out.view(-1)
    rflat = None if residual is None else residual.reshape(-1)
    for s in range(n_slots):  # one accumulator tile, n_slots stores (+ each slot's own residual): fused PnP injection
        lo = s * slot_stride
        ys = y if rflat is None else y + rflat[lo:lo + M * Cout].double().view(M, Cout)
        flat[lo:lo + M * Cout] = ys.to(torch.float16).reshape(-1)
    return out


def upsample2x_conv3x3(x, w_phases, bias=None, out=None):
    """the four 2 x 2 phase convolutions of nearest-up x 2 + conv 3 x 3, interleaved into the [NF, 2H, 2W, Cout] image"""
    _f16(x, "upsample2x_conv3x3.x")
    NF, H, W, Cin = x.shape
    Cout = w_phases.shape[1]
    xp = F.pad(x.double().permute(0, 3, 1, 2), (1, 1, 1, 1))  # zero border: input offsets -1 .. +1
    y = torch.zeros(NF, 2 * H, 2 * W, Cout, dtype=torch.float64)
    for ph in range(4):
        py, px = ph >> 1, ph & 1
        w = w_phases[ph].double().view(Cout, 2, 2, Cin).permute(0, 3, 1, 2)  # OIHW, tap (a, b) reads input (i + a - 1 + py, j + b - 1 + px)
        acc = F.conv2d(xp[:, :, py:py + H + 1, px:px + W + 1], w, None if bias is None else bias.double())  # [NF, Cout, H, W]
        y[:, py::2, px::2, :] = acc.permute(0, 2, 3, 1)
        _count()
    return _store(out, y, (NF, 2 * H, 2 * W, Cout))


def tconv3(x, w_packed, F_, HW, bias=None, residual=None, out=None):
    _f16(x, "tconv3.x")
    assert x.dim() == 3 and x.is_contiguous() and x.shape[1] == F_ * HW
    B, R, Cin = x.shape
    Cout = w_packed.shape[0]
    w = w_packed.double().view(Cout, 3, Cin).permute(0, 2, 1)[:, :, :, None, None]  # [Cout][kt][Cin] -> [O, I, kt, 1, 1]
    x5 = x.double().view(B, F_, HW, 1, Cin).permute(0, 4, 1, 2, 3)
    y = F.conv3d(x5, w, None, padding=(1, 0, 0)).permute(0, 2, 3, 4, 1).reshape(B * R, Cout)
    y = _epilogue(y, bias, None, 0, None if residual is None else residual.reshape(B * R, Cout))
    _count()
    return _store(out, y, (B, R, Cout))


# ------------------------------------------------------------------------------------------------------------- attention
def attention(q, k, v, heads, seq, batch, out, scale=0.125, n_v=1, v_branch_stride=0, o_branch_stride=0,
              frames_mode=False, HW=0, seq_kv=0, kv_batch_div=0):
    for name, t in (("q", q), ("k", k), ("v", v), ("o", out)):
        _f16(t, "attention." + name)
        assert t.dim() == 2 and t.stride(1) == 1
    C = heads * 64
    ldv, ldo = v.stride(0), out.stride(0)

    def sdpa(qh, kh, vh):  # [..., L, heads, 64]
        p = torch.softmax(torch.einsum("...qhd,...khd->...hqk", qh.double(), kh.double()) * scale, dim=-1)
        return torch.einsum("...hqk,...khd->...qhd", p, vh.double())

    branches = range(n_v)
    if not frames_mode:
        div = kv_batch_div if kv_batch_div > 0 else 1
        nk = seq_kv if seq_kv > 0 else seq
        kvb = batch // div
        qh = q[:batch * seq, :C].reshape(batch, seq, heads, 64)
        kh = k[:kvb * nk, :C].reshape(kvb, nk, heads, 64).repeat_interleave(div, dim=0)
        vrows = v_branch_stride // ldv if n_v == 3 else 0
        orows = o_branch_stride // ldo if n_v == 3 else 0
        for b in branches:
            vh = v[b * vrows:b * vrows + kvb * nk, :C].reshape(kvb, nk, heads, 64).repeat_interleave(div, dim=0)
            o = sdpa(qh, kh, vh).reshape(batch * seq, C)
            out[b * orows:b * orows + batch * seq, :C] = o.to(torch.float16)
    else:
        assert batch % HW == 0 and (seq_kv <= 0 or seq_kv == seq) and kv_batch_div <= 1
        clips, Fr = batch // HW, seq
        rows = clips * Fr * HW
        to_seq = lambda t: t.reshape(clips, Fr, HW, heads, 64).permute(0, 2, 1, 3, 4)  # [clips, HW, F, heads, 64]
        qh, kh = to_seq(q[:rows, :C]), to_seq(k[:rows, :C])
        vrows = v_branch_stride // ldv if n_v == 3 else 0
        orows = o_branch_stride // ldo if n_v == 3 else 0
        for b in branches:
            vh = to_seq(v[b * vrows:b * vrows + rows, :C])
            o = sdpa(qh, kh, vh).permute(0, 2, 1, 3, 4).reshape(rows, C)  # back to frame-major tokens
            out[b * orows:b * orows + rows, :C] = o.to(torch.float16)
    _count()
    return out


def temporal_attention_fused(x, wqkv, heads, F_, HW, clips, out, scale=0.125, n_v=1):
    """Q/K/V projection (rounded to fp16, as the QKV GEMM would store them) + frames-mode attention; n_v = 3: Q, K of every
    clip from the source clip of the same index (clips ordered [source | uncond | cond])"""
    _f16(x, "temporal_attention_fused.x")
    C = heads * 64
    qkv = (x.double() @ wqkv.double().t()).to(torch.float16)
    _count(0)
    if n_v == 1:
        return attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], heads, F_, clips * HW, out, scale=scale, frames_mode=True, HW=HW)
    assert n_v == 3 and clips % 3 == 0
    src_rows = (clips // 3) * F_ * HW
    return attention(qkv[:src_rows, :C], qkv[:src_rows, C:2 * C], qkv[:, 2 * C:], heads, F_, (clips // 3) * HW, out, scale=scale, n_v=3,
                     v_branch_stride=src_rows * qkv.stride(0), o_branch_stride=src_rows * out.stride(0), frames_mode=True, HW=HW)


CONTRACTS = dict(upsample2x_conv3x3=upsample2x_conv3x3, temporal_attention_fused=temporal_attention_fused, ddim_step=ddim_step, groupnorm=groupnorm, layernorm=layernorm, geglu_pack=geglu_pack, linear=linear,
                 conv3x3=conv3x3, tconv3=tconv3, attention=attention, launch_count=launch_count)

"""Tensor-level wrappers over the C ABI: PyTorch is used for device memory and streams only.

Every function enqueues hand-written sm_100a kernels on the current CUDA stream; none has a PyTorch fallback.
Activations are channels-last (see include/anyv2v_b200.h).  ``launch_count()`` reports how many of OUR kernels
were launched (bench.py's ``gpu_launches``).
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib as L

_launches = 0
_gn_ws: dict = {}
_gn_ws_keepalive: list = []


def launch_count() -> int:
    return _launches


# -- optional in-situ timing (development aid): CUDA events around every wrapper call, keyed by op + shape
_prof = None


def profile_begin():
    global _prof
    _prof = []


def profile_end():
    """-> {key: (calls, total_ms)} sorted by time; synchronises."""
    global _prof
    torch.cuda.synchronize()
    out = {}
    for key, e0, e1 in _prof or []:
        c, t = out.get(key, (0, 0.0))
        out[key] = (c + 1, t + e0.elapsed_time(e1))
    _prof = None
    return dict(sorted(out.items(), key=lambda kv: -kv[1][1]))


class _timed:
    def __init__(self, key):
        self.key = key

    def __enter__(self):
        if _prof is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if _prof is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _prof.append((self.key, self.e0, e1))


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _f16_cuda(t: torch.Tensor, name: str) -> None:
    if not (t.is_cuda and t.dtype == torch.float16):
        raise L.Av2vError(f"{name}: expected a CUDA fp16 tensor, got {t.device}/{t.dtype} (no CPU fallback exists)")


# ----------------------------------------------------------------------------------------------------------- K7
def ddim_step(x, v_neg, v_edit, guidance: float, ca: float, cb: float, cc: float, cd: float, out=None,
              inverse: bool = False, coef_dev=None):
    """Fused CFG + v-prediction DDIM update (pipeline_i2vgen_xl.py:1159-1176 / :1407-1420). Elementwise."""
    global _launches
    _f16_cuda(x, "ddim_step.x")
    _f16_cuda(v_neg, "ddim_step.v_neg")
    assert x.is_contiguous() and v_neg.is_contiguous() and x.numel() == v_neg.numel()
    if v_edit is not None:
        _f16_cuda(v_edit, "ddim_step.v_edit")
        assert v_edit.is_contiguous() and v_edit.numel() == x.numel()
    if out is None:
        out = torch.empty_like(x)
    if coef_dev is not None:
        assert coef_dev.is_cuda and coef_dev.dtype == torch.float32 and coef_dev.numel() >= 5
    a = L.DdimArgs(_p(x), _p(v_neg), _p(v_edit), _p(out), x.numel(), guidance, ca, cb, cc, cd, _p(coef_dev))
    fn = L.lib().av2v_ddim_inverse_step_f16 if inverse else L.lib().av2v_ddim_step_cfg_f16
    with _timed("ddim_step"):
        L.check(fn(ctypes.byref(a), _stream()), "av2v_ddim_step")
    _launches += 1
    return out


# ----------------------------------------------------------------------------------------------------------- K6
def groupnorm(x, gamma, beta, groups: int, eps: float, silu: bool, out=None, x2=None):
    """GroupNorm(+SiLU) over x[n_samples, rows, C] (channels-last). pnp_utils.py:48-49,92,104.
    x2: second source — the logical input is [x | x2] along the channels (skip-concat without torch.cat); out is [n, rows, C1 + C2]."""
    global _launches
    _f16_cuda(x, "groupnorm.x")
    assert x.dim() == 3 and x.is_contiguous()
    n, rows, C = x.shape
    C1 = C
    if x2 is not None:
        _f16_cuda(x2, "groupnorm.x2")
        assert x2.dim() == 3 and x2.is_contiguous() and x2.shape[:2] == x.shape[:2]
        C = C1 + x2.shape[2]
    if out is None:
        out = torch.empty((n, rows, C), dtype=x.dtype, device=x.device)
    need = L.lib().av2v_groupnorm_workspace_floats(n, C)
    key = x.device.index
    ws = _gn_ws.get(key)
    if ws is None or ws.numel() < need:
        # never free an old workspace: a captured CUDA graph may still hold its address
        _gn_ws_keepalive.append(ws)
        ws = torch.empty(max(need, 1 << 22), dtype=torch.float32, device=x.device)
        _gn_ws[key] = ws
    a = L.GroupNormArgs(_p(x), _p(out), _p(gamma), _p(beta), _p(ws), n, rows, C, groups, eps, 1 if silu else 0, _p(x2), C1)
    with _timed(f"groupnorm n={n} rows={rows} C={C}"):
        L.check(L.lib().av2v_groupnorm_silu_f16(ctypes.byref(a), _stream()), "av2v_groupnorm_silu_f16")
    _launches += 1
    return out


# ----------------------------------------------------------------------------------------------------------- GEMM
def _gemm(args: L.GemmArgs):
    global _launches
    kind = ("linear", "conv3x3", "tconv3")[args.mode] + ("+geglu" if args.geglu else "") + ("+res" if args.residual else "")
    with _timed(f"{kind} M={args.M} N={args.N} K={args.K} slots={args.n_slots}"):
        L.check(L.lib().av2v_gemm_f16(ctypes.byref(args), _stream()), "av2v_gemm_f16")
    _launches += 1


def linear(a, w, bias=None, residual=None, out=None, rowbias=None, rows_per_rowbias: int = 0, geglu: bool = False, a2=None):
    """out[M,N] = a[M,K] @ w[N,K]^T (+bias) (+rowbias[m//rpr]) (+residual). a may be a row-strided view.
    geglu=True: w/bias are block-32 interleaved [h|gate] (see geglu_pack) and out is [M, N/2] = h * gelu_erf(gate).
    a2: second source of the K loop — the logical A is [a | a2] along K (a.shape[1] % 64 == 0): the skip-connection concat of
    the up blocks without a materialised torch.cat."""
    _f16_cuda(a, "linear.a")
    assert a.dim() == 2 and a.stride(1) == 1 and w.is_contiguous()
    M, K = a.shape
    if a2 is not None:
        _f16_cuda(a2, "linear.a2")
        assert a2.dim() == 2 and a2.stride(1) == 1 and a2.shape[0] == M and K % 64 == 0
        K = K + a2.shape[1]
    N = w.shape[0]
    assert w.shape[1] == K
    if out is None:
        out = torch.empty((M, N // 2 if geglu else N), dtype=torch.float16, device=a.device)
    assert out.stride(1) == 1
    if residual is not None:
        assert residual.stride(1) == 1 and residual.stride(0) == out.stride(0)
    g = L.GemmArgs()
    g.mode = L.A_LINEAR
    g.a, g.w, g.M, g.N, g.K, g.lda = _p(a), _p(w), M, N, K, a.stride(0)
    g.bias, g.rowbias, g.rows_per_rowbias = _p(bias), _p(rowbias), rows_per_rowbias
    g.residual, g.out, g.ldo, g.n_slots, g.slot_stride = _p(residual), _p(out), out.stride(0), 1, 0
    g.geglu = 1 if geglu else 0
    if a2 is not None:
        g.a2, g.k_split, g.lda2 = _p(a2), a.shape[1], a2.stride(0)
    _gemm(g)
    return out


def geglu_pack(w, bias):
    """Interleave the h / gate halves of GEGLU.proj in blocks of 32 output features (layout of av2v_gemm_args.geglu)."""
    n2, k = w.shape
    inner = n2 // 2
    assert inner % 32 == 0
    wp = torch.stack([w[:inner].view(inner // 32, 32, k), w[inner:].view(inner // 32, 32, k)], dim=1).reshape(n2, k)
    bp = torch.stack([bias[:inner].view(inner // 32, 32), bias[inner:].view(inner // 32, 32)], dim=1).reshape(n2)
    return wp.contiguous(), bp.contiguous()


def layernorm(x, gamma, beta, eps: float = 1e-5, out=None):
    """LayerNorm over the last dim of a contiguous [..., C] tensor."""
    global _launches
    _f16_cuda(x, "layernorm.x")
    assert x.is_contiguous()
    C = x.shape[-1]
    rows = x.numel() // C
    if out is None:
        out = torch.empty_like(x)
    a = L.LayerNormArgs(_p(x), _p(out), _p(gamma), _p(beta), rows, C, eps)
    with _timed(f"layernorm rows={rows} C={C}"):
        L.check(L.lib().av2v_layernorm_f16(ctypes.byref(a), _stream()), "av2v_layernorm_f16")
    _launches += 1
    return out


def conv3x3(x, w_packed, bias=None, rowbias=None, rows_per_rowbias: int = 0, residual=None, out=None,
            n_slots: int = 1, slot_stride: int = 0, stride: int = 1):
    """3x3 / pad 1 convolution as an implicit GEMM. x: [NF,H,W,C] contiguous (channels-last),
    w_packed: [Cout, 9*Cin] (= conv.weight.permute(0,2,3,1).reshape). out: [n_slots][NF*(H/stride)*(W/stride), Cout].
    stride 2 = Downsample2D (the taps are sampled with TMA element strides).  C < Cin (conv_in: 8 channels): the weights are
    zero-padded per tap to Cin = 64 and the missing channels of every K block read as zeros (TMA out-of-bounds fill)."""
    _f16_cuda(x, "conv3x3.x")
    assert x.dim() == 4 and x.is_contiguous()
    NF, H, W, C = x.shape
    Cout = w_packed.shape[0]
    assert w_packed.shape[1] % 9 == 0 and w_packed.is_contiguous()
    Cin = w_packed.shape[1] // 9
    assert C <= Cin and C % 8 == 0
    M = NF * (H // stride) * (W // stride)
    if out is None:
        assert n_slots == 1
        out = torch.empty((NF, H // stride, W // stride, Cout), dtype=torch.float16, device=x.device)
    g = L.GemmArgs()
    g.mode = L.A_CONV3X3
    g.a, g.w, g.M, g.N, g.K = _p(x), _p(w_packed), M, Cout, 9 * Cin
    g.NF, g.H, g.W, g.Cin = NF, H, W, Cin
    g.stride, g.a_channels = stride, (C if C != Cin else 0)
    g.bias, g.rowbias, g.rows_per_rowbias = _p(bias), _p(rowbias), rows_per_rowbias
    g.residual, g.out, g.ldo, g.n_slots, g.slot_stride = _p(residual), _p(out), Cout, n_slots, slot_stride
    _gemm(g)
    return out


def upsample2x_conv3x3(x, w_phases, bias=None, out=None):
    """Upsample2D = nearest-neighbour x 2 followed by a 3x3 / pad 1 convolution (twin at seine/models/resnet.py:24-76)
    WITHOUT the up-sampled tensor: output pixel (2i+py, 2j+px) only sees a 2 x 2 neighbourhood of the input, with the 3 x 3 taps
    that land on the same input pixel pre-summed (``pack_upsample_weights``).  x: [NF,H,W,Cin]; w_phases: [4][Cout, 4*Cin];
    -> [NF, 2H, 2W, Cout].  Four launches of the implicit GEMM (K = 4*Cin each): 4/9 of the direct layer's FLOPs, and the 4x
    larger intermediate is neither written nor read."""
    _f16_cuda(x, "upsample2x_conv3x3.x")
    assert x.dim() == 4 and x.is_contiguous() and w_phases.dim() == 3 and w_phases.shape[0] == 4 and w_phases.is_contiguous()
    NF, H, W, Cin = x.shape
    Cout = w_phases.shape[1]
    assert w_phases.shape[2] == 4 * Cin
    if out is None:
        out = torch.empty((NF, 2 * H, 2 * W, Cout), dtype=torch.float16, device=x.device)
    for ph in range(4):
        g = L.GemmArgs()
        g.mode = L.A_CONV3X3
        g.a, g.w, g.M, g.N, g.K = _p(x), _p(w_phases[ph]), NF * H * W, Cout, 4 * Cin
        g.NF, g.H, g.W, g.Cin = NF, H, W, Cin
        g.bias, g.out, g.ldo, g.n_slots, g.slot_stride = _p(bias), _p(out), Cout, 1, 0
        g.up2_phase = ph + 1
        _gemm(g)
    return out


def pack_upsample_weights(w):
    """conv weight [Cout, Cin, 3, 3] -> [4][Cout, 2*2*Cin]: for output phase (py, px) the tap (a, b) of the 2 x 2 neighbourhood
    (input offsets a - 1 + py, b - 1 + px) carries the sum of the 3 x 3 taps that read the same input pixel after the nearest
    up-sampling: rows {0 | 1,2} for py = 0, {0,1 | 2} for py = 1 (columns alike).  Summed in fp32, rounded to fp16 once."""
    sets = {0: ((0,), (1, 2)), 1: ((0, 1), (2,))}
    wf = w.float()
    co, ci = w.shape[0], w.shape[1]
    out = torch.empty((4, co, 2, 2, ci), dtype=torch.float32, device=w.device)
    for py in (0, 1):
        for px in (0, 1):
            for a in (0, 1):
                for b in (0, 1):
                    acc = 0
                    for ky in sets[py][a]:
                        for kx in sets[px][b]:
                            acc = acc + wf[:, :, ky, kx]
                    out[py * 2 + px, :, a, b, :] = acc
    return out.reshape(4, co, 4 * ci).to(w.dtype).contiguous()


def tconv3(x, w_packed, F: int, HW: int, bias=None, residual=None, out=None):
    """Conv3d (3,1,1) / pad (1,0,0) over frames. x: [B, F*HW, Cin] contiguous (frame-major channels-last),
    w_packed: [Cout, 3*Cin] (= conv.weight[:, :, :, 0, 0].permute(0,2,1).reshape)."""
    _f16_cuda(x, "tconv3.x")
    assert x.dim() == 3 and x.is_contiguous() and x.shape[1] == F * HW
    B, R, Cin = x.shape
    Cout = w_packed.shape[0]
    assert w_packed.shape[1] == 3 * Cin and w_packed.is_contiguous()
    if out is None:
        out = torch.empty((B, R, Cout), dtype=torch.float16, device=x.device)
    g = L.GemmArgs()
    g.mode = L.A_TCONV3
    g.a, g.w, g.M, g.N, g.K = _p(x), _p(w_packed), B * R, Cout, 3 * Cin
    g.B, g.rows_per_clip, g.HW, g.Cin = B, R, HW, Cin
    g.bias, g.residual, g.out, g.ldo, g.n_slots, g.slot_stride = _p(bias), _p(residual), _p(out), Cout, 1, 0
    _gemm(g)
    return out


# ----------------------------------------------------------------------------------------------------------- attention
def attention(q, k, v, heads: int, seq: int, batch: int, out, scale: float = 0.125, n_v: int = 1,
              v_branch_stride: int = 0, o_branch_stride: int = 0, frames_mode: bool = False, HW: int = 0,
              seq_kv: int = 0, kv_batch_div: int = 0):
    """PnP self-attention core (pnp_utils.py:189-210 / 295-316). q,k,v,out: 2-D token matrices (row-strided views ok)."""
    global _launches
    for name, t in (("q", q), ("k", k), ("v", v), ("o", out)):
        _f16_cuda(t, "attention." + name)
        assert t.dim() == 2 and t.stride(1) == 1
    a = L.AttnArgs()
    a.seq_mode = L.SEQ_FRAMES if frames_mode else L.SEQ_ROWS
    a.q, a.k, a.v, a.o = _p(q), _p(k), _p(v), _p(out)
    a.ldq, a.ldk, a.ldv, a.ldo = q.stride(0), k.stride(0), v.stride(0), out.stride(0)
    a.batch, a.seq, a.heads, a.HW, a.n_v = batch, seq, heads, HW, n_v
    a.v_branch_stride, a.o_branch_stride, a.scale = v_branch_stride, o_branch_stride, scale
    a.seq_kv, a.kv_batch_div = seq_kv, kv_batch_div
    with _timed(f"attention {'frames' if frames_mode else 'rows'} nv={n_v} batch={batch} seq={seq} heads={heads}"):
        L.check(L.lib().av2v_attn_pnp_f16(ctypes.byref(a), _stream()), "av2v_attn_pnp_f16")
    _launches += 1
    return out


def temporal_attention_fused(x, wqkv, heads: int, F: int, HW: int, clips: int, out, scale: float = 0.125, n_v: int = 1):
    """Temporal self-attention with the Q/K/V projection fused in (csrc/attention_tfused_tcgen05.cu; pnp_utils.py:247-334).
    x: frame-major tokens [clips*F*HW, Cx]; wqkv: [3*heads*64, Cx]; out: [clips*F*HW, heads*64].  n_v = 3: PnP-injected step,
    clips ordered [source | uncond | cond]; Q, K of every clip come from the source clip of the same index (pnp_utils.py:295-302)."""
    global _launches
    for name, t in (("x", x), ("wqkv", wqkv), ("o", out)):
        _f16_cuda(t, "temporal_attention_fused." + name)
        assert t.dim() == 2 and t.stride(1) == 1
    assert wqkv.is_contiguous() and wqkv.shape[0] == 3 * heads * 64 and wqkv.shape[1] == x.shape[1]
    assert x.shape[0] == clips * F * HW == out.shape[0]
    a = L.TAttnFusedArgs(_p(x), _p(wqkv), _p(out), x.stride(0), out.stride(0), clips, F, HW, heads, x.shape[1], scale, n_v)
    with _timed(f"temporal attention fused nv={n_v} clips={clips} F={F} HW={HW} heads={heads} Cx={x.shape[1]}"):
        L.check(L.lib().av2v_tattn_fused_f16(ctypes.byref(a), _stream()), "av2v_tattn_fused_f16")
    _launches += 1
    return out

"""Library-backed ops for what is NOT on the denoising step's hot path: the once-per-step embedding MLPs' activations
(a few KB), the image-latent stem's 2-head x dim-4 attention, the VAE's stem convolutions / mid attention / upsampling
(once per clip, SURVEY.md 8(f) "next").

Everything here is a PyTorch / cuDNN call; every op of the UNet's blocks (convs, norms, attention, feed-forward, skip
concats, up/down-sampling) goes through anyv2v_b200.ops and the C-ABI only.  There is still no CPU path: callers hand in
CUDA tensors.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def layer_norm(x, weight, bias, eps: float = 1e-5):
    return F.layer_norm(x, (x.shape[-1],), weight, bias, eps)


def silu(x):
    return F.silu(x)


def conv2d_nhwc(x_nhwc, weight, bias, stride: int = 1, padding: int = 1):
    """Small / strided convolutions (conv_in 8->320, conv_out 320->4, Downsample2D stride 2, image-latent stems)."""
    y = F.conv2d(x_nhwc.permute(0, 3, 1, 2), weight, bias, stride=stride, padding=padding)
    return y.permute(0, 2, 3, 1).contiguous()


def nearest_up2_nhwc(x_nhwc):
    """Upsample2D's F.interpolate(scale_factor=2, mode="nearest") in channels-last."""
    n, h, w, c = x_nhwc.shape
    return x_nhwc[:, :, None, :, None, :].expand(n, h, 2, w, 2, c).reshape(n, 2 * h, 2 * w, c)


def cross_attention(q, k, v, heads: int):
    """Cross-attention to the 145-token context (attn2 of the spatial transformers): q [B,Nq,C], k/v [B,Nk,C]."""
    b, nq, c = q.shape
    hd = c // heads
    qh = q.view(b, nq, heads, hd).transpose(1, 2)
    kh = k.view(b, -1, heads, hd).transpose(1, 2)
    vh = v.view(b, -1, heads, hd).transpose(1, 2)
    o = F.scaled_dot_product_attention(qh, kh, vh)
    return o.transpose(1, 2).reshape(b, nq, c)


def tiny_self_attention(q, k, v, heads: int):
    """image_latents_temporal_encoder: 2 heads x dim 4 — far below any tensor-core tile."""
    return cross_attention(q, k, v, heads)

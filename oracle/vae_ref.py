"""ORACLE (test infrastructure only) — AutoencoderKL restatement for the steps either side of the sampling loops
(SURVEY §8f row 4): `encode_vae_video` (i2vgen-xl/pipelines/pipeline_i2vgen_xl.py:565-592) and `decode_latents`
(:443-463).

PARITY UNPINNED: the arithmetic lives in diffusers==0.26.3 (`models/autoencoders/autoencoder_kl.py`, `vae.py`,
`unet_2d_blocks.py::{DownEncoderBlock2D, UpDecoderBlock2D, UNetMidBlock2D}`, `resnet.py::ResnetBlock2D(temb_channels=None)`,
`attention_processor.py::Attention(_from_deprecated_attn_block=True)`), which is absent here; the reference holds no VAE
tests or fixtures.  This file restates the published architecture of the Stable-Diffusion KL-f8 VAE the checkpoint
`ali-vilab/i2vgen-xl` ships (block_out_channels 128/256/512/512, 2 layers per block, 4 latent channels, GroupNorm(32,
eps 1e-6), one 512-wide single-head attention in each mid block, scaling_factor 0.18215) with diffusers' attribute
names, so a real state_dict loads unchanged.  Structural known answer: 83 653 863 parameters (the public
`diffusion_pytorch_model` of that VAE family), asserted in tests/test_oracle_kat.py.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

SD_VAE_CONFIG = dict(in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512),
                     layers_per_block=2, norm_num_groups=32, scaling_factor=0.18215)
TINY_VAE_CONFIG = dict(in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(64, 128), layers_per_block=1,
                       norm_num_groups=32, scaling_factor=0.18215)


class VaeResnetBlock2D(nn.Module):
    """ResnetBlock2D without a time embedding: GN -> SiLU -> conv1 -> GN -> SiLU -> conv2 -> shortcut + h."""

    def __init__(self, in_channels, out_channels, groups=32, eps=1e-6):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class VaeAttention(nn.Module):
    """The VAE's spatial self-attention block: GroupNorm -> q,k,v (bias) -> softmax(q k^T / sqrt(C)) v -> out proj ->
    + input (residual_connection=True, one head of width C, rescale_output_factor 1)."""

    def __init__(self, channels, groups=32, eps=1e-6):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, channels, eps=eps)
        self.to_q = nn.Linear(channels, channels)
        self.to_k = nn.Linear(channels, channels)
        self.to_v = nn.Linear(channels, channels)
        self.to_out = nn.ModuleList([nn.Linear(channels, channels), nn.Dropout(0.0)])
        self.heads = 1

    def forward(self, x):
        b, c, h, w = x.shape
        y = self.group_norm(x.view(b, c, h * w)).transpose(1, 2)  # [b, hw, c]
        q, k, v = self.to_q(y), self.to_k(y), self.to_v(y)
        o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
        o = self.to_out[0](o)
        return o.transpose(1, 2).reshape(b, c, h, w) + x


class _Down(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=0)

    def forward(self, x):  # diffusers Downsample2D(padding=0): asymmetric zero pad on the right / bottom
        return self.conv(F.pad(x, (0, 1, 0, 1)))


class _Up(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownEncoderBlock2D(nn.Module):
    def __init__(self, cin, cout, layers, groups, add_downsample):
        super().__init__()
        self.resnets = nn.ModuleList([VaeResnetBlock2D(cin if i == 0 else cout, cout, groups) for i in range(layers)])
        self.downsamplers = nn.ModuleList([_Down(cout)]) if add_downsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
        return x


class UpDecoderBlock2D(nn.Module):
    def __init__(self, cin, cout, layers, groups, add_upsample):
        super().__init__()
        self.resnets = nn.ModuleList([VaeResnetBlock2D(cin if i == 0 else cout, cout, groups) for i in range(layers)])
        self.upsamplers = nn.ModuleList([_Up(cout)]) if add_upsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class UNetMidBlock2D(nn.Module):
    def __init__(self, channels, groups):
        super().__init__()
        self.resnets = nn.ModuleList([VaeResnetBlock2D(channels, channels, groups), VaeResnetBlock2D(channels, channels, groups)])
        self.attentions = nn.ModuleList([VaeAttention(channels, groups)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class Encoder(nn.Module):
    def __init__(self, in_channels, latent_channels, chans: Sequence[int], layers, groups):
        super().__init__()
        self.conv_in = nn.Conv2d(in_channels, chans[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        c = chans[0]
        for i, co in enumerate(chans):
            self.down_blocks.append(DownEncoderBlock2D(c, co, layers, groups, add_downsample=i < len(chans) - 1))
            c = co
        self.mid_block = UNetMidBlock2D(c, groups)
        self.conv_norm_out = nn.GroupNorm(groups, c, eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(c, 2 * latent_channels, 3, padding=1)  # double_z: mean and log-variance

    def forward(self, x):
        x = self.conv_in(x)
        for blk in self.down_blocks:
            x = blk(x)
        x = self.mid_block(x)
        return self.conv_out(self.conv_act(self.conv_norm_out(x)))


class Decoder(nn.Module):
    def __init__(self, latent_channels, out_channels, chans: Sequence[int], layers, groups):
        super().__init__()
        rev = list(reversed(chans))
        self.conv_in = nn.Conv2d(latent_channels, rev[0], 3, padding=1)
        self.mid_block = UNetMidBlock2D(rev[0], groups)
        self.up_blocks = nn.ModuleList()
        c = rev[0]
        for i, co in enumerate(rev):
            self.up_blocks.append(UpDecoderBlock2D(c, co, layers + 1, groups, add_upsample=i < len(rev) - 1))
            c = co
        self.conv_norm_out = nn.GroupNorm(groups, c, eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(c, out_channels, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for blk in self.up_blocks:
            x = blk(x)
        return self.conv_out(self.conv_act(self.conv_norm_out(x)))


class DiagonalGaussianDistribution:
    """diffusers `DiagonalGaussianDistribution`: parameters = cat(mean, logvar) on dim 1, logvar clamped to [-30, 20]."""

    def __init__(self, parameters: torch.Tensor):
        self.mean, logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def sample(self, generator=None) -> torch.Tensor:
        noise = torch.randn(self.mean.shape, generator=generator, device=self.mean.device, dtype=self.mean.dtype)
        return self.mean + self.std * noise

    def mode(self) -> torch.Tensor:
        return self.mean


class AutoencoderKL(nn.Module):
    def __init__(self, in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512),
                 layers_per_block=2, norm_num_groups=32, scaling_factor=0.18215):
        super().__init__()
        self.encoder = Encoder(in_channels, latent_channels, block_out_channels, layers_per_block, norm_num_groups)
        self.decoder = Decoder(latent_channels, out_channels, block_out_channels, layers_per_block, norm_num_groups)
        self.quant_conv = nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1)
        self.post_quant_conv = nn.Conv2d(latent_channels, latent_channels, 1)
        self.config = SimpleNamespace(scaling_factor=scaling_factor, latent_channels=latent_channels,
                                      block_out_channels=tuple(block_out_channels))

    def encode(self, x):
        return SimpleNamespace(latent_dist=DiagonalGaussianDistribution(self.quant_conv(self.encoder(x))))

    def decode(self, z):
        return SimpleNamespace(sample=self.decoder(self.post_quant_conv(z)))


def seeded_vae(config: dict, seed: int = 8888, dtype=torch.float32) -> AutoencoderKL:
    g = torch.Generator().manual_seed(seed)
    vae = AutoencoderKL(**config)
    with torch.no_grad():
        for p in vae.parameters():  # deterministic init independent of the global RNG
            if p.dim() > 1:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) / fan_in ** 0.5)
            else:
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
        for m in vae.modules():
            if isinstance(m, nn.GroupNorm):
                m.weight.copy_(1.0 + 0.1 * torch.randn(m.weight.shape, generator=g))
    return vae.to(dtype).eval()


def decode_latents(vae, latents: torch.Tensor, decode_chunk_size=None) -> torch.Tensor:
    """pipeline_i2vgen_xl.py:443-463 — latents [b, c, f, h, w] -> video [b, 3, f, 8h, 8w] float32."""
    latents = 1 / vae.config.scaling_factor * latents
    b, c, f, h, w = latents.shape
    latents = latents.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
    if decode_chunk_size is not None:
        image = torch.cat([vae.decode(latents[i:i + decode_chunk_size]).sample for i in range(0, latents.shape[0], decode_chunk_size)], dim=0)
    else:
        image = vae.decode(latents).sample
    video = image[None, :].reshape((b, f, -1) + image.shape[2:]).permute(0, 2, 1, 3, 4)
    return video.float()


def encode_vae_video(vae, frames: torch.Tensor, generator=None) -> torch.Tensor:
    """pipeline_i2vgen_xl.py:565-592 after image pre-processing — frames [f, 3, H, W] in [-1, 1], one VAE call per frame,
    posterior SAMPLE scaled by scaling_factor -> [1, c, f, H/8, W/8]."""
    lat = []
    for i in range(frames.shape[0]):
        z = vae.encode(frames[i:i + 1]).latent_dist.sample(generator) * vae.config.scaling_factor
        lat.append(z.squeeze(0))
    lat = torch.stack(lat)
    return lat.reshape(1, frames.shape[0], *lat.shape[1:]).permute(0, 2, 1, 3, 4)

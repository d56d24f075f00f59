"""ORACLE (test infrastructure, not product code): plain-PyTorch restatement of diffusers==0.26.3 ``I2VGenXLUNet``.

PARITY UNPINNED at the diffusers boundary: diffusers is an un-vendored pip dependency of the reference
(i2vgen-xl/environment.yml:15) that is absent from /root/reference and from this image, and the reference ships no
tests or golden vectors.  This file restates the published architecture and is anchored on
what the reference repo itself pins:
  * attribute paths the hooks walk               — i2vgen-xl/pnp_utils.py:19-28, 130, 235-242, 340-346
  * the full ResnetBlock2D.forward arithmetic    — i2vgen-xl/pnp_utils.py:41-126
  * the full AttnProcessor2_0.__call__ arithmetic — i2vgen-xl/pnp_utils.py:142-228
  * Attention.__init__ (q/k/v no bias, out bias)  — consisti2v/consisti2v/models/videoldm_attention.py:64-177
  * BasicTransformerBlock structure               — consisti2v/consisti2v/models/videoldm_transformer_blocks.py:361-563
  * the UNet call signature                       — i2vgen-xl/pipelines/pipeline_i2vgen_xl.py:1146-1155
  * total parameter count 1.420 B                 — public fp16 checkpoint size (tests/test_oracle_unet.py)
Module / parameter names equal diffusers' so that the reference's UNMODIFIED pnp_utils.py can be executed on this
model through oracle/diffusers_shim.py.  Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline import it.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------- attention
class AttnProcessor2_0:
    """Default processor: q/k/v projection, SDPA (scale = head_dim**-0.5), out projection (pnp_utils.py:142-228
    without the injection block)."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, scale=1.0):
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        b = hidden_states.shape[0]
        q = attn.to_q(hidden_states)
        k = attn.to_k(ctx)
        v = attn.to_v(ctx)
        hd = k.shape[-1] // attn.heads
        q = q.view(b, -1, attn.heads, hd).transpose(1, 2)
        k = k.view(b, -1, attn.heads, hd).transpose(1, 2)
        v = v.view(b, -1, attn.heads, hd).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(b, -1, attn.heads * hd).to(q.dtype)
        o = attn.to_out[0](o)
        o = attn.to_out[1](o)
        return o / attn.rescale_output_factor


class Attention(nn.Module):
    """diffusers Attention with the attributes pnp_utils.py reads (videoldm_attention.py:64-177)."""

    def __init__(self, query_dim: int, cross_attention_dim: Optional[int] = None, heads: int = 8, dim_head: int = 64):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.spatial_norm = None
        self.group_norm = None
        self.norm_cross = None
        self.residual_connection = False
        self.rescale_output_factor = 1.0
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(cross_attention_dim or query_dim, inner, bias=False)
        self.to_v = nn.Linear(cross_attention_dim or query_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, bias=True), nn.Dropout(0.0)])
        self.processor = AttnProcessor2_0()

    def prepare_attention_mask(self, attention_mask, target_length, batch_size):  # never used on this path
        raise NotImplementedError("attention masks are not used by I2VGen-XL self/cross attention")

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kw):
        # dispatch through the INSTANCE attribute at call time: hooks assign module.processor = ... (pnp_utils.py:242)
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **kw)


class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)  # exact erf GELU


class GELU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out)

    def forward(self, x):
        return F.gelu(self.proj(x))


class FeedForward(nn.Module):
    def __init__(self, dim: int, mult: int = 4, activation_fn: str = "geglu", inner_dim: Optional[int] = None):
        super().__init__()
        inner = inner_dim or dim * mult
        act = GEGLU(dim, inner) if activation_fn == "geglu" else GELU(dim, inner)
        self.net = nn.ModuleList([act, nn.Dropout(0.0), nn.Linear(inner, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    """LN -> attn1(self) -> + ; LN -> attn2(cross | self when double_self_attention) -> + ; LN -> GEGLU FF -> +
    (videoldm_transformer_blocks.py:461-562)."""

    def __init__(self, dim, heads, head_dim, cross_attention_dim=None, double_self_attention=False):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-5)
        self.attn1 = Attention(dim, None, heads, head_dim)
        self.norm2 = nn.LayerNorm(dim, eps=1e-5)
        self.attn2 = Attention(dim, None if double_self_attention else cross_attention_dim, heads, head_dim)
        self.norm3 = nn.LayerNorm(dim, eps=1e-5)
        self.ff = FeedForward(dim, activation_fn="geglu")

    def forward(self, x, encoder_hidden_states=None):
        x = self.attn1(self.norm1(x), encoder_hidden_states=None) + x
        x = self.attn2(self.norm2(x), encoder_hidden_states=encoder_hidden_states) + x
        x = self.ff(self.norm3(x)) + x
        return x


class Transformer2DModel(nn.Module):
    def __init__(self, heads, head_dim, in_channels, cross_attention_dim, groups=32):
        super().__init__()
        inner = heads * head_dim
        self.norm = nn.GroupNorm(groups, in_channels, eps=1e-6)
        self.proj_in = nn.Linear(in_channels, inner)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(inner, heads, head_dim, cross_attention_dim)])
        self.proj_out = nn.Linear(inner, in_channels)

    def forward(self, x, encoder_hidden_states=None):
        bf, c, h, w = x.shape
        res = x
        y = self.norm(x).permute(0, 2, 3, 1).reshape(bf, h * w, c)
        y = self.proj_in(y)
        for blk in self.transformer_blocks:
            y = blk(y, encoder_hidden_states=encoder_hidden_states)
        y = self.proj_out(y)
        y = y.reshape(bf, h, w, c).permute(0, 3, 1, 2).contiguous()
        return y + res


class TransformerTemporalModel(nn.Module):
    def __init__(self, heads, head_dim, in_channels, groups=32):
        super().__init__()
        inner = heads * head_dim
        self.norm = nn.GroupNorm(groups, in_channels, eps=1e-6)
        self.proj_in = nn.Linear(in_channels, inner)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, heads, head_dim, None, double_self_attention=True)])
        self.proj_out = nn.Linear(inner, in_channels)

    def forward(self, x, num_frames: int):
        bf, c, h, w = x.shape
        b = bf // num_frames
        res = x
        y = x[None, :].reshape(b, num_frames, c, h, w).permute(0, 2, 1, 3, 4)
        y = self.norm(y)  # stats over (C/groups, F, h, w)
        y = y.permute(0, 3, 4, 2, 1).reshape(b * h * w, num_frames, c)  # batch-major: index = b*hw + y*w + x
        y = self.proj_in(y)
        for blk in self.transformer_blocks:
            y = blk(y, encoder_hidden_states=None)
        y = self.proj_out(y)
        y = y[None, None, :].reshape(b, h, w, num_frames, c).permute(0, 3, 4, 1, 2).contiguous()
        return y.reshape(bf, c, h, w) + res


# --------------------------------------------------------------------------------------------- conv blocks
class ResnetBlock2D(nn.Module):
    """GN -> SiLU -> conv1 -> +temb -> GN -> SiLU -> conv2 -> shortcut + h  (pnp_utils.py:41-126)."""

    def __init__(self, in_channels, out_channels, temb_channels, groups=32, eps=1e-5):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.nonlinearity = nn.SiLU()
        self.upsample = None
        self.downsample = None
        self.skip_time_act = False
        self.time_embedding_norm = "default"
        self.output_scale_factor = 1.0
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None

    def forward(self, input_tensor, temb, scale: float = 1.0):
        h = self.conv1(self.nonlinearity(self.norm1(input_tensor)))
        h = h + self.time_emb_proj(self.nonlinearity(temb))[:, :, None, None]
        h = self.conv2(self.dropout(self.nonlinearity(self.norm2(h))))
        if self.conv_shortcut is not None:
            input_tensor = self.conv_shortcut(input_tensor)
        return (input_tensor + h) / self.output_scale_factor


class TemporalConvLayer(nn.Module):
    def __init__(self, dim, groups=32, dropout=0.1):
        super().__init__()
        self.conv1 = nn.Sequential(nn.GroupNorm(groups, dim), nn.SiLU(), nn.Conv3d(dim, dim, (3, 1, 1), padding=(1, 0, 0)))
        for name in ("conv2", "conv3", "conv4"):
            setattr(self, name, nn.Sequential(nn.GroupNorm(groups, dim), nn.SiLU(), nn.Dropout(dropout),
                                              nn.Conv3d(dim, dim, (3, 1, 1), padding=(1, 0, 0))))
        nn.init.zeros_(self.conv4[-1].weight)  # diffusers zero-inits the last conv (identity block at init)
        nn.init.zeros_(self.conv4[-1].bias)

    def forward(self, x, num_frames: int):
        y = x[None, :].reshape((-1, num_frames) + x.shape[1:]).permute(0, 2, 1, 3, 4)
        identity = y
        y = self.conv4(self.conv3(self.conv2(self.conv1(y))))
        y = identity + y
        return y.permute(0, 2, 1, 3, 4).reshape((y.shape[0] * y.shape[2], -1) + y.shape[3:])


class Downsample2D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=1)

    def forward(self, x, scale: float = 1.0):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x, output_size=None, scale: float = 1.0):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class _Block3D(nn.Module):
    """Shared body of the 3-D down / up blocks: per layer resnet -> temp_conv [-> attn -> temp_attn]."""

    def _layer(self, i, x, temb, ctx, num_frames):
        x = self.resnets[i](x, temb)
        x = self.temp_convs[i](x, num_frames=num_frames)
        if self.has_cross_attention:
            x = self.attentions[i](x, encoder_hidden_states=ctx)
            x = self.temp_attentions[i](x, num_frames=num_frames)
        return x


class DownBlock3D(_Block3D):
    def __init__(self, in_ch, out_ch, temb_ch, layers, heads, cross_dim, groups, attn: bool, add_downsample: bool):
        super().__init__()
        self.has_cross_attention = attn
        self.resnets = nn.ModuleList([ResnetBlock2D(in_ch if i == 0 else out_ch, out_ch, temb_ch, groups) for i in range(layers)])
        self.temp_convs = nn.ModuleList([TemporalConvLayer(out_ch, groups) for _ in range(layers)])
        if attn:
            self.attentions = nn.ModuleList([Transformer2DModel(out_ch // heads, heads, out_ch, cross_dim, groups) for _ in range(layers)])
            self.temp_attentions = nn.ModuleList([TransformerTemporalModel(out_ch // heads, heads, out_ch, groups) for _ in range(layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(out_ch)]) if add_downsample else None

    def forward(self, x, temb, ctx, num_frames):
        outs = ()
        for i in range(len(self.resnets)):
            x = self._layer(i, x, temb, ctx, num_frames)
            outs += (x,)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs += (x,)
        return x, outs


class UpBlock3D(_Block3D):
    def __init__(self, in_ch, out_ch, prev_ch, temb_ch, layers, heads, cross_dim, groups, attn: bool, add_upsample: bool):
        super().__init__()
        self.has_cross_attention = attn
        res = []
        for i in range(layers):
            skip = in_ch if i == layers - 1 else out_ch
            rin = prev_ch if i == 0 else out_ch
            res.append(ResnetBlock2D(rin + skip, out_ch, temb_ch, groups))
        self.resnets = nn.ModuleList(res)
        self.temp_convs = nn.ModuleList([TemporalConvLayer(out_ch, groups) for _ in range(layers)])
        if attn:
            self.attentions = nn.ModuleList([Transformer2DModel(out_ch // heads, heads, out_ch, cross_dim, groups) for _ in range(layers)])
            self.temp_attentions = nn.ModuleList([TransformerTemporalModel(out_ch // heads, heads, out_ch, groups) for _ in range(layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(out_ch)]) if add_upsample else None

    def forward(self, x, skips, temb, ctx, num_frames):
        for i in range(len(self.resnets)):
            x = torch.cat([x, skips[-1]], dim=1)
            skips = skips[:-1]
            x = self._layer(i, x, temb, ctx, num_frames)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class MidBlock3D(nn.Module):
    def __init__(self, ch, temb_ch, heads, cross_dim, groups):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, temb_ch, groups), ResnetBlock2D(ch, ch, temb_ch, groups)])
        self.temp_convs = nn.ModuleList([TemporalConvLayer(ch, groups), TemporalConvLayer(ch, groups)])
        self.attentions = nn.ModuleList([Transformer2DModel(ch // heads, heads, ch, cross_dim, groups)])
        self.temp_attentions = nn.ModuleList([TransformerTemporalModel(ch // heads, heads, ch, groups)])

    def forward(self, x, temb, ctx, num_frames):
        x = self.temp_convs[0](self.resnets[0](x, temb), num_frames=num_frames)
        x = self.attentions[0](x, encoder_hidden_states=ctx)
        x = self.temp_attentions[0](x, num_frames=num_frames)
        return self.temp_convs[1](self.resnets[1](x, temb), num_frames=num_frames)


# --------------------------------------------------------------------------------------------- embeddings
def timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0): [cos | sin], fp32."""
    half = dim // 2
    freq = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    arg = t[:, None].float() * freq[None, :]
    return torch.cat([torch.cos(arg), torch.sin(arg)], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_ch, dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_ch, dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(self.act(self.linear_1(x)))


class I2VGenXLTransformerTemporalEncoder(nn.Module):
    def __init__(self, dim, heads, head_dim, ff_inner_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-5)
        self.attn1 = Attention(dim, None, heads, head_dim)
        self.ff = FeedForward(dim, activation_fn="gelu", inner_dim=ff_inner_dim)

    def forward(self, x):
        x = self.attn1(self.norm1(x), encoder_hidden_states=None) + x
        return self.ff(x) + x


# --------------------------------------------------------------------------------------------- the UNet
I2VGEN_XL_CONFIG = dict(in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
                        norm_num_groups=32, cross_attention_dim=1024, head_dim=64, transformer_in_heads=8)
#: small geometry for CPU tests: same topology (4 levels, attention at the same places, head_dim 64, 32 groups)
TINY_CONFIG = dict(in_channels=4, out_channels=4, block_out_channels=(64, 128, 128, 128), layers_per_block=2,
                   norm_num_groups=32, cross_attention_dim=64, head_dim=64, transformer_in_heads=2)


class I2VGenXLUNet(nn.Module):
    def __init__(self, in_channels=4, out_channels=4, block_out_channels: Sequence[int] = (320, 640, 1280, 1280),
                 layers_per_block=2, norm_num_groups=32, cross_attention_dim=1024, head_dim=64, transformer_in_heads=8):
        super().__init__()
        self.config = dict(in_channels=in_channels, out_channels=out_channels, block_out_channels=tuple(block_out_channels),
                           layers_per_block=layers_per_block, norm_num_groups=norm_num_groups,
                           cross_attention_dim=cross_attention_dim, head_dim=head_dim,
                           transformer_in_heads=transformer_in_heads)
        c0 = block_out_channels[0]
        g = norm_num_groups
        temb = c0 * 4
        self.conv_in = nn.Conv2d(in_channels * 2, c0, 3, padding=1)
        self.transformer_in = TransformerTemporalModel(transformer_in_heads, head_dim, c0, g)
        self.image_latents_proj_in = nn.Sequential(
            nn.Conv2d(4, in_channels * 4, 3, padding=1), nn.SiLU(),
            nn.Conv2d(in_channels * 4, in_channels * 4, 3, padding=1), nn.SiLU(),
            nn.Conv2d(in_channels * 4, in_channels, 3, padding=1))
        self.image_latents_temporal_encoder = I2VGenXLTransformerTemporalEncoder(in_channels, 2, in_channels, in_channels * 4)
        self.image_latents_context_embedding = nn.Sequential(
            nn.Conv2d(4, in_channels * 8, 3, padding=1), nn.SiLU(), nn.AdaptiveAvgPool2d((32, 32)),
            nn.Conv2d(in_channels * 8, in_channels * 16, 3, stride=2, padding=1), nn.SiLU(),
            nn.Conv2d(in_channels * 16, cross_attention_dim, 3, stride=2, padding=1))
        self.time_embedding = TimestepEmbedding(c0, temb)
        self.context_embedding = nn.Sequential(nn.Linear(cross_attention_dim, temb), nn.SiLU(),
                                               nn.Linear(temb, cross_attention_dim * in_channels))
        self.fps_embedding = nn.Sequential(nn.Linear(c0, temb), nn.SiLU(), nn.Linear(temb, temb))

        n = len(block_out_channels)
        self.down_blocks = nn.ModuleList()
        out_ch = c0
        for i, ch in enumerate(block_out_channels):
            in_ch, out_ch = out_ch, ch
            self.down_blocks.append(DownBlock3D(in_ch, out_ch, temb, layers_per_block, head_dim, cross_attention_dim, g,
                                                attn=i < n - 1, add_downsample=i < n - 1))
        self.mid_block = MidBlock3D(block_out_channels[-1], temb, head_dim, cross_attention_dim, g)
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(block_out_channels))
        out_ch = rev[0]
        for i in range(n):
            prev, out_ch = out_ch, rev[i]
            in_ch = rev[min(i + 1, n - 1)]
            self.up_blocks.append(UpBlock3D(in_ch, out_ch, prev, temb, layers_per_block + 1, head_dim, cross_attention_dim,
                                            g, attn=i > 0, add_upsample=i < n - 1))
        self.conv_norm_out = nn.GroupNorm(g, c0, eps=1e-5)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(c0, out_channels, 3, padding=1)

    @property
    def dtype(self):
        return self.conv_in.weight.dtype

    def forward(self, sample, timestep, fps, image_latents, image_embeddings=None, encoder_hidden_states=None,
                cross_attention_kwargs=None, return_dict: bool = False):
        b, c, f, h, w = sample.shape
        dt = self.dtype
        t = timestep if torch.is_tensor(timestep) else torch.tensor([timestep], device=sample.device)
        t = t.reshape(-1).to(sample.device).expand(b)
        c0 = self.config["block_out_channels"][0]
        emb = self.time_embedding(timestep_embedding(t, c0).to(dt))
        emb = emb + self.fps_embedding(timestep_embedding(fps.reshape(-1).expand(b), c0).to(dt))
        emb = emb.repeat_interleave(f, dim=0)

        first = image_latents[:, :, 0]  # [b,4,h,w]
        lat_ctx = self.image_latents_context_embedding(first)
        lat_ctx = lat_ctx.permute(0, 2, 3, 1).reshape(b, -1, lat_ctx.shape[1])
        img_ctx = self.context_embedding(image_embeddings).view(-1, self.config["in_channels"], self.config["cross_attention_dim"])
        ctx = torch.cat([encoder_hidden_states, lat_ctx, img_ctx], dim=1).repeat_interleave(f, dim=0)

        il = image_latents.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
        il = self.image_latents_proj_in(il)
        il = il[None, :].reshape(b, f, c, h, w).permute(0, 3, 4, 1, 2).reshape(b * h * w, f, c)
        il = self.image_latents_temporal_encoder(il)
        il = il.reshape(b, h, w, f, c).permute(0, 4, 3, 1, 2)

        x = torch.cat([sample, il], dim=1)
        x = x.permute(0, 2, 1, 3, 4).reshape(b * f, 2 * c, h, w)
        x = self.conv_in(x)
        x = self.transformer_in(x, num_frames=f)

        skips = (x,)
        for blk in self.down_blocks:
            x, outs = blk(x, emb, ctx, f)
            skips += outs
        x = self.mid_block(x, emb, ctx, f)
        for blk in self.up_blocks:
            k = len(blk.resnets)
            x = blk(x, skips[-k:], emb, ctx, f)
            skips = skips[:-k]
        x = self.conv_out(self.conv_act(self.conv_norm_out(x)))
        x = x[None, :].reshape((-1, f) + x.shape[1:]).permute(0, 2, 1, 3, 4)
        return (x,)


def seeded_unet(config: dict, seed: int = 8888, dtype=torch.float32, device="cpu") -> I2VGenXLUNet:
    """Deterministic random-init UNet. Zero-initialised TemporalConvLayer.conv4 is re-randomised (std 0.02) so the
    temporal convolutions are exercised."""
    gen_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    try:
        net = I2VGenXLUNet(**config)
        g = torch.Generator().manual_seed(seed + 1)
        for m in net.modules():
            if isinstance(m, TemporalConvLayer):
                conv = m.conv4[-1]
                with torch.no_grad():
                    conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * 0.02)
                    conv.bias.copy_(torch.randn(conv.bias.shape, generator=g) * 0.02)
    finally:
        torch.random.set_rng_state(gen_state)
    return net.to(device=device, dtype=dtype).eval()

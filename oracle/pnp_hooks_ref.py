"""ORACLE (test infrastructure): compact restatement of the reference's three PnP hooks + register_time.

Follows /root/reference/i2vgen-xl/pnp_utils.py: register_time :19-28, conv_forward :41-126 (injection :109-115),
ModifiedSpaAttnProcessor :141-228 (injection :189-196), ModifiedTmpAttnProcessor :247-334 (injection :295-302),
site tables :130, :235-242, :340-346.  oracle/make_golden.py proves (in the build container, where /root/reference
exists) that these give bit-identical outputs to the UNMODIFIED reference file executed through
oracle/diffusers_shim.py; the GPU box has no /root/reference, so tests there use this restatement.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

SPATIAL_SITES = {1: [1, 2], 2: [0, 1, 2], 3: [0, 1, 2]}   # pnp_utils.py:235
TIMED_SITES = {1: [0, 1, 2], 2: [0, 1, 2], 3: [0, 1, 2]}   # pnp_utils.py:22


def _fires(obj) -> bool:
    sched = obj.injection_schedule
    if sched is None:
        return False
    return bool(obj.t in sched) or obj.t == 1000


def register_time(model, t):
    setattr(model.unet.up_blocks[1].resnets[1], "t", t)
    for res, blocks in TIMED_SITES.items():
        for blk in blocks:
            up = model.unet.up_blocks[res]
            setattr(up.attentions[blk].transformer_blocks[0].attn1.processor, "t", t)
            setattr(up.temp_attentions[blk].transformer_blocks[0].attn1.processor, "t", t)


def register_conv_injection(model, injection_schedule):
    mod = model.unet.up_blocks[1].resnets[1]

    def forward(input_tensor, temb, scale: float = 1.0):
        h = mod.conv1(mod.nonlinearity(mod.norm1(input_tensor)))
        h = h + mod.time_emb_proj(mod.nonlinearity(temb))[:, :, None, None]
        h = mod.conv2(mod.dropout(mod.nonlinearity(mod.norm2(h))))
        if _fires(mod):
            n = h.shape[0] // 3
            h[n:2 * n] = h[:n]
            h[2 * n:] = h[:n]
        x = mod.conv_shortcut(input_tensor) if mod.conv_shortcut is not None else input_tensor
        return (x + h) / mod.output_scale_factor

    mod.forward = forward
    mod.injection_schedule = injection_schedule


class PnPAttnProcessor:
    """Self-attention with the source chunk's q,k copied over the uncond / cond chunks on scheduled timesteps."""

    def __init__(self, injection_schedule):
        self.injection_schedule = injection_schedule
        self.t = None

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, scale=1.0):
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        b = hidden_states.shape[0]
        q, k, v = attn.to_q(hidden_states), attn.to_k(ctx), attn.to_v(ctx)
        if _fires(self):
            n = b // 3
            q[n:2 * n] = q[:n]
            k[n:2 * n] = k[:n]
            q[2 * n:] = q[:n]
            k[2 * n:] = k[:n]
        hd = k.shape[-1] // attn.heads
        q = q.view(b, -1, attn.heads, hd).transpose(1, 2)
        k = k.view(b, -1, attn.heads, hd).transpose(1, 2)
        v = v.view(b, -1, attn.heads, hd).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(b, -1, attn.heads * hd).to(q.dtype)
        o = attn.to_out[1](attn.to_out[0](o))
        return o / attn.rescale_output_factor


def register_spatial_attention_pnp(model, injection_schedule):
    for res, blocks in SPATIAL_SITES.items():
        for blk in blocks:
            model.unet.up_blocks[res].attentions[blk].transformer_blocks[0].attn1.processor = PnPAttnProcessor(injection_schedule)


def register_temp_attention_pnp(model, injection_schedule):
    for res, blocks in SPATIAL_SITES.items():
        for blk in blocks:
            model.unet.up_blocks[res].temp_attentions[blk].transformer_blocks[0].attn1.processor = PnPAttnProcessor(injection_schedule)


def init_pnp(model, scheduler, n_steps: int, pnp_f_t: float, pnp_spatial_attn_t: float, pnp_temp_attn_t: float):
    """run_group_pnp_edit.py:35-48 — schedule = first int(n_steps*frac) entries of the FULL timestep list."""
    def sched(frac):
        k = int(n_steps * frac)
        return scheduler.timesteps[:k] if k >= 0 else []
    register_conv_injection(model, sched(pnp_f_t))
    register_spatial_attention_pnp(model, sched(pnp_spatial_attn_t))
    register_temp_attention_pnp(model, sched(pnp_temp_attn_t))

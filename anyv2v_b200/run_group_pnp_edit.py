"""Group runner, phase 2 — drop-in for the reference's ``i2vgen-xl/run_group_pnp_edit.py``.

Same CLI (``--template_config``, ``--configs_json``; reference :186-192), same YAML / JSON keys and override
semantics (template merged with each JSON entry, ``"active": false`` skips; :74-81), same ``init_pnp`` arithmetic
(:35-48), same ``ddim_latents_{t}.pt`` inputs and the same output directory naming (:154-168).
What changes: clips are sharded one-per-GPU when launched under torchrun (the reference loops over them on one
device, :74), the UNet is this package's B200 model, and VAE / CLIP / video export — which bracket the loop and are
outside the metric — are pluggable: without encoders the runner uses the seeded synthetic conditioning of
SURVEY 8d (``synthetic: true`` in the config, the default when no encoders are attached) and writes the edited
LATENTS (``edited_latents.pt``) instead of mp4/gif/png.
"""
from __future__ import annotations

import argparse
import json
import logging
import os
from pathlib import Path

import torch

from .config import OmegaConf
from .latent_store import LatentStore, load_ddim_latents_at_t
from .pipeline import I2VGenXLPipeline, frame_position_latents
from .pnp_utils import register_conv_injection, register_spatial_attention_pnp, register_temp_attention_pnp
from .schedulers import DDIMScheduler

logger = logging.getLogger(__name__)


def seed_everything(seed: int) -> None:
    """i2vgen-xl/utils.py:17-22."""
    import random

    import numpy as np
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    random.seed(seed)
    np.random.seed(seed)


def pnp_schedules(scheduler, config):
    """reference :36-45 — k = int(n_steps * frac); schedule = first k entries of the FULL descending timestep list;
    a negative fraction gives an empty schedule."""
    out = []
    for frac in (config.pnp_f_t, config.pnp_spatial_attn_t, config.pnp_temp_attn_t):
        k = int(config.n_steps * frac)
        out.append(scheduler.timesteps[:k] if k >= 0 else [])
    return out


def init_pnp(pipe, scheduler, config):
    conv_t, spa_t, tmp_t = pnp_schedules(scheduler, config)
    register_conv_injection(pipe, conv_t)
    register_spatial_attention_pnp(pipe, spa_t)
    register_temp_attention_pnp(pipe, tmp_t)
    logger.debug("conv_injection_timesteps: %s", conv_t)
    logger.debug("spatial_attn_qk_injection_timesteps: %s", spa_t)
    logger.debug("temp_attn_qk_injection_timesteps: %s", tmp_t)


def synthetic_conditioning(n_frames: int, h: int, w: int, cross_dim: int, seed: int, device, dtype=torch.float16):
    """Seeded stand-ins for the CLIP / VAE outputs; same generator order as oracle.loops_ref."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g).to(device=device, dtype=dtype)
    c = {}
    c["video_latents"] = rn(1, 4, n_frames, h, w)
    c["src_image_latents"] = frame_position_latents(rn(1, 4, h, w), n_frames)
    c["edit_image_latents"] = frame_position_latents(rn(1, 4, h, w), n_frames)
    c["inv_prompt"], c["neg_prompt"], c["edit_prompt"] = rn(1, 77, cross_dim), rn(1, 77, cross_dim), rn(1, 77, cross_dim)
    c["src_image_emb"], c["edit_image_emb"] = rn(1, 1, cross_dim), rn(1, 1, cross_dim)
    return c


def config_suffix(config) -> str:
    """reference :154-167."""
    return ("ddim_init_latents_t_idx_" + str(config.ddim_init_latents_t_idx) + "_nsteps_" + str(config.n_steps) + "_cfg_"
            + str(config.cfg) + "_pnpf" + str(config.pnp_f_t) + "_pnps" + str(config.pnp_spatial_attn_t) + "_pnpt"
            + str(config.pnp_temp_attn_t))


def edit_one(pipe, ddim_scheduler, config, device, rank_seed_offset: int = 0):
    config.video_path = os.path.join(config.video_dir, config.video_name + ".mp4")
    config.video_frames_path = os.path.join(config.video_dir, config.video_name)
    config.edited_first_frame_path = os.path.join(config.data_dir, config.edited_first_frame_path)
    for k, v in config.items():
        if "ReplaceMe" in str(v):
            logger.error("Field %s contains 'ReplaceMe'", k)
    h, w = config.image_size[1] // 8, config.image_size[0] // 8
    cross_dim = pipe.unet.config["cross_attention_dim"]
    cond = synthetic_conditioning(config.n_frames, h, w, cross_dim, config.seed + rank_seed_offset, device)

    ddim_scheduler.set_timesteps(config.n_steps)
    t_idx = config.ddim_init_latents_t_idx
    t0 = int(ddim_scheduler.timesteps[t_idx])
    store = LatentStore(config.ddim_latents_path, write_files=False)
    ddim_latents_at_t = load_ddim_latents_at_t(t0, config.ddim_latents_path, map_location=device)
    random_latents = torch.randn_like(ddim_latents_at_t)
    logger.info("Blending random_ratio (1 means random latent): %s", config.random_ratio)
    mixed = random_latents * config.random_ratio + ddim_latents_at_t * (1 - config.random_ratio)

    init_pnp(pipe, ddim_scheduler, config)
    pipe.register_modules(scheduler=ddim_scheduler)
    out = pipe.sample_with_pnp(
        latents=mixed, prompt_embeds=cond["edit_prompt"], negative_prompt_embeds=cond["neg_prompt"],
        ddim_inv_prompt_embeds=cond["inv_prompt"], image_embeddings=cond["edit_image_emb"],
        image_latents=cond["edit_image_latents"], ddim_inv_image_embeddings=cond["src_image_emb"],
        ddim_inv_image_latents=cond["src_image_latents"], num_frames=config.n_frames,
        num_inference_steps=config.n_steps, guidance_scale=config.cfg, target_fps=config.target_fps,
        ddim_init_latents_t_idx=t_idx, ddim_inv_latents_path=config.ddim_latents_path, latent_store=store,
        output_type="latent").frames
    output_dir = os.path.join(config.output_dir, config_suffix(config))
    os.makedirs(output_dir, exist_ok=True)
    torch.save(out.cpu(), os.path.join(output_dir, "edited_latents.pt"))
    logger.info("Saved edited latents to: %s", output_dir)
    return out


def build_pipeline(device, unet_config=None, seed: int = 8888, broadcast: bool = True):
    """Random-init I2VGen-XL UNet (no checkpoint can be downloaded here).  Under torchrun rank 0 initialises the
    weights and every other rank receives them by ONE NCCL broadcast of the flat fp16 buffer (anyv2v_b200.distributed)."""
    from . import distributed
    from .unet_i2vgen_xl import I2VGEN_XL_CONFIG, I2VGenXLUNet
    cfg = dict(unet_config or I2VGEN_XL_CONFIG)
    unet = distributed.build_unet_replicated(I2VGenXLUNet, cfg, seed, device, broadcast=broadcast)
    return I2VGenXLPipeline(unet, DDIMScheduler())


def main(template_config, configs_list, device, unet_config=None):
    from . import distributed
    rank, world = distributed.rank_world()
    pipe = build_pipeline(device, unet_config, seed=template_config.seed)
    ddim_scheduler = DDIMScheduler.from_pretrained("ali-vilab/i2vgen-xl", subfolder="scheduler")
    active = [e for e in configs_list if e.get("active", True)]
    for e in configs_list:
        if not e.get("active", True):
            logger.info("Skipping config_entry: %s", e)
    results = []
    for i, entry in enumerate(active):
        if i % world != rank:  # clips shard one-per-GPU; no data-path collective
            continue
        logger.info("Processing config_entry: %s", entry)
        config = OmegaConf.merge(template_config, OmegaConf.create(entry))
        results.append(edit_one(pipe, ddim_scheduler, config, device, rank_seed_offset=i))
    return results


def cli(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--template_config", type=str, default="./configs/group_pnp_edit/template.yaml")
    parser.add_argument("--configs_json", type=str, default="./configs/group_config.json")
    args = parser.parse_args(argv)
    template_config = OmegaConf.load(args.template_config)
    logging.basicConfig(level=logging.DEBUG if template_config.debug else logging.INFO,
                        format="%(asctime)s - %(levelname)s - [%(funcName)s] - %(message)s")
    logger.info("template_config: %s", OmegaConf.to_yaml(template_config))
    assert Path(args.configs_json).exists()
    with open(args.configs_json, "r") as fh:
        configs_list = json.load(fh)
    logger.info("Loaded %d configs from %s", len(configs_list), args.configs_json)
    from . import distributed
    device = distributed.pick_device(template_config.device)
    torch.set_grad_enabled(False)
    seed_everything(template_config.seed)
    return main(template_config, configs_list, device)


if __name__ == "__main__":
    cli()

"""Group runner, phase 1 — drop-in for the reference's ``i2vgen-xl/run_group_ddim_inversion.py``.

Same CLI (:195-198), same template keys (configs/group_ddim_inversion/template.yaml: ``inverse_config.{cfg,
target_fps, prompt, negative_prompt, n_steps, output_dir, ...}``, ``recon_config.*``), same skip rule (existing
``output_dir`` and not ``force_recompute_latents``; :118-120) and the same per-timestep ``ddim_latents_{t}.pt`` files
(pipeline :1424-1428).  ``ddim_inversion(config, first_frame, frame_list, pipe, inverse_scheduler, g)`` keeps the
reference signature (:29); ``first_frame`` / ``frame_list`` may be PIL images when ``pipe.encoders`` is attached, or
pre-encoded tensors / None for the synthetic conditioning of SURVEY 8d.
"""
from __future__ import annotations

import argparse
import json
import logging
import os
from pathlib import Path

import torch

from .config import OmegaConf
from .pipeline import I2VGenXLPipeline
from .run_group_pnp_edit import build_pipeline, seed_everything, synthetic_conditioning
from .schedulers import DDIMInverseScheduler, DDIMScheduler

logger = logging.getLogger(__name__)


def ddim_inversion(config, first_frame, frame_list, pipe: I2VGenXLPipeline, inverse_scheduler, g, cond=None):
    """reference :29-55.  Returns the stacked inverted latents [b, steps, c, f, h, w]."""
    pipe.scheduler = inverse_scheduler
    if cond is None:
        raise ValueError("pre-encoded conditioning is required (VAE / CLIP are outside the hot path)")
    return pipe.invert(
        latents=cond["video_latents"], prompt_embeds=cond["inv_prompt"], image_latents=cond["src_image_latents"],
        image_embeddings=cond["src_image_emb"], num_frames=config.n_frames, num_inference_steps=config.n_steps,
        guidance_scale=config.cfg, target_fps=config.target_fps, output_dir=config.output_dir, return_dict=False)


def ddim_sampling(config, first_frame, ddim_latents_path, pipe, ddim_scheduler, ddim_init_latents_t_idx, g, cond=None):
    """reference :58-77 (DDIM reconstruction, the authors' sanity check): plain CFG sampling from x_t without hooks."""
    from .latent_store import load_ddim_latents_at_t
    ddim_scheduler.set_timesteps(config.n_steps)
    ts = ddim_scheduler.timesteps.tolist()[ddim_init_latents_t_idx:]
    latents = load_ddim_latents_at_t(ts[0], ddim_latents_path, map_location=pipe.device)
    dev = pipe.device
    prompts = torch.cat([cond["neg_prompt"], cond["inv_prompt"]])
    img_emb = torch.cat([torch.zeros_like(cond["src_image_emb"]), cond["src_image_emb"]])
    img_lat = torch.cat([cond["src_image_latents"]] * 2)
    c2 = pipe.unet.precompute_conditioning(torch.tensor([config.target_fps] * 2, device=dev), img_lat, img_emb, prompts)
    for t in ts:
        v = pipe.unet(torch.cat([latents, latents]), torch.tensor([t], device=dev), cond=c2)[0]
        latents = ddim_scheduler.step(v[0:1], t, latents, model_output_cond=v[1:2], guidance_scale=config.cfg).prev_sample
    return latents


def main(template_config, configs_list, device, unet_config=None):
    from . import distributed
    rank, world = distributed.rank_world()
    pipe = build_pipeline(device, unet_config, seed=template_config.seed)
    g = torch.Generator(device=device).manual_seed(template_config.seed)
    inverse_scheduler = DDIMInverseScheduler.from_pretrained("ali-vilab/i2vgen-xl", subfolder="scheduler")
    ddim_scheduler = DDIMScheduler.from_pretrained("ali-vilab/i2vgen-xl", subfolder="scheduler")
    assert len(configs_list) > 0
    active = [e for e in configs_list if e.get("active", True)]
    out = []
    for i, entry in enumerate(active):
        if i % world != rank:
            continue
        logger.info("Processing config_entry: %s", entry)
        config = OmegaConf.merge(template_config, OmegaConf.create(entry))
        config.video_path = os.path.join(config.video_dir, config.video_name + ".mp4")
        config.video_frames_path = os.path.join(config.video_dir, config.video_name)
        if os.path.exists(config.output_dir) and not config.get("force_recompute_latents", False):
            logger.info("= Inverted latents already exist at %s. Skip.", config.output_dir)
            continue
        h, w = config.image_size[1] // 8, config.image_size[0] // 8
        cond = synthetic_conditioning(config.n_frames, h, w, pipe.unet.config["cross_attention_dim"], config.seed + i, device)
        inv = ddim_inversion(config.inverse_config, None, None, pipe, inverse_scheduler, g, cond=cond)
        out.append(inv)
        rc = config.recon_config
        if rc.enable_recon:
            rec = ddim_sampling(rc, None, rc.ddim_latents_path, pipe, ddim_scheduler, rc.ddim_init_latents_t_idx, g, cond=cond)
            os.makedirs(os.path.join(config.output_dir, "ddim_reconstruction"), exist_ok=True)
            torch.save(rec.cpu(), os.path.join(config.output_dir, "ddim_reconstruction", "latents.pt"))
    return out


def cli(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--template_config", type=str, default="./configs/group_ddim_inversion/template.yaml")
    parser.add_argument("--configs_json", type=str, default="./configs/group_ddim_inversion/group_config.json")
    args = parser.parse_args(argv)
    template_config = OmegaConf.load(args.template_config)
    logging.basicConfig(level=logging.DEBUG if template_config.debug else logging.INFO,
                        format="%(asctime)s - %(levelname)s - [%(funcName)s] - %(message)s")
    assert Path(args.configs_json).exists()
    with open(args.configs_json, "r") as fh:
        configs_list = json.load(fh)
    from . import distributed
    device = distributed.pick_device(template_config.device)
    torch.set_grad_enabled(False)
    seed_everything(template_config.seed)
    return main(template_config, configs_list, device)


if __name__ == "__main__":
    cli()

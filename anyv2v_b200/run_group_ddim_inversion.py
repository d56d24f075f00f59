"""Group runner, phase 1 — drop-in for the reference's ``i2vgen-xl/run_group_ddim_inversion.py``.

Same CLI (:195-198), same template keys (configs/group_ddim_inversion/template.yaml: ``inverse_config.{cfg,
target_fps, prompt, negative_prompt, n_steps, output_dir, ...}``, ``recon_config.*``), same skip rule (existing
``output_dir`` and not ``force_recompute_latents``; :118-120) and the same per-timestep ``ddim_latents_{t}.pt`` files
(pipeline :1424-1428).  ``ddim_inversion(config, first_frame, frame_list, pipe, inverse_scheduler, g)`` keeps the
reference signature (:29) and return value (``[steps, c, f, h, w]``, :54): with PIL ``first_frame`` / ``frame_list`` it runs
the real path (VAE-encode the frames, CLIP-encode prompt and first frame — the pipeline must carry ``encoders`` and
``vae``); ``cond=`` (pre-encoded tensors) is the ``synthetic: true`` opt-in of SURVEY 8d.  The source frames are read like the
reference does (:125-139; png frames, mp4 fallback), incl. ``inverse_static_video`` / ``null_image_inversion`` (:143-151).
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
from .run_group_pnp_edit import _model_dir, build_pipeline, load_source_frames, seed_everything, synthetic_conditioning
from .schedulers import DDIMInverseScheduler, DDIMScheduler

logger = logging.getLogger(__name__)


def ddim_inversion(config, first_frame, frame_list, pipe: I2VGenXLPipeline, inverse_scheduler, g, cond=None):
    """reference :29-55.  Returns the inverted latents of the clip, [steps, c, f, h, w] (descending t), like :54."""
    pipe.scheduler = inverse_scheduler
    if cond is not None:  # synthetic opt-in: pre-encoded conditioning
        return pipe.invert(
            latents=cond["video_latents"], prompt_embeds=cond["inv_prompt"], negative_prompt_embeds=cond.get("neg_prompt"),
            image_latents=cond["src_image_latents"], image_embeddings=cond["src_image_emb"], num_frames=config.n_frames,
            num_inference_steps=config.n_steps, guidance_scale=config.cfg, target_fps=config.target_fps,
            output_dir=config.output_dir, return_dict=False)[0]
    if first_frame is None or frame_list is None:
        raise ValueError("ddim_inversion needs the source frames (PIL) or `cond=` (pre-encoded, `synthetic: true`)")
    width, height = int(config.image_size[0]), int(config.image_size[1])
    video_latents_at_0 = pipe.encode_vae_video(frame_list, device=pipe._execution_device, height=height, width=width, generator=g)
    return pipe.invert(
        prompt=config.prompt, image=first_frame, height=height, width=width, num_frames=config.n_frames,
        num_inference_steps=config.n_steps, guidance_scale=config.cfg, negative_prompt=config.negative_prompt,
        target_fps=config.target_fps, latents=video_latents_at_0, generator=g, return_dict=False,
        output_dir=config.output_dir)[0]


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


def main(template_config, configs_list, device, unet_config=None, pipeline_kwargs=None):
    from . import distributed
    rank, world = distributed.rank_world()
    active = [e for e in configs_list if e.get("active", True)]
    need_real = any(not OmegaConf.merge(template_config, OmegaConf.create(e)).get("synthetic", False) for e in active)
    pipe = build_pipeline(device, unet_config, seed=template_config.seed, with_encoders=need_real,
                          model_dir=_model_dir(template_config), **(pipeline_kwargs or {}))
    g = torch.Generator(device=device).manual_seed(template_config.seed)
    inverse_scheduler = DDIMInverseScheduler.from_pretrained("ali-vilab/i2vgen-xl", subfolder="scheduler")
    ddim_scheduler = DDIMScheduler.from_pretrained("ali-vilab/i2vgen-xl", subfolder="scheduler")
    assert len(configs_list) > 0
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
        if config.get("synthetic", False):
            h, w = config.image_size[1] // 8, config.image_size[0] // 8
            cond = synthetic_conditioning(config.n_frames, h, w, pipe.unet.config["cross_attention_dim"], config.seed + i, device)
            inv = ddim_inversion(config.inverse_config, None, None, pipe, inverse_scheduler, g, cond=cond)
        else:
            from PIL import Image

            from . import image_io
            cond = None
            frame_list = load_source_frames(config)
            if not os.path.exists(os.path.join(config.video_frames_path, config.video_name + ".gif")):
                try:  # reference :137-141 saves the source frames as a gif next to them
                    image_io.export_to_gif(frame_list, os.path.join(config.video_frames_path, config.video_name + ".gif"))
                except OSError as e:
                    logger.info("source gif not written (%s)", e)
            first_frame = frame_list[0]
            if config.inverse_config.get("inverse_static_video", False):
                logger.info("### Inverse a static video!")
                frame_list = [frame_list[0]] * config.n_frames
            if config.inverse_config.get("null_image_inversion", False):
                logger.info("### Inverse a null image!")
                first_frame = Image.new("RGB", (config.image_size[0], config.image_size[1]), (0, 0, 0))
            inv = ddim_inversion(config.inverse_config, first_frame, frame_list, pipe, inverse_scheduler, g)
        out.append(inv)
        rc = config.recon_config
        if rc.enable_recon:
            if cond is None:  # real inputs: encode what the reconstruction's plain CFG sampling needs (reference :58-77)
                emb, lat = pipe.encode_first_frame(first_frame, int(config.image_size[1]), int(config.image_size[0]), config.n_frames)
                cond = {"neg_prompt": pipe.encode_prompt(rc.negative_prompt), "inv_prompt": pipe.encode_prompt(rc.prompt),
                        "src_image_emb": emb, "src_image_latents": lat}
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

"""GPU tests of the runner / config surface: the two group runners end to end (tiny UNet), the reference's file
hand-off between the phases, and CUDA-graph replay vs eager launches."""
import json
import os
from types import SimpleNamespace

import pytest
import torch
import yaml

pytestmark = pytest.mark.gpu
dev = "cuda"

INV_TEMPLATE = {
    "seed": 8888, "device": "cuda:0", "debug": False, "data_dir": "DATA", "model_name": "i2vgen-xl", "exp_name": "${video_name}",
    "output_dir": "${data_dir}/inversions/${model_name}/${exp_name}", "image_size": [128, 128], "video_dir": "${data_dir}/demo",
    "video_name": "ReplaceMe", "video_path": "ReplaceMe", "video_frames_path": "ReplaceMe", "n_frames": 4,
    "inverse_config": {"image_size": "${image_size}", "n_frames": "${n_frames}", "cfg": 1.0, "target_fps": 8, "prompt": "",
                       "negative_prompt": "", "n_steps": 5, "output_dir": "${output_dir}/ddim_latents",
                       "inverse_static_video": False, "null_image_inversion": False},
    "recon_config": {"enable_recon": True, "image_size": "${image_size}", "n_frames": "${n_frames}", "cfg": 9.0, "target_fps": 8,
                     "prompt": "", "negative_prompt": "x", "n_steps": 5, "ddim_init_latents_t_idx": 1,
                     "ddim_latents_path": "${inverse_config.output_dir}"},
}
EDIT_TEMPLATE = {
    "seed": 8888, "device": "cuda:0", "debug": False, "data_dir": "DATA", "model_name": "i2vgen-xl", "task_name": "Prompt-Based-Editing",
    "edited_video_name": "ReplaceMe", "output_dir": "${data_dir}/Results/${task_name}/${model_name}/${video_name}/${edited_video_name}/",
    "image_size": [128, 128], "video_dir": "${data_dir}/demo", "video_name": "ReplaceMe", "video_path": "ReplaceMe",
    "video_frames_path": "ReplaceMe", "edited_first_frame_path": "ReplaceMe",
    "ddim_latents_path": "${data_dir}/inversions/${model_name}/${video_name}/ddim_latents", "n_frames": 4, "cfg": 9.0, "target_fps": 8,
    "editing_prompt": "ReplaceMe", "editing_negative_prompt": "bad", "n_steps": 5, "ddim_init_latents_t_idx": 1, "ddim_inv_prompt": "",
    "random_ratio": 0.0, "pnp_f_t": 0.2, "pnp_spatial_attn_t": 0.2, "pnp_temp_attn_t": 0.5,
}


def test_group_runners_end_to_end(tmp_path):
    """run_group_ddim_inversion -> ddim_latents_{t}.pt files -> run_group_pnp_edit, same template/JSON API as the reference."""
    from anyv2v_b200 import run_group_ddim_inversion as inv, run_group_pnp_edit as edit
    from anyv2v_b200.config import OmegaConf
    from oracle.unet_ref import TINY_CONFIG
    data = str(tmp_path)
    for name, tpl in (("inv.yaml", INV_TEMPLATE), ("edit.yaml", EDIT_TEMPLATE)):
        t = dict(tpl, data_dir=data)
        (tmp_path / name).write_text(yaml.safe_dump(t))
    entries = [{"active": True, "video_name": "clipA", "edited_first_frame_path": "demo/clipA/edited.png", "editing_prompt": "a robot",
                "edited_video_name": "robot", "ddim_init_latents_t_idx": 0, "pnp_f_t": 1.0, "pnp_spatial_attn_t": 1.0, "pnp_temp_attn_t": 1.0},
               {"active": False, "video_name": "skipped", "edited_first_frame_path": "x", "editing_prompt": "x", "edited_video_name": "x"}]
    torch.set_grad_enabled(False)
    device = torch.device("cuda", 0)
    out = inv.main(OmegaConf.load(str(tmp_path / "inv.yaml")), entries, device, unet_config=TINY_CONFIG)
    assert len(out) == 1 and out[0].shape == (1, 5, 4, 4, 16, 16)
    lat_dir = os.path.join(data, "inversions", "i2vgen-xl", "clipA", "ddim_latents")
    assert sorted(os.listdir(lat_dir)) == sorted(f"ddim_latents_{t}.pt" for t in (1, 201, 401, 601, 801))
    assert os.path.exists(os.path.join(data, "inversions", "i2vgen-xl", "clipA", "ddim_reconstruction", "latents.pt"))
    # second call: the reference's skip rule (existing output_dir)
    assert inv.main(OmegaConf.load(str(tmp_path / "inv.yaml")), entries, device, unet_config=TINY_CONFIG) == []
    res = edit.main(OmegaConf.load(str(tmp_path / "edit.yaml")), entries, device, unet_config=TINY_CONFIG)
    assert len(res) == 1 and res[0].shape == (1, 4, 4, 16, 16) and torch.isfinite(res[0]).all()
    suffix = "ddim_init_latents_t_idx_0_nsteps_5_cfg_9.0_pnpf1.0_pnps1.0_pnpt1.0"
    saved = os.path.join(data, "Results", "Prompt-Based-Editing", "i2vgen-xl", "clipA", "robot", suffix, "edited_latents.pt")
    assert os.path.exists(saved) and torch.equal(torch.load(saved), res[0].cpu())


def test_cuda_graph_replay_equals_eager():
    """A captured loop iteration replayed with new device-resident scalars == the same kernels launched eagerly."""
    from anyv2v_b200.pipeline import I2VGenXLPipeline
    from anyv2v_b200.run_group_pnp_edit import build_pipeline, init_pnp, synthetic_conditioning
    from anyv2v_b200.schedulers import DDIMInverseScheduler, DDIMScheduler
    from oracle.unet_ref import TINY_CONFIG
    torch.set_grad_enabled(False)
    device = torch.device("cuda", 0)
    pipe = build_pipeline(device, TINY_CONFIG, seed=1, broadcast=False)
    c = synthetic_conditioning(4, 16, 16, 64, 3, device)
    results = {}
    for graphs in (False, True):
        I2VGenXLPipeline.use_cuda_graphs = graphs
        pipe.scheduler = DDIMInverseScheduler()
        inv = pipe.invert(latents=c["video_latents"], prompt_embeds=c["inv_prompt"], image_latents=c["src_image_latents"],
                          image_embeddings=c["src_image_emb"], target_fps=8, num_inference_steps=6, guidance_scale=1.0, write_files=False)
        store = pipe.latent_store
        es = DDIMScheduler()
        es.set_timesteps(6)
        pipe.scheduler = es
        init_pnp(pipe, es, SimpleNamespace(n_steps=6, pnp_f_t=0.5, pnp_spatial_attn_t=0.5, pnp_temp_attn_t=0.34))
        out = pipe.sample_with_pnp(latents=store.get(es.timesteps.tolist()[0]).clone(), prompt_embeds=c["edit_prompt"],
                                   negative_prompt_embeds=c["neg_prompt"], ddim_inv_prompt_embeds=c["inv_prompt"],
                                   image_embeddings=c["edit_image_emb"], image_latents=c["edit_image_latents"],
                                   ddim_inv_image_embeddings=c["src_image_emb"], ddim_inv_image_latents=c["src_image_latents"],
                                   target_fps=8, num_inference_steps=6, guidance_scale=9.0, ddim_init_latents_t_idx=0, latent_store=store).frames
        results[graphs] = (inv.clone(), out.clone())
    I2VGenXLPipeline.use_cuda_graphs = True
    assert torch.equal(results[False][0], results[True][0]), "inversion: graph replay differs from eager launches"
    assert torch.equal(results[False][1], results[True][1]), "edit: graph replay differs from eager launches"

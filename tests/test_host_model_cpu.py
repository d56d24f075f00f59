"""Host logic of the product model on CPU: the channels-last UNet wiring, the PnP de-duplication (shared probabilities,
source-only conv branch, dead-source steps), the hooks and both loops run on top of tests/kernel_contracts.py (the CPU
restatement of each kernel's contract) and are compared with the oracle.  No product code path is involved in the
emulation: the ``emulated_ops`` fixture patches anyv2v_b200.ops for one test only."""
from types import SimpleNamespace

import pytest
import torch

from parity_utils import err_stats

F_, H_, W_ = 4, 16, 16


def _models():
    from anyv2v_b200.unet_i2vgen_xl import I2VGenXLUNet
    from oracle import unet_ref
    ref32 = unet_ref.seeded_unet(unet_ref.TINY_CONFIG, seed=8888, dtype=torch.float32, device="cpu")
    ours = I2VGenXLUNet(**unet_ref.TINY_CONFIG)
    ours.load_state_dict(ref32.state_dict())
    return ref32, ours.to(dtype=torch.float16).eval()


def _inputs(dtype):
    from oracle import loops_ref
    ns = loops_ref.synthetic_inputs(F_, H_, W_, cross_dim=64, seed=8888, dtype=dtype, device="cpu")
    prompts, img_lat, img_emb, fps = loops_ref.edit_conditioning(ns)
    g = torch.Generator().manual_seed(8895)
    x3 = torch.randn(3, 4, F_, H_, W_, generator=g).to(dtype=dtype)
    return ns, x3, prompts, img_lat, img_emb, fps


def _close(got, ref32, what, rms=4e-3, mx=1.5e-2):
    assert torch.isfinite(got).all(), what
    e = err_stats(got, ref32)
    assert e["rms_rel"] <= rms and e["rel_to_max"] <= mx, (what, e)


def test_ddim_contract_is_bit_exact_against_the_oracle(emulated_ops):
    from oracle import schedulers_ref
    torch.manual_seed(0)
    x, vn, ve = (torch.randn(1001).half() for _ in range(3))
    for cls, inverse in ((schedulers_ref.DDIMScheduler, False), (schedulers_ref.DDIMInverseScheduler, True)):
        s = cls()
        s.set_timesteps(50)
        for t in (981, 501, 1):
            ca, cb, cc, cd = s.coefficients(t)
            if inverse:
                ref, _ = s.step(vn, t, x)
                got = emulated_ops.ddim_step(x, vn, None, 1.0, ca, cb, cc, cd, inverse=True)
            else:
                ref, _ = s.step(schedulers_ref.cfg_combine(vn, ve, 9.0), t, x)
                got = emulated_ops.ddim_step(x, vn, ve, 9.0, ca, cb, cc, cd)
            assert torch.equal(got, ref)


@torch.no_grad()
@pytest.mark.parametrize("t,expect_inject", [(901, True), (101, False), (1000, True)])
def test_product_unet_and_hooks_match_the_oracle_on_cpu(emulated_ops, t, expect_inject):
    from anyv2v_b200 import pnp_utils as ours_hooks
    from oracle import pnp_hooks_ref, schedulers_ref
    ref32, ours = _models()
    s = schedulers_ref.DDIMScheduler()
    s.set_timesteps(10)
    schedule = s.timesteps[:5]
    outs = {}
    for name, net, dt, hooks in (("ref32", ref32, torch.float32, pnp_hooks_ref), ("ours", ours, torch.float16, ours_hooks)):
        pipe = SimpleNamespace(unet=net)
        hooks.register_conv_injection(pipe, schedule)
        hooks.register_spatial_attention_pnp(pipe, schedule)
        hooks.register_temp_attention_pnp(pipe, schedule)
        hooks.register_time(pipe, t)
        _, x3, prompts, img_lat, img_emb, fps = _inputs(dt)
        outs[name] = net(x3, torch.tensor([t]), fps, img_lat, img_emb, prompts)[0]
    assert outs["ours"].shape == (3, 4, F_, H_, W_)
    _close(outs["ours"], outs["ref32"], f"hooked UNet t={t}")
    proc = ours.up_blocks[2].attentions[1].transformer_blocks[0].attn1.processor
    assert proc.inject_now() == expect_inject
    assert emulated_ops.launch_count() > 0


@torch.no_grad()
def test_unhooked_unet_and_batch_sizes_on_cpu(emulated_ops):
    """B = 1 (inversion), B = 2 (dead-source edit step) and B = 3 run through the same wiring"""
    ref32, ours = _models()
    for b in (1, 2, 3):
        ns32, x3, prompts, img_lat, img_emb, fps = _inputs(torch.float32)
        ref = ref32(x3[:b], torch.tensor([501]), fps[:b], img_lat[:b], img_emb[:b], prompts[:b])[0]
        _, x3h, prompts, img_lat, img_emb, fps = _inputs(torch.float16)
        got = ours(x3h[:b], torch.tensor([501]), fps[:b], img_lat[:b], img_emb[:b], prompts[:b])[0]
        _close(got, ref, f"UNet forward B={b}")


@torch.no_grad()
def test_both_loops_on_cpu_teacher_forced(emulated_ops, tmp_path):
    """pipeline.invert + sample_with_pnp (host-side loop logic, latent store, dead-source steps, fused CFG/DDIM contract)
    against oracle/loops_ref.py, teacher-forced per step"""
    from anyv2v_b200 import pnp_utils as ours_hooks
    from anyv2v_b200.pipeline import I2VGenXLPipeline
    from anyv2v_b200.run_group_pnp_edit import init_pnp
    from anyv2v_b200.schedulers import DDIMInverseScheduler, DDIMScheduler
    from oracle import loops_ref, pnp_hooks_ref, schedulers_ref
    ref32, ours = _models()
    n_steps = 4
    ns32 = loops_ref.synthetic_inputs(F_, H_, W_, cross_dim=64, dtype=torch.float32, device="cpu")
    ns16 = loops_ref.synthetic_inputs(F_, H_, W_, cross_dim=64, dtype=torch.float16, device="cpu")
    inv_ref = loops_ref.invert_loop(ref32, ns32.video_latents, ns32.inv_prompt, ns32.src_image_latents, ns32.src_image_emb, ns32.fps, n_steps)
    pipe = I2VGenXLPipeline(ours, DDIMInverseScheduler())
    stacked = pipe.invert(latents=ns16.video_latents, prompt_embeds=ns16.inv_prompt, image_latents=ns16.src_image_latents,
                          image_embeddings=ns16.src_image_emb, target_fps=8, num_inference_steps=n_steps, guidance_scale=1.0,
                          output_dir=str(tmp_path / "ddim_latents"))
    assert stacked.shape == (1, n_steps, 4, F_, H_, W_)
    store = pipe.latent_store
    ts = sorted(inv_ref)
    # the first step is exactly teacher-forced (same x_0); later ones drift with the fp16 state — compare loosely
    _close(store.get(ts[0], device="cpu"), inv_ref[ts[0]], "inversion step 1", rms=3e-3, mx=1e-2)
    _close(store.get(ts[-1], device="cpu"), inv_ref[ts[-1]], "inversion, free-running", rms=2e-2, mx=6e-2)

    # edit phase: conv on the first 2 of 4 steps, attention on the first step only -> injected, conv-only, dead-source steps
    cfg = SimpleNamespace(n_steps=n_steps, pnp_f_t=0.5, pnp_spatial_attn_t=0.25, pnp_temp_attn_t=0.25)
    edit_sched = DDIMScheduler()
    edit_sched.set_timesteps(n_steps)
    pipe.scheduler = edit_sched
    init_pnp(pipe, edit_sched, cfg)
    seen = []
    out = pipe.sample_with_pnp(latents=ns16.video_latents.clone(), prompt_embeds=ns16.edit_prompt, negative_prompt_embeds=ns16.neg_prompt,
                               ddim_inv_prompt_embeds=ns16.inv_prompt, image_embeddings=ns16.edit_image_emb,
                               image_latents=ns16.edit_image_latents, ddim_inv_image_embeddings=ns16.src_image_emb,
                               ddim_inv_image_latents=ns16.src_image_latents, target_fps=8, num_inference_steps=n_steps,
                               guidance_scale=9.0, ddim_init_latents_t_idx=0, latent_store=store,
                               callback=lambda i, t, x: seen.append((i, t, x.clone())), return_dict=False)[0]
    assert out.shape == (1, 4, F_, H_, W_) and len(seen) == n_steps
    # oracle, teacher-forced: restart every step from OUR previous latent and OUR stored source latent
    sref = schedulers_ref.DDIMScheduler()
    sref.set_timesteps(n_steps)
    rp = SimpleNamespace(unet=ref32)
    pnp_hooks_ref.init_pnp(rp, sref, n_steps, pnp_f_t=0.5, pnp_spatial_attn_t=0.25, pnp_temp_attn_t=0.25)
    prompts, img_lat, img_emb, fps3 = loops_ref.edit_conditioning(ns32)
    x_prev = ns16.video_latents.float()
    for i, t, x_ours in seen:
        pnp_hooks_ref.register_time(rp, t)
        src = store.get(t, device="cpu").float()
        v = ref32(torch.cat([src, x_prev, x_prev]), torch.tensor(t), fps3, img_lat, img_emb, prompts)[0]
        x_ref, _ = sref.step(schedulers_ref.cfg_combine(v[1:2], v[2:3], 9.0), t, x_prev)
        _close(x_ours, x_ref, f"edit step {i} (t={t})", rms=6e-3, mx=3e-2)
        x_prev = x_ours.float()


@torch.no_grad()
@pytest.mark.parametrize("t", [901, 101])
def test_shared_uncond_cond_prefix_is_a_pure_deduplication(emulated_ops, t):
    """shared_edit_prefix (AV2V_SHARED_PREFIX): with identical uncond / cond latents the prefix up to the first
    cross-attention is computed once — same output as the plain forward, for B = 3 (hooks firing or not) and B = 2."""
    from anyv2v_b200 import pnp_utils as ours_hooks
    from oracle import schedulers_ref
    _, ours = _models()
    pipe = SimpleNamespace(unet=ours)
    s = schedulers_ref.DDIMScheduler()
    s.set_timesteps(10)
    for reg in (ours_hooks.register_conv_injection, ours_hooks.register_spatial_attention_pnp, ours_hooks.register_temp_attention_pnp):
        reg(pipe, s.timesteps[:5])
    ours_hooks.register_time(pipe, t)
    _, x3, prompts, img_lat, img_emb, fps = _inputs(torch.float16)
    x3 = torch.cat([x3[:2], x3[1:2]])                        # [source, x, x] as in pipeline :1136
    img_lat = torch.cat([img_lat[:2], img_lat[1:2]])         # [source, edited, edited] (:1099)
    for b0 in (0, 1):                                        # B = 3 and the dead-source B = 2 batch
        args = (x3[b0:], torch.tensor([t]), fps[b0:], img_lat[b0:], img_emb[b0:], prompts[b0:])
        plain = ours(*args)[0]
        n0 = emulated_ops.launch_count()
        ours(*args)
        n_plain = emulated_ops.launch_count() - n0
        n0 = emulated_ops.launch_count()
        shared = ours(*args, shared_edit_prefix=True)[0]
        n_shared = emulated_ops.launch_count() - n0
        assert n_shared == n_plain                          # same kernels, smaller batches in the prefix
        assert not torch.equal(plain[-1], plain[-2])        # the two edit branches do differ (different contexts)
        assert torch.equal(shared, plain)                   # float64 contracts are batch-invariant: bit-identical


@torch.no_grad()
def test_edit_loop_with_shared_prefix_switch(emulated_ops, monkeypatch, tmp_path):
    from anyv2v_b200.latent_store import LatentStore
    from anyv2v_b200.pipeline import I2VGenXLPipeline
    from anyv2v_b200.run_group_pnp_edit import init_pnp
    from anyv2v_b200.schedulers import DDIMScheduler
    from oracle import loops_ref
    _, ours = _models()
    n_steps = 3
    ns = loops_ref.synthetic_inputs(F_, H_, W_, cross_dim=64, dtype=torch.float16, device="cpu")
    sched = DDIMScheduler()
    sched.set_timesteps(n_steps)
    outs = []
    for flag in (False, True):  # False: the reference's full 3-branch batch on every step; True: all de-duplications
        pipe = I2VGenXLPipeline(ours, sched)
        init_pnp(pipe, sched, SimpleNamespace(n_steps=n_steps, pnp_f_t=0.67, pnp_spatial_attn_t=0.34, pnp_temp_attn_t=0.0))
        store = LatentStore(None, write_files=False)
        g = torch.Generator().manual_seed(5)
        for t in sched.timesteps.tolist():
            store.put(int(t), torch.randn(1, 4, F_, H_, W_, generator=g).half())
        out = pipe.sample_with_pnp(latents=ns.video_latents.clone(), prompt_embeds=ns.edit_prompt, negative_prompt_embeds=ns.neg_prompt,
                                   ddim_inv_prompt_embeds=ns.inv_prompt, image_embeddings=ns.edit_image_emb,
                                   image_latents=ns.edit_image_latents, ddim_inv_image_embeddings=ns.src_image_emb,
                                   ddim_inv_image_latents=ns.src_image_latents, target_fps=8, num_inference_steps=n_steps,
                                   guidance_scale=9.0, ddim_init_latents_t_idx=0, latent_store=store, return_dict=False,
                                   skip_dead_source_branch=flag)[0]
        outs.append(out.float())
    assert torch.isfinite(outs[1]).all() and torch.equal(outs[0], outs[1])


@torch.no_grad()
@pytest.mark.parametrize("fracs,site", [((1.0, 0.0, 0.0), (1, 1, "resnet")), ((1.0, 1.0, 0.0), (3, 2, "spatial")),
                                        ((1.0, 1.0, 1.0), (3, 2, "temporal")), ((0.0, 0.0, 1.0), (3, 2, "temporal"))])
def test_source_branch_pruning_keeps_the_edit_branches(emulated_ops, fracs, site):
    """prune_source_after (set by the edit loop): dropping the source branch after its last firing site leaves [uncond, cond] unchanged"""
    from anyv2v_b200 import pnp_utils as ours_hooks
    from anyv2v_b200.pipeline import I2VGenXLPipeline
    from oracle import schedulers_ref
    _, ours = _models()
    pipe = SimpleNamespace(unet=ours)
    s = schedulers_ref.DDIMScheduler()
    s.set_timesteps(10)
    full, none = s.timesteps, []
    ours_hooks.register_conv_injection(pipe, full if fracs[0] else none)
    ours_hooks.register_spatial_attention_pnp(pipe, full if fracs[1] else none)
    ours_hooks.register_temp_attention_pnp(pipe, full if fracs[2] else none)
    t = int(s.timesteps[2])
    ours_hooks.register_time(pipe, t)
    assert I2VGenXLPipeline._prune_site(tuple(bool(f) for f in fracs)) == site
    _, x3, prompts, img_lat, img_emb, fps = _inputs(torch.float16)
    args = (x3, torch.tensor([t]), fps, img_lat, img_emb, prompts)
    plain = ours(*args)[0]
    n0 = emulated_ops.launch_count()
    pruned = ours(*args, prune_source_after=site)[0]
    assert pruned.shape[0] == 2 and emulated_ops.launch_count() - n0 > 0
    assert torch.equal(pruned, plain[1:])  # the float64 contracts are batch-invariant: a pure re-grouping of the same work
    both = ours(torch.cat([x3[:2], x3[1:2]]), torch.tensor([t]), fps, torch.cat([img_lat[:2], img_lat[1:2]]), img_emb, prompts,
                shared_edit_prefix=True, prune_source_after=site)[0]
    ref = ours(torch.cat([x3[:2], x3[1:2]]), torch.tensor([t]), fps, torch.cat([img_lat[:2], img_lat[1:2]]), img_emb, prompts)[0]
    assert torch.equal(both, ref[1:])


@torch.no_grad()
def test_edit_loop_with_every_host_level_switch(emulated_ops, monkeypatch):
    """the whole edit loop (injected, conv-only and dead-source steps) with every parity-preserving de-duplication (dead
    source branch skipped, shared uncond / cond prefix, source pruned after its last firing site) against the reference's
    full 3-branch batch on every step (`skip_dead_source_branch=False`)"""
    from anyv2v_b200.latent_store import LatentStore
    from anyv2v_b200.pipeline import I2VGenXLPipeline
    from anyv2v_b200.run_group_pnp_edit import init_pnp
    from anyv2v_b200.schedulers import DDIMScheduler
    from oracle import loops_ref
    _, ours = _models()
    n_steps = 4
    ns = loops_ref.synthetic_inputs(F_, H_, W_, cross_dim=64, dtype=torch.float16, device="cpu")
    sched = DDIMScheduler()
    sched.set_timesteps(n_steps)
    outs = []
    for dedup in (False, True):
        pipe = I2VGenXLPipeline(ours, sched)
        init_pnp(pipe, sched, SimpleNamespace(n_steps=n_steps, pnp_f_t=0.75, pnp_spatial_attn_t=0.5, pnp_temp_attn_t=0.25))
        store = LatentStore(None, write_files=False)
        g = torch.Generator().manual_seed(5)
        for t in sched.timesteps.tolist():
            store.put(int(t), torch.randn(1, 4, F_, H_, W_, generator=g).half())
        out = pipe.sample_with_pnp(latents=ns.video_latents.clone(), prompt_embeds=ns.edit_prompt, negative_prompt_embeds=ns.neg_prompt,
                                   ddim_inv_prompt_embeds=ns.inv_prompt, image_embeddings=ns.edit_image_emb,
                                   image_latents=ns.edit_image_latents, ddim_inv_image_embeddings=ns.src_image_emb,
                                   ddim_inv_image_latents=ns.src_image_latents, target_fps=8, num_inference_steps=n_steps,
                                   guidance_scale=9.0, ddim_init_latents_t_idx=0, latent_store=store, return_dict=False,
                                   skip_dead_source_branch=dedup)[0]
        outs.append(out.float())
    assert torch.isfinite(outs[1]).all() and torch.equal(outs[0], outs[1])


def test_group_runners_end_to_end_on_cpu(emulated_ops, tmp_path):
    """run_group_ddim_inversion -> ddim_latents_{t}.pt -> run_group_pnp_edit with the reference's template / JSON API, on CPU
    through the kernel contracts (the GPU twin is tests/test_gpu_runners.py)"""
    import os
    import yaml
    from test_gpu_runners import EDIT_TEMPLATE, INV_TEMPLATE
    from anyv2v_b200 import run_group_ddim_inversion as inv, run_group_pnp_edit as edit
    from anyv2v_b200.config import OmegaConf
    from oracle.unet_ref import TINY_CONFIG
    data = str(tmp_path)
    for name, tpl in (("inv.yaml", INV_TEMPLATE), ("edit.yaml", EDIT_TEMPLATE)):
        (tmp_path / name).write_text(yaml.safe_dump(dict(tpl, data_dir=data, device="cpu")))
    entries = [{"active": True, "video_name": "clipA", "edited_first_frame_path": "demo/clipA/edited.png", "editing_prompt": "a robot",
                "edited_video_name": "robot", "ddim_init_latents_t_idx": 0, "pnp_f_t": 1.0, "pnp_spatial_attn_t": 0.4, "pnp_temp_attn_t": 0.2},
               {"active": False, "video_name": "skipped", "edited_first_frame_path": "x", "editing_prompt": "x", "edited_video_name": "x"}]
    prev = torch.is_grad_enabled()
    torch.set_grad_enabled(False)
    try:
        device = torch.device("cpu")
        out = inv.main(OmegaConf.load(str(tmp_path / "inv.yaml")), entries, device, unet_config=TINY_CONFIG)
        assert len(out) == 1 and out[0].shape == (5, 4, 4, 16, 16)
        lat_dir = os.path.join(data, "inversions", "i2vgen-xl", "clipA", "ddim_latents")
        assert sorted(os.listdir(lat_dir)) == sorted(f"ddim_latents_{t}.pt" for t in (1, 201, 401, 601, 801))
        assert inv.main(OmegaConf.load(str(tmp_path / "inv.yaml")), entries, device, unet_config=TINY_CONFIG) == []  # skip rule
        res = edit.main(OmegaConf.load(str(tmp_path / "edit.yaml")), entries, device, unet_config=TINY_CONFIG)
        assert len(res) == 1 and res[0].shape == (1, 4, 4, 16, 16) and torch.isfinite(res[0]).all()
        suffix = "ddim_init_latents_t_idx_0_nsteps_5_cfg_9.0_pnpf1.0_pnps0.4_pnpt0.2"
        saved = os.path.join(data, "Results", "Prompt-Based-Editing", "i2vgen-xl", "clipA", "robot", suffix, "edited_latents.pt")
        assert os.path.exists(saved) and torch.equal(torch.load(saved), res[0].cpu())
    finally:
        torch.set_grad_enabled(prev)


@torch.no_grad()
@pytest.mark.parametrize("t,inject", [(901, True), (101, False)])
def test_temporal_self_attention_goes_through_the_fused_kernel(emulated_ops, monkeypatch, t, inject):
    """every temporal self-attention (attn1 and attn2 of the temporal transformers, transformer_in) is ONE launch of
    ops.temporal_attention_fused; on injected steps the hooked attn1 sites pass n_v = 3 (Q, K from the source clip)"""
    from anyv2v_b200 import ops, pnp_utils as ours_hooks
    from oracle import schedulers_ref
    _, ours = _models()
    pipe = SimpleNamespace(unet=ours)
    s = schedulers_ref.DDIMScheduler()
    s.set_timesteps(10)
    ours_hooks.register_temp_attention_pnp(pipe, s.timesteps[:5])
    ours_hooks.register_time(pipe, t)
    calls = []
    real = ops.temporal_attention_fused

    def spy(*a, **kw):
        calls.append(kw.get("n_v", 1))
        return real(*a, **kw)

    monkeypatch.setattr(ops, "temporal_attention_fused", spy)
    _, x3, prompts, img_lat, img_emb, fps = _inputs(torch.float16)
    out = ours(x3, torch.tensor([t]), fps, img_lat, img_emb, prompts)[0]
    assert torch.isfinite(out).all()
    n_temporal = 1 + sum(len(b.temp_attentions) for b in list(ours.down_blocks) + list(ours.up_blocks) if b.has_cross_attention) + 1
    # attn1 + attn2 of every temporal transformer (transformer_in, mid); short spatial sequences (< 128 tokens) take the same kernel
    assert len(calls) >= 2 * n_temporal
    assert calls.count(3) == (8 if inject else 0)             # the 8 hooked attn1 sites (pnp_utils.py:340-346)
    ours_hooks.register_temp_attention_pnp(pipe, [])
    ours_hooks.register_time(pipe, -1)


@torch.no_grad()
def test_fullwidth_gpu_test_module_runs_on_cpu_with_the_tiny_config(emulated_ops, monkeypatch):
    """tests/test_gpu_fullwidth.py (the BASELINE-width GPU parity module) re-run here with the topology-equivalent tiny
    config on the float64 kernel contracts: exercises every line of its host-side wiring without a GPU."""
    import test_gpu_fullwidth as fw
    from oracle import unet_ref
    monkeypatch.setattr(fw, "CONFIG_OVERRIDE", unet_ref.TINY_CONFIG)
    monkeypatch.setattr(fw, "dev", "cpu")
    monkeypatch.setattr(fw, "H_", 16)
    monkeypatch.setattr(fw, "W_", 16)
    monkeypatch.setattr(fw._inputs, "__defaults__", (fw.F_, 16, 16))
    full = fw.build_models("cpu")
    fw.test_fullwidth_hooked_step_matches_oracle(full, 901, True)
    fw.test_fullwidth_unhooked_forward_and_inversion_batch(full)
    fw.test_fullwidth_teacher_forced_edit_and_inversion_steps(full)
    fw.test_fullwidth_finest_level_block_alone(full, True)


def test_group_runners_on_real_inputs_on_cpu(emulated_ops, tmp_path):
    """the REAL input path of both runners (png frames, prompt strings, edited first frame -> VAE / CLIP -> loops -> VAE decode
    -> png / gif) on CPU through the kernel contracts; the GPU twin is tests/test_gpu_runners.py"""
    from test_gpu_runners import run_real_input_runners
    prev = torch.is_grad_enabled()
    torch.set_grad_enabled(False)
    try:
        run_real_input_runners(tmp_path, torch.device("cpu"), cfg_inv=2.0)  # also: inversion with guidance
    finally:
        torch.set_grad_enabled(prev)

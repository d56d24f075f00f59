"""CPU tests: the oracle against the known-answer material the reference holds and against the
golden vectors frozen by oracle/make_golden.py from the reference's own hook / scheduler code."""
import math
import os
from types import SimpleNamespace

import pytest
import torch

from oracle import loops_ref, pnp_hooks_ref, schedulers_ref, unet_ref
from oracle.make_golden import run_hooks


def test_timesteps_kat():
    # demo.ipynb:1201-1204 (50 steps) and :498,997 (500-step inversion)
    s = schedulers_ref.DDIMScheduler()
    s.set_timesteps(50)
    assert s.timesteps.tolist() == list(range(981, 0, -20))
    inv = schedulers_ref.DDIMInverseScheduler()
    inv.set_timesteps(50)
    assert inv.timesteps.tolist() == list(range(1, 982, 20))
    inv.set_timesteps(500)
    assert inv.timesteps[0] == 1 and inv.timesteps[-1] == 999 and len(inv.timesteps) == 500


def test_alphas_cumprod_kat(golden_dir):
    a = schedulers_ref.alphas_cumprod()
    kat = {0: 9.999587536e-01, 1: 9.999126196e-01, 21: 9.979711771e-01, 41: 9.940957427e-01, 481: 5.218712687e-01,
           501: 4.907063246e-01, 961: 3.497560509e-03, 981: 7.840304170e-04}
    for i, v in kat.items():
        assert abs(float(a[i]) - v) < 1e-9
    assert float(a[999]) == 0.0
    g = torch.load(os.path.join(golden_dir, "scheduler_kat.pt"))
    assert torch.equal(a[g["alphas_idx"]], g["alphas"])


def test_inverse_step_matches_reference_golden(golden_dir):
    g = torch.load(os.path.join(golden_dir, "scheduler_kat.pt"))
    s = schedulers_ref.DDIMInverseScheduler()
    for n in (10, 50, 500):
        s.set_timesteps(n)
        assert torch.equal(s.timesteps, g[f"timesteps_{n}"])
        for t in (int(s.timesteps[0]), int(s.timesteps[n // 2]), int(s.timesteps[-1])):
            out, _ = s.step(g["v"], t, g["x"])
            torch.testing.assert_close(out, g[f"inv_step_{n}_{t}"], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("cls", [schedulers_ref.DDIMScheduler, schedulers_ref.DDIMInverseScheduler])
def test_step_closed_form(cls):
    # Appendix C.6: feed a consistent (x, v) pair -> the step must land exactly on the target noise level
    s = cls()
    s.set_timesteps(50)
    torch.manual_seed(0)
    x0 = torch.randn(4, 4, 8, 8, dtype=torch.float64)
    eps = torch.randn(4, 4, 8, 8, dtype=torch.float64)
    for t in (981, 501, 21):
        ca, cb, cc, cd = s.coefficients(t)
        x = ca * x0 + cb * eps
        v = ca * eps - cb * x0
        out, pred_x0 = s.step(v, t, x)
        torch.testing.assert_close(pred_x0, x0, rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(out, cc * x0 + cd * eps, rtol=1e-5, atol=1e-5)


def test_init_pnp_truncation():
    # run_group_pnp_edit.py:36-38 float truncation
    assert [int(50 * f) for f in (0.8, 0.5, 0.2, 0.58)] == [40, 25, 10, 28]
    s = schedulers_ref.DDIMScheduler()
    s.set_timesteps(50)
    pipe = SimpleNamespace(unet=unet_ref.seeded_unet(unet_ref.TINY_CONFIG))
    pnp_hooks_ref.init_pnp(pipe, s, 50, 0.8, 0.5, 0.2)
    res = pipe.unet.up_blocks[1].resnets[1]
    assert res.injection_schedule.tolist() == list(range(981, 981 - 40 * 20, -20))
    sp = pipe.unet.up_blocks[3].attentions[2].transformer_blocks[0].attn1.processor
    tp = pipe.unet.up_blocks[2].temp_attentions[0].transformer_blocks[0].attn1.processor
    assert len(sp.injection_schedule) == 25 and len(tp.injection_schedule) == 10
    # un-patched sites keep the default processor (pnp_utils.py:235 skips up_blocks[1].attentions[0])
    assert isinstance(pipe.unet.up_blocks[1].attentions[0].transformer_blocks[0].attn1.processor, unet_ref.AttnProcessor2_0)


def test_parameter_count_full_model():
    with torch.device("meta"):
        net = unet_ref.I2VGenXLUNet(**unet_ref.I2VGEN_XL_CONFIG)
    n = sum(p.numel() for p in net.parameters())
    assert abs(n / 1e9 - 1.420) < 0.001, n  # public fp16 checkpoint = 2.84 GB


def test_timestep_embedding_closed_form():
    t = torch.tensor([0, 1, 981])
    e = unet_ref.timestep_embedding(t, 320)
    assert e.shape == (3, 320)
    assert torch.allclose(e[0, :160], torch.ones(160)) and torch.allclose(e[0, 160:], torch.zeros(160))
    f5 = math.exp(-math.log(10000.0) * 5 / 160)
    assert abs(float(e[2, 5]) - math.cos(981 * f5)) < 5e-4 and abs(float(e[2, 165]) - math.sin(981 * f5)) < 5e-4  # fp32 argument


def test_restated_hooks_match_reference_golden(golden_dir):
    """tests/golden/tiny_unet_pnp.pt was produced by the UNMODIFIED reference pnp_utils.py (oracle/make_golden.py)."""
    g = torch.load(os.path.join(golden_dir, "tiny_unet_pnp.pt"))
    for name, t in (("injected", 901), ("not_injected", 101), ("t1000", 1000)):
        out = run_hooks(pnp_hooks_ref, t, g["schedule"])
        torch.testing.assert_close(out, g[name], rtol=1e-4, atol=1e-5)


def test_hook_invariants():
    # Appendix C.5: injected => identical attention probabilities / conv features across the 3 branches
    torch.manual_seed(0)
    attn = unet_ref.Attention(128, None, 2, 64)
    proc = pnp_hooks_ref.PnPAttnProcessor(torch.tensor([901]))
    proc.t = 901
    x = torch.randn(3, 16, 128)
    x[1] = x[0]
    x[2] = x[0] * 1.0
    v_same = proc(attn, x.clone())
    assert torch.allclose(v_same[0], v_same[1]) and torch.allclose(v_same[0], v_same[2])
    proc.t = 101  # not scheduled -> plain attention
    y = torch.randn(3, 16, 128)
    assert torch.allclose(proc(attn, y.clone()), unet_ref.AttnProcessor2_0()(attn, y.clone()))
    proc.t = 1000  # forced
    z = proc(attn, y.clone())
    assert not torch.allclose(z[1], unet_ref.AttnProcessor2_0()(attn, y.clone())[1])


def test_tiny_inversion_then_edit_runs():
    """BASELINE config 0 in miniature: inversion loop then PnP edit loop, CPU, plumbing only."""
    net = unet_ref.seeded_unet(unet_ref.TINY_CONFIG)
    ns = loops_ref.synthetic_inputs(4, 16, 16, cross_dim=64)
    inv = loops_ref.invert_loop(net, ns.video_latents, ns.inv_prompt, ns.src_image_latents, ns.src_image_emb, ns.fps, 3)
    assert sorted(inv) == [1, 334, 667]
    pipe = SimpleNamespace(unet=net)
    s = schedulers_ref.DDIMScheduler()
    s.set_timesteps(3)
    pnp_hooks_ref.init_pnp(pipe, s, 3, 1.0, 1.0, 0.5)
    prompts, img_lat, img_emb, fps = loops_ref.edit_conditioning(ns)
    out = loops_ref.pnp_edit_loop(pipe, pnp_hooks_ref.register_time, inv, inv[667].clone(), prompts, img_lat, img_emb, fps, 3, 9.0)
    assert out.shape == (1, 4, 4, 16, 16) and torch.isfinite(out).all()


# ---------------------------------------------------------------------------------------------- VAE
@torch.no_grad()
def test_vae_restatement_structure_and_loop_brackets():
    """Structural known answers of the AutoencoderKL restatement (parity unpinned: diffusers is absent): the parameter
    count of the public SD KL-f8 VAE, state_dict names shared with the product module, the f8 geometry, and that the
    reference's frame-chunked decode (`decode_chunk_size=1`, pipeline_i2vgen_xl.py:449-454) equals the batched one."""
    from oracle import vae_ref
    from anyv2v_b200 import vae as product
    full = vae_ref.AutoencoderKL(**vae_ref.SD_VAE_CONFIG)
    assert sum(p.numel() for p in full.parameters()) == 83_653_863
    ours = product.AutoencoderKL(**product.SD_VAE_CONFIG)
    sd_ref, sd_ours = full.state_dict(), ours.state_dict()
    assert list(sd_ref) == list(sd_ours)
    assert all(sd_ref[k].shape == sd_ours[k].shape for k in sd_ref)
    assert "decoder.mid_block.attentions.0.to_q.weight" in sd_ref and "encoder.down_blocks.0.downsamplers.0.conv.weight" in sd_ref

    tiny = vae_ref.seeded_vae(vae_ref.TINY_VAE_CONFIG, seed=8888)
    g = torch.Generator().manual_seed(1)
    lat = torch.randn(1, 4, 3, 6, 5, generator=g)
    v_all = vae_ref.decode_latents(tiny, lat, None)
    v_one = vae_ref.decode_latents(tiny, lat, 1)
    assert v_all.shape == (1, 3, 3, 12, 10) and v_all.dtype == torch.float32
    torch.testing.assert_close(v_all, v_one, rtol=1e-5, atol=1e-5)
    frames = torch.randn(3, 3, 16, 12, generator=g).clamp(-1, 1)
    z1 = vae_ref.encode_vae_video(tiny, frames, torch.Generator().manual_seed(7))
    z2 = vae_ref.encode_vae_video(tiny, frames, torch.Generator().manual_seed(7))
    assert z1.shape == (1, 4, 3, 8, 6) and torch.equal(z1, z2)
    d = tiny.encode(frames[:1]).latent_dist
    assert float(d.logvar.max()) <= 20.0 and float(d.logvar.min()) >= -30.0
    torch.testing.assert_close(d.std, torch.exp(0.5 * d.logvar))

"""GPU parity tests for the steps either side of the loops: the product AutoencoderKL on the
hand-written kernels against the oracle restatement (fp32 and fp16 torch on the same GPU), and the wide-image path of the
implicit-GEMM convolution it relies on."""
import pytest
import torch

from parity_utils import assert_fp16_close, err_stats

pytestmark = pytest.mark.gpu
dev = "cuda"


@pytest.fixture(scope="module")
def ops():
    from anyv2v_b200 import ops as o
    return o


@pytest.mark.parametrize("geo", [(1, 4, 256, 64, 64), (2, 3, 512, 128, 128), (1, 5, 384, 64, 192)])
def test_conv3x3_wide_images(ops, geo):
    """W > 128: tiles are 128-pixel segments of one image row (VAE resolutions)."""
    NF, H, W, Cin, Cout = geo
    torch.manual_seed(5)
    x = torch.randn(NF, H, W, Cin, device=dev).half()
    w = (torch.randn(Cout, Cin, 3, 3, device=dev) / (9 * Cin) ** 0.5).half()
    bias = torch.randn(Cout, device=dev).half()
    res = torch.randn(NF, H, W, Cout, device=dev).half()
    out = ops.conv3x3(x, w.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous(), bias=bias, residual=res)
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).float(), w.float(), bias.float(), padding=1).permute(0, 2, 3, 1) + res.float()
    assert_fp16_close(out, ref, f"conv3x3 wide {geo}")


@pytest.fixture(scope="module")
def vaes():
    from types import SimpleNamespace
    from anyv2v_b200 import vae as product
    from oracle import vae_ref
    cfg = dict(vae_ref.TINY_VAE_CONFIG)
    ref32 = vae_ref.seeded_vae(cfg, seed=8888, dtype=torch.float32).to(dev)
    ref16 = vae_ref.seeded_vae(cfg, seed=8888, dtype=torch.float16).to(dev)
    ours = product.AutoencoderKL(**cfg)
    ours.load_state_dict(ref32.state_dict())
    ours = ours.to(device=dev, dtype=torch.float16).eval()
    return SimpleNamespace(ref32=ref32, ref16=ref16, ours=ours)


def _as_close_as_fp16_torch(got, ref32, ref16, what, slack=3.0):
    assert torch.isfinite(got).all(), what
    e_ours, e_ref = err_stats(got, ref32), err_stats(ref16, ref32)
    assert e_ours["rms_rel"] <= max(slack * e_ref["rms_rel"], 2e-3), (what, e_ours, e_ref)


@torch.no_grad()
def test_vae_decode_matches_oracle(vaes):
    from anyv2v_b200 import vae as product
    from oracle import vae_ref
    g = torch.Generator().manual_seed(11)
    lat = torch.randn(1, 4, 3, 16, 32, generator=g).to(dev)          # 3 frames, 16 x 32 latents -> 32 x 64 ... x2 levels
    ref32 = vae_ref.decode_latents(vaes.ref32, lat.float(), 1)
    ref16 = vae_ref.decode_latents(vaes.ref16, lat.half(), 1)
    got = product.decode_latents(vaes.ours, lat.half(), None)
    assert got.shape == ref32.shape and got.dtype == torch.float32
    _as_close_as_fp16_torch(got, ref32, ref16, "vae decode")
    got1 = product.decode_latents(vaes.ours, lat.half(), 1)             # the reference's chunking gives the same frames
    assert err_stats(got1, got)["rms_rel"] < 2e-3


@torch.no_grad()
def test_vae_encode_matches_oracle(vaes):
    from anyv2v_b200 import vae as product
    from oracle import vae_ref
    g = torch.Generator().manual_seed(12)
    frames = torch.randn(2, 3, 32, 256, generator=g).clamp(-1, 1).to(dev)   # W = 256 exercises the wide-image conv path
    d32 = vaes.ref32.encode(frames.float()).latent_dist
    d16 = vaes.ref16.encode(frames.half()).latent_dist
    dours = vaes.ours.encode(frames.half()).latent_dist
    _as_close_as_fp16_torch(dours.mean, d32.mean, d16.mean, "vae posterior mean")
    _as_close_as_fp16_torch(dours.logvar, d32.logvar, d16.logvar, "vae posterior logvar")
    z = product.encode_vae_video(vaes.ours, frames.half(), torch.Generator(device=dev).manual_seed(3))
    assert z.shape == (1, 4, 2, 16, 128) and torch.isfinite(z).all()
    # same generator state -> same posterior draws as the per-frame reference order
    zr = vae_ref.encode_vae_video(vaes.ours, frames.half(), torch.Generator(device=dev).manual_seed(3))
    assert err_stats(z, zr)["rms_rel"] < 2e-3

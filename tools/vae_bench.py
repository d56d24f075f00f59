"""Full-size AutoencoderKL (SD KL-f8, random init) around the loops: 16 frames x 512^2 decode / encode on the
hand-written kernels vs the same module tree in plain torch fp16 (cuDNN / cuBLAS / SDPA) on the same GPU."""
import sys
import torch

sys.path.insert(0, ".")
from anyv2v_b200 import vae as product  # noqa: E402
from oracle import vae_ref  # noqa: E402  (timing comparison only: development tool, not the product path)
from tools.gpu_check import timeit  # noqa: E402

dev = "cuda"
torch.set_grad_enabled(False)
ref = vae_ref.seeded_vae(vae_ref.SD_VAE_CONFIG, seed=8888, dtype=torch.float16).to(dev)
ours = product.AutoencoderKL(**product.SD_VAE_CONFIG)
ours.load_state_dict(ref.state_dict())
ours = ours.to(device=dev, dtype=torch.float16).eval()
g = torch.Generator().manual_seed(0)
lat = (torch.randn(1, 4, 16, 64, 64, generator=g) * 0.18215).to(dev).half()
frames = torch.randn(16, 3, 512, 512, generator=g).clamp(-1, 1).to(dev).half()

v_ours = product.decode_latents(ours, lat, None)
v_ref = vae_ref.decode_latents(ref, lat, 1)
e = (v_ours - v_ref).double()
print(f"decode 16f x 512^2: rms rel diff ours-vs-torch-fp16 {float(e.pow(2).mean().sqrt() / v_ref.double().pow(2).mean().sqrt()):.3e}", flush=True)
t_o = timeit(lambda: product.decode_latents(ours, lat, None), iters=3, warm=1)
t_r = timeit(lambda: vae_ref.decode_latents(ref, lat, 1), iters=3, warm=1)
t_r16 = timeit(lambda: vae_ref.decode_latents(ref, lat, None), iters=3, warm=1)
print(f"decode: ours {t_o*1e3:.1f} ms | torch fp16 chunk=1 (reference setting) {t_r*1e3:.1f} ms | torch fp16 batched {t_r16*1e3:.1f} ms", flush=True)
m_o = ours.encode(frames).latent_dist.mean
m_r = torch.cat([ref.encode(frames[i:i + 1]).latent_dist.mean for i in range(16)])
e = (m_o - m_r).double()
print(f"encode 16f x 512^2: rms rel diff of the posterior mean {float(e.pow(2).mean().sqrt() / m_r.double().pow(2).mean().sqrt()):.3e}", flush=True)
t_o = timeit(lambda: ours.encode(frames), iters=3, warm=1)
t_r = timeit(lambda: [ref.encode(frames[i:i + 1]) for i in range(16)], iters=3, warm=1)
print(f"encode: ours {t_o*1e3:.1f} ms | torch fp16 per frame (reference setting) {t_r*1e3:.1f} ms", flush=True)
print(f"peak memory {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")

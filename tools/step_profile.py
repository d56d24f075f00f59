"""In-situ per-op GPU time of one inversion step and one PnP edit step (eager, warm) — development aid."""
import sys
import time
from types import SimpleNamespace

import torch

sys.path.insert(0, ".")
from anyv2v_b200 import distributed, ops  # noqa: E402
from anyv2v_b200.pipeline import I2VGenXLPipeline  # noqa: E402
from anyv2v_b200.run_group_pnp_edit import init_pnp, synthetic_conditioning  # noqa: E402
from anyv2v_b200.schedulers import DDIMInverseScheduler, DDIMScheduler  # noqa: E402
from anyv2v_b200.unet_i2vgen_xl import I2VGEN_XL_CONFIG, I2VGenXLUNet  # noqa: E402

dev = torch.device("cuda", 0)
torch.set_grad_enabled(False)
I2VGenXLPipeline.use_cuda_graphs = False
unet = distributed.build_unet_replicated(I2VGenXLUNet, I2VGEN_XL_CONFIG, 8888, dev)
pipe = I2VGenXLPipeline(unet, DDIMInverseScheduler())
c = synthetic_conditioning(16, 64, 64, 1024, 8888, dev)
st_inv = pipe.prepare_invert(c["video_latents"], c["inv_prompt"], c["src_image_latents"], c["src_image_emb"], 8, 50, 1.0, None, False)
inv_sched = pipe.scheduler
es = DDIMScheduler()
es.set_timesteps(50)
pipe.scheduler = es
init_pnp(pipe, es, SimpleNamespace(n_steps=50, pnp_f_t=1.0, pnp_spatial_attn_t=1.0, pnp_temp_attn_t=float(sys.argv[1]) if len(sys.argv) > 1 else 0.0))
for t in es.timesteps.tolist()[:4]:
    st_inv.store._mem[int(t)] = torch.randn(1, 4, 16, 64, 64, device=dev).half()
st_edit = pipe.prepare_edit(c["video_latents"].clone(), c["edit_prompt"], c["neg_prompt"], c["inv_prompt"], c["edit_image_emb"],
                            c["edit_image_latents"], c["src_image_emb"], c["src_image_latents"], 8, 50, 9.0, 0, None, st_inv.store, True)


def timed(name, fn):
    fn()  # warm
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ops.profile_begin()
    e0.record()
    fn()
    e1.record()
    prof = ops.profile_end()
    total = e0.elapsed_time(e1)
    ours = sum(v[1] for v in prof.values())
    print(f"=== {name}: step {total:.2f} ms (eager, incl. host gaps); inside our kernels {ours:.2f} ms; other (torch ops + gaps) {total-ours:.2f} ms")
    groups = {}
    for k, (n, ms) in prof.items():
        g = k.split()[0]
        a = groups.setdefault(g, [0, 0.0])
        a[0] += n
        a[1] += ms
    for g, (n, ms) in sorted(groups.items(), key=lambda kv: -kv[1][1]):
        print(f"   {g:22s} n={n:4d} {ms:8.2f} ms")
    for k, (n, ms) in list(prof.items())[:28]:
        print(f"      {ms:7.3f} ms n={n:3d} avg {ms/n*1e3:8.1f} us  {k}")


pipe.scheduler = inv_sched
timed("inversion step (B=1)", lambda: pipe.invert_step(st_inv, 0))
pipe.scheduler = es
timed("PnP edit step (B=3, conv+spatial injected)", lambda: pipe.edit_step(st_edit, 0))

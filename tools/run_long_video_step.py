"""BASELINE config 5 sanity: one inversion step + one PnP edit step (all three injections) of a 128-frame 512x512 clip."""
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

F = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda", 0)
torch.set_grad_enabled(False)
unet = distributed.build_unet_replicated(I2VGenXLUNet, I2VGEN_XL_CONFIG, 8888, dev)
pipe = I2VGenXLPipeline(unet, DDIMInverseScheduler())
c = synthetic_conditioning(F, 64, 64, 1024, 8888, dev)
st_inv = pipe.prepare_invert(c["video_latents"], c["inv_prompt"], c["src_image_latents"], c["src_image_emb"], 8, 50, 1.0, None, False)
inv_sched = pipe.scheduler
es = DDIMScheduler()
es.set_timesteps(50)
pipe.scheduler = es
init_pnp(pipe, es, SimpleNamespace(n_steps=50, pnp_f_t=1.0, pnp_spatial_attn_t=1.0, pnp_temp_attn_t=1.0))
for t in es.timesteps.tolist()[:8]:
    st_inv.store._mem[int(t)] = torch.randn(1, 4, F, 64, 64, device=dev).half()
st_edit = pipe.prepare_edit(c["video_latents"].clone(), c["edit_prompt"], c["neg_prompt"], c["inv_prompt"], c["edit_image_emb"],
                            c["edit_image_latents"], c["src_image_emb"], c["src_image_latents"], 8, 50, 9.0, 0, None, st_inv.store, True)


def timed(name, fn, n=2):
    fn()
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1)/n:.1f} ms/step  peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)


i = [0, 0]
def inv():
    pipe.scheduler = inv_sched
    pipe.invert_step(st_inv, i[0]); i[0] += 1
def edit():
    pipe.scheduler = es
    pipe.edit_step(st_edit, i[1]); i[1] += 1
timed(f"inversion step, {F} frames (B=1)", inv)
timed(f"PnP edit step, {F} frames (B=3, conv+spatial+temporal injection)", edit)
print("finite:", bool(torch.isfinite(st_inv.latents).all() and torch.isfinite(st_edit.latents).all()))

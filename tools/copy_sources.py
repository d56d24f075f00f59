"""Where do the library copy kernels of a step come from?  One eager PnP edit step + one inversion step under torch.profiler with
Python stacks: every aten op that launched a copy / cat / elementwise library kernel, grouped by its innermost anyv2v_b200 frame.
  python tools/copy_sources.py     (B200; development aid)"""
import collections
import runpy
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.argv = [sys.argv[0], "0.0"]
ns = runpy.run_path("tools/step_profile.py")  # builds the model / states and runs one warm step of each kind
pipe, st_inv, st_edit, es, inv_sched = ns["pipe"], ns["st_inv"], ns["st_edit"], ns["es"], ns["inv_sched"]
for name, fn, sched in (("inversion step", lambda: pipe.invert_step(st_inv, 1), inv_sched), ("edit step", lambda: pipe.edit_step(st_edit, 1), es)):
    pipe.scheduler = sched
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=True) as prof:
        fn()
        torch.cuda.synchronize()
    by = collections.defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type != torch.autograd.DeviceType.CPU or not ev.name.startswith("aten::"):
            continue
        cuda_us = sum(k.duration for k in ev.kernels) if ev.kernels else 0.0
        if cuda_us <= 0:
            continue
        frame = next((f for f in ev.stack if "anyv2v_b200" in f), ev.stack[0] if ev.stack else "?")
        key = (ev.name, frame.strip()[-90:], str(ev.input_shapes)[:80])
        by[key][0] += 1
        by[key][1] += cuda_us
    tot = sum(v[1] for v in by.values())
    print(f"=== {name}: library (aten) kernels {tot / 1e3:.3f} ms")
    for k, (n, us) in sorted(by.items(), key=lambda kv: -kv[1][1])[:22]:
        print(f"  {us / 1e3:7.3f} ms n={n:3d}  {k[0]:28s} {k[1]}  {k[2]}")

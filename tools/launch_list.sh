#!/bin/bash
# per-launch device times of an inversion step + a PnP edit step (eager; each step runs twice: warm + measured) under
# `ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised launches: compare SHARES, not absolutes)
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/launches.csv \
  python tools/step_profile.py ${1:-0.0} > gpurun_out/launches_step.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/launches.csv") if l.startswith('"')))
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
t = collections.defaultdict(float); n = collections.Counter()
for r in rows[1:]:
    try: v = float(r[iv].replace(",", ""))
    except ValueError: continue
    name = r[ik][:100]
    t[name] += v; n[name] += 1
tot = sum(t.values())
ours = sum(v for k, v in t.items() if any(s in k for s in ("gemm_tcgen05", "attn", "gn_persistent", "layernorm", "ddim_step")))
print(f"total {tot/1e6:.1f} ms over {sum(n.values())} launches (model init + conditioning + 2 x inversion step + 2 x PnP edit step, eager); "
      f"this package's kernels {ours/1e6:.1f} ms = {100*ours/tot:.1f} %")
for k, v in sorted(t.items(), key=lambda kv: -kv[1])[:45]:
    print(f"{v/1e6:9.2f} ms {100*v/tot:5.1f}%  n={n[k]:5d}  avg {v/n[k]/1e3:8.1f} us  {k}")
PY

#!/bin/bash
# per-launch device times of one inversion + one edit step (cold-cache, serialised: compare SHARES)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/launches.csv") if l.startswith('"')))
hdr = rows[0]
ik, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID")
t = collections.defaultdict(float); n = collections.Counter()
for r in rows[1:]:
    try: v = float(r[iv].replace(",", ""))
    except ValueError: continue
    name = r[ik][:90]
    t[name] += v; n[name] += 1
tot = sum(t.values())
print(f"total {tot/1e6:.1f} ms over {sum(n.values())} launches")
for k, v in sorted(t.items(), key=lambda kv: -kv[1])[:40]:
    print(f"{v/1e6:9.2f} ms {100*v/tot:5.1f}%  n={n[k]:5d}  avg {v/n[k]/1e3:8.1f} us  {k}")
PY

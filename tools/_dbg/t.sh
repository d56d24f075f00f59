timeout 900 python bench.py --gpus 1 --steps 10 --warmup 4 2>&1 | tail -1 > gpurun_out/r02_final_bench.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_final_bench.json").read())
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"])
for r in d["roofline_more"]: print(r["kernel"][:60], r["bound"], r["achieved"], r["peak"], r["frac"], r["us_per_launch"], r["traffic"])
PY

# round-end check on the GPU box: pytest -m gpu, smoke(), bench.py (1 GPU, reference arm, 128 frames)
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 4 2>&1 | tail -1 > gpurun_out/r02_final_bench.json
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/r02_final_ref.json
timeout 900 python bench.py --gpus 1 --steps 4 --warmup 3 --frames 128 2>&1 | tail -1 > gpurun_out/r02_final_f128.json
python - <<'PY'
import json
for f in ("r02_final_bench","r02_final_ref","r02_final_f128"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read())
        print(f, d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("clocks"), d.get("gpu_launches"))
    except Exception as e: print(f, "ERR", e, open(f"gpurun_out/{f}.json").read()[-300:])
PY

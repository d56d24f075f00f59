timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python tools/kernel_bench.py 2>&1 | grep -i "^linear\|^conv"
timeout 600 python bench.py --steps 10 --warmup 4 2>&1 | tail -1 > gpurun_out/r02q_bench.txt; python -c "
import json; d=json.loads(open('gpurun_out/r02q_bench.txt').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['sub_records']['config3']['value'], d['sub_records']['config3']['ms_per_edit_step'], d['sub_records']['config3']['ms_per_inversion_step'])"

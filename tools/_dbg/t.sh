for i in 1 2; do
for lib in tools/_dbg/libprev.so anyv2v_b200/lib/libanyv2v_b200.so; do
AV2V_LIB=$lib timeout 600 python bench.py --steps 10 --warmup 4 2>&1 | tail -1 > gpurun_out/ab_bench.txt; python -c "
import json,sys; d=json.loads(open('gpurun_out/ab_bench.txt').read()); print('$lib'.split('/')[-1], d['value'], d['clocks']['sm_mhz'], d['sub_records']['config3']['ms_per_edit_step'], d['sub_records']['config3']['ms_per_inversion_step'])"
done; done

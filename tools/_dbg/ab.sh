for i in 1 2 3; do
AV2V_LIB=tools/_dbg/libold.so timeout 300 python tools/kernel_bench.py 2>&1 | grep -i "conv3x3" | sed "s/TF frac.*//" > gpurun_out/ab_old_$i.txt
AV2V_LIB=tools/_dbg/libanyv2v_b200_timers.so timeout 300 python tools/kernel_bench.py 2>&1 | grep -i "conv3x3" | sed "s/.*://; s/TF frac.*//" > gpurun_out/ab_new_$i.txt
AV2V_GEMM_DEBUG=512 AV2V_LIB=tools/_dbg/libanyv2v_b200_timers.so timeout 300 python tools/kernel_bench.py 2>&1 | grep -i "conv3x3" | sed "s/.*://; s/TF frac.*//" > gpurun_out/ab_gen_$i.txt
done
echo "shape | old new generic x3"
paste -d"|" gpurun_out/ab_old_1.txt gpurun_out/ab_new_1.txt gpurun_out/ab_gen_1.txt gpurun_out/ab_old_2.txt gpurun_out/ab_new_2.txt gpurun_out/ab_gen_2.txt gpurun_out/ab_old_3.txt gpurun_out/ab_new_3.txt gpurun_out/ab_gen_3.txt | sed "s/ us *[0-9.]* *//g"

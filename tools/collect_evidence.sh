# round-end evidence: ncu --set full of the seven kernel shapes of tools/ncu_one.py, the kernel bench and the launch list (run under gpurun)
for op in gn tattn attn3 attn1 lin960 linres geglu; do
  k=regex:gn_persistent
  case $op in tattn) k=regex:tattn_fused;; attn3) k=regex:attn_pnp;; attn1) k=regex:attn2q;; lin960|linres|geglu) k=regex:gemm_tcgen05;; esac
  timeout 250 ncu --set full --clock-control none --import-source on -k $k -s 2 -c 1 -o gpurun_out/r02z_$op python tools/ncu_one.py $op > /dev/null 2>&1
done
ls gpurun_out/r02z_*
timeout 300 python tools/kernel_bench.py > gpurun_out/r02z_kernel_bench.txt 2>&1
bash tools/launch_list.sh > gpurun_out/r02z_launches.txt 2>&1
tail -3 gpurun_out/r02z_kernel_bench.txt; head -3 gpurun_out/r02z_launches.txt

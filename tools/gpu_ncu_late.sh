# ncu --set full captures of the kernels added late in round 2 (second training step of tools/one_step.py), summarised on the box
mkdir -p gpurun_out/ncu_late
for k in heads_fwd_mma_kernel heads_bwd_mma_kernel conv_cin1_mma_kernel wgrad_cin1_mma_kernel loss_sums_kernel loss_backward_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/ncu_late/$k python tools/one_step.py 32 192 2 > gpurun_out/ncu_late/$k.log 2>&1
  timeout 120 python tools/ncu_summary.py gpurun_out/ncu_late/$k.ncu-rep > gpurun_out/ncu_late/$k.txt 2>&1
  rm -f gpurun_out/ncu_late/$k.ncu-rep
  head -12 gpurun_out/ncu_late/$k.txt | grep "duration\|dram__bytes\|issue_active\|registers"
done
du -sh gpurun_out/ncu_late

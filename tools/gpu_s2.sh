set -x
mkdir -p gpurun_out/s2
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/s2/pytest.log
timeout 300 python tools/profile_layers.py 32 192 bf16 > gpurun_out/s2/layers.txt 2>&1
for shape in "32 256 256 24 24 3" "32 512 512 12 12 3" "32 1024 1024 6 6 3" "32 512 256 24 24 3"; do
  for mode in 0 1; do
    FU_TC_BN_MAX=256 timeout 120 python tools/conv_time.py $shape $mode 2>&1 | grep TFLOP
  done
done > gpurun_out/s2/conv_time_bn256.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"small_cin|heads_|act_bwd|bn_bwd_reduce|bn_finalize_apply|channel_sum|nchw" --launch-skip 0 -c 80 -o gpurun_out/s2/simt python tools/one_step.py 32 192 1 > gpurun_out/s2/ncu.log 2>&1
ncu -i gpurun_out/s2/simt.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__average_warp_latency_issue_stalled_long_scoreboard.pct,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio > gpurun_out/s2/simt_raw.csv 2>&1

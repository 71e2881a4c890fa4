set -x
mkdir -p gpurun_out/s8
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/s8/pytest.log
timeout 300 python tools/step_profile.py > gpurun_out/s8/step_profile.txt 2>&1
FU_TC_WGRAD_WAVES=1 FU_TC_WGRAD3_WAVES=1 timeout 300 python tools/step_profile.py > gpurun_out/s8/step_profile_waves1.txt 2>&1
FU_TC_WGRAD_WAVES=3 FU_TC_WGRAD3_WAVES=3 timeout 300 python tools/step_profile.py > gpurun_out/s8/step_profile_waves3.txt 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"tc_conv|tc_wgrad" -c 12 -o gpurun_out/s8/tc python tools/conv_shapes.py "32 1024 1024 6 6 3 0" "32 512 512 12 12 3 0" "32 512 512 12 12 3 2" "32 256 256 24 24 3 2" "32 32 32 192 192 3 0" "32 32 32 192 192 3 2" "32 64 32 192 192 3 2" "32 128 128 48 48 3 0" "32 128 128 48 48 3 2" "32 64 32 192 192 1 0" > gpurun_out/s8/ncu_tc.log 2>&1
bash tools/ncu_csv.sh gpurun_out/s8/tc.ncu-rep gpurun_out/s8/tc_raw.csv
timeout 600 ncu --set full --clock-control none -k regex:"act_bwd|bn_bwd_reduce|bn_finalize_apply|heads_|cin1|loss_|unpack|pack_batched|channel_sum" -c 24 -o gpurun_out/s8/simt python tools/one_step.py 32 192 1 > gpurun_out/s8/ncu_simt.log 2>&1
bash tools/ncu_csv.sh gpurun_out/s8/simt.ncu-rep gpurun_out/s8/simt_raw.csv
(timeout 300 python bench.py 2>&1 | tail -2) > gpurun_out/s8/bench.log
du -sh gpurun_out

set -x
mkdir -p gpurun_out/s19
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/s19/pytest.log
timeout 200 python tools/step_profile.py > gpurun_out/s19/step_profile.txt 2>&1
(timeout 300 python bench.py 2>gpurun_out/s19/bench.err | tail -2) > gpurun_out/s19/bench.log
timeout 300 ncu --set full --clock-control none -k regex:"pack_batched|heads_bwd_fused|loss_" -c 5 -o gpurun_out/s19/misc python tools/one_step.py 32 192 1 > gpurun_out/s19/ncu_misc.log 2>&1
bash tools/ncu_csv.sh gpurun_out/s19/misc.ncu-rep gpurun_out/s19/misc_raw.csv
du -sh gpurun_out

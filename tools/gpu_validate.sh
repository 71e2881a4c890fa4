# One-GPU validation + evidence run of the round: GPU tests, smoke, bench (ours + reference arm), step / layer profiles,
# ncu launch list of an eager bench step and --set full metric extracts of the tensor-core kernels.
# usage: gpurun --timeout 2400 -- bash tools/gpu_validate.sh ; then tools/ncu_traffic.py on the launch list
set -x
mkdir -p gpurun_out/validate
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/validate/pytest.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4) > gpurun_out/validate/smoke.log
(timeout 300 python bench.py 2>gpurun_out/validate/bench.err | tail -2) > gpurun_out/validate/bench.log
(timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -2) > gpurun_out/validate/bench_ref.log
FU_STREAMS=1 timeout 200 python tools/step_profile.py > gpurun_out/validate/step_profile.txt 2>&1
timeout 300 python tools/profile_layers.py 32 192 bf16 > gpurun_out/validate/layers.txt 2>&1
FU_STREAMS=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/validate/ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-profile > gpurun_out/validate/ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:"tc_conv|tc_wgrad" -c 14 -o gpurun_out/validate/tc python tools/conv_shapes.py "32 32 32 192 192 3 0" "32 64 64 96 96 3 0" "32 128 128 48 48 3 0" "32 256 256 24 24 3 0" "32 512 512 12 12 3 0" "32 1024 1024 6 6 3 0" "32 128 128 48 48 3 1" "32 32 32 192 192 3 2" "32 128 128 48 48 3 2" "32 512 512 12 12 3 2" "32 64 32 192 192 1 0" "32 64 32 96 96 2 0 -2" > gpurun_out/validate/ncu_tc.log 2>&1
bash tools/ncu_csv.sh gpurun_out/validate/tc.ncu-rep gpurun_out/validate/tc_raw.csv
du -sh gpurun_out

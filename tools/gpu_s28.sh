mkdir -p gpurun_out/s28
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/s28/pytest.log
SH='"32 32 32 192 192 2 1 2" "32 64 32 96 96 2 0 -2" "32 64 64 96 96 2 1 2" "32 128 64 48 48 2 0 -2" "32 1024 512 6 6 2 0 -2"'
eval timeout 300 python tools/conv_shapes.py --time $SH > gpurun_out/s28/time.txt 2>&1
eval FU_TC_SCATTER_MERGE=0 timeout 300 python tools/conv_shapes.py --time $SH > gpurun_out/s28/time_nomerge.txt 2>&1
timeout 200 python tools/step_profile.py > gpurun_out/s28/step_profile.txt 2>&1
(timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | tail -1) > gpurun_out/s28/bench.log

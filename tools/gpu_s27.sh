mkdir -p gpurun_out/s27
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/s27/pytest.log
SH='"32 32 32 192 192 2 0 2" "32 32 32 192 192 2 1 2" "32 64 32 96 96 2 0 -2" "32 64 32 192 192 1 0" "32 32 32 192 192 3 0" "32 64 64 96 96 3 0" "32 128 128 48 48 3 0" "32 512 512 12 12 3 0" "32 32 32 192 192 3 2" "32 512 512 12 12 3 2"'
eval timeout 300 python tools/conv_shapes.py --time $SH > gpurun_out/s27/time.txt 2>&1
timeout 200 python tools/step_profile.py > gpurun_out/s27/step_profile.txt 2>&1
(timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | tail -1) > gpurun_out/s27/bench.log

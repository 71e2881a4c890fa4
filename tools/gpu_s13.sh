set -x
mkdir -p gpurun_out/s13
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/s13/pytest.log
SH='"32 32 32 192 192 3 0" "32 32 32 192 192 3 1" "32 64 32 192 192 3 0" "32 32 64 192 192 3 1" "32 64 64 96 96 3 0" "32 128 128 48 48 3 0" "32 64 32 192 192 1 0" "32 512 512 12 12 3 0"'
eval timeout 300 python tools/conv_shapes.py --time $SH > gpurun_out/s13/kc_default.txt 2>&1
eval FU_TC_KC=16 timeout 300 python tools/conv_shapes.py --time $SH > gpurun_out/s13/kc16.txt 2>&1
eval FU_TC_KC=32 timeout 300 python tools/conv_shapes.py --time $SH > gpurun_out/s13/kc32.txt 2>&1
timeout 300 python tools/step_profile.py > gpurun_out/s13/step_profile.txt 2>&1
(timeout 300 python bench.py 2>gpurun_out/s13/bench.err | tail -2) > gpurun_out/s13/bench.log
du -sh gpurun_out

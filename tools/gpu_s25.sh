set -x
mkdir -p gpurun_out/s25
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4) > gpurun_out/s25/smoke.log
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/s25/pytest.log
timeout 200 python tools/step_profile.py > gpurun_out/s25/step_profile.txt 2>&1
(timeout 300 python bench.py 2>gpurun_out/s25/bench.err | tail -2) > gpurun_out/s25/bench.log
timeout 200 python tools/conv_shapes.py --time "32 1024 1024 6 6 3 0" "32 1024 1024 6 6 3 1" "32 512 1024 6 6 3 0" "32 1024 512 6 6 3 1" > gpurun_out/s25/shapes6.txt 2>&1

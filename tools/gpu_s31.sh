mkdir -p gpurun_out/s31
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/s31/pytest.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) > gpurun_out/s31/smoke.log
timeout 200 python tools/step_profile.py > gpurun_out/s31/sp.txt 2>&1
FU_STREAMS=1 timeout 200 python tools/step_profile.py > gpurun_out/s31/sp_1stream.txt 2>&1
(timeout 300 python bench.py --no-cpu-baseline 2>gpurun_out/s31/bench.err | tail -1) > gpurun_out/s31/bench.log
(FU_STREAMS=1 timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | tail -1) > gpurun_out/s31/bench_1stream.log
timeout 300 python tools/profile_layers.py 32 192 bf16 > gpurun_out/s31/layers.txt 2>&1

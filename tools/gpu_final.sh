# Final GPU check of the round: the whole GPU suite, smoke, one bench line.
set -x
mkdir -p gpurun_out/final
(timeout 330 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/final/pytest.log
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4) > gpurun_out/final/smoke.log
(timeout 200 python bench.py 2>gpurun_out/final/bench.err | tail -2) > gpurun_out/final/bench.log
cat gpurun_out/final/pytest.log gpurun_out/final/smoke.log gpurun_out/final/bench.log

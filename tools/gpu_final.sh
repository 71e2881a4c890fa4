set -x
mkdir -p gpurun_out/final
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/final/pytest.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5) > gpurun_out/final/smoke.log
(timeout 300 python bench.py 2>gpurun_out/final/bench.err | tail -2) > gpurun_out/final/bench.log
(timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -2) > gpurun_out/final/bench_ref.log
timeout 200 python tools/step_profile.py > gpurun_out/final/step_profile.txt 2>&1
timeout 300 python tools/profile_layers.py 32 192 bf16 > gpurun_out/final/layers.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/final/ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-profile > gpurun_out/final/ncu_bench.log 2>&1
du -sh gpurun_out

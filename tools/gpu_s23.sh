mkdir -p gpurun_out/s23
(timeout 200 python tools/smoke_probe.py 2>&1 | grep env=) > gpurun_out/s23/default.txt
(FU_TC_DISABLE=1 timeout 200 python tools/smoke_probe.py 2>&1 | grep env=) > gpurun_out/s23/simt.txt
(FU_TC_M64=0 FU_TC_FUSE_RES=0 timeout 200 python tools/smoke_probe.py 2>&1 | grep env=) > gpurun_out/s23/old_paths.txt
(timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/s23/pytest.log
(timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | tail -1) > gpurun_out/s23/bench.log

mkdir -p gpurun_out/s24
cd _old && (timeout 200 python tools/smoke_probe.py 2>&1 | grep env=) > ../gpurun_out/s24/old_commit.txt; (FU_TC_DISABLE=1 timeout 200 python tools/smoke_probe.py 2>&1 | grep env=) > ../gpurun_out/s24/old_commit_simt.txt

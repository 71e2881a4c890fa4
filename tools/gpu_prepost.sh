# GPU check of the sample-prep / post-processing kernels: parity tests, CUDA-event timings, ncu launch list (time + DRAM bytes).
set -x
mkdir -p gpurun_out/prepost
(timeout 240 python -m pytest tests/test_prepost_gpu.py -m gpu -q 2>&1 | tail -40) > gpurun_out/prepost/pytest_prepost.log
(timeout 120 python tools/prepost_time.py 256 3 2>&1 | tail -8) > gpurun_out/prepost/time_b256.log
(timeout 120 python tools/prepost_time.py 32 3 2>&1 | tail -8) > gpurun_out/prepost/time_b32.log
(timeout 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv -k regex:"prep_|heatmap_targets|ens_|extract_landmarks" -c 120 --log-file gpurun_out/prepost/ncu_prepost.csv python tools/prepost_time.py 256 3 > gpurun_out/prepost/ncu.log 2>&1)

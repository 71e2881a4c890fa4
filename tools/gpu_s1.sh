set -x
mkdir -p gpurun_out/s1
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/s1/pytest.log
(timeout 300 python bench.py 2>&1 | tail -3) > gpurun_out/s1/bench.log
for shape in "32 32 32 192 192 3" "32 64 32 192 192 3" "32 64 64 96 96 3" "32 128 128 48 48 3" "32 256 256 24 24 3" "32 512 512 12 12 3" "32 1024 1024 6 6 3"; do
  for mode in 0 1 2; do
    timeout 120 python tools/conv_time.py $shape $mode 2>&1 | grep -v Warn | tail -4
  done
done > gpurun_out/s1/conv_time.log 2>&1
for shape in "32 32 32 192 192 3" "32 128 128 48 48 3"; do
  for mode in 0 1; do
    FU_TC_DBG=1 timeout 120 python tools/conv_time.py $shape $mode 2>&1 | grep -A13 timeline | head -30
  done
done > gpurun_out/s1/timeline.log 2>&1

set -x
mkdir -p gpurun_out/s4
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/s4/pytest.log
timeout 300 python tools/profile_layers.py 32 192 bf16 > gpurun_out/s4/layers.txt 2>&1
for shape in "32 32 32 192 192 3" "32 64 32 192 192 3" "32 64 64 96 96 3"; do
  for mode in 0 1; do
  FU_TC_EPI_SETS=1 timeout 120 python tools/conv_time.py $shape $mode 2>&1 | grep TFLOP
  FU_TC_EPI_SETS=2 timeout 120 python tools/conv_time.py $shape $mode 2>&1 | grep TFLOP
  done
done > gpurun_out/s4/conv_time_sets.log 2>&1
for shape in "32 64 32 192 192 1" "32 32 64 96 96 1" "32 128 64 96 96 1" "32 256 128 48 48 1"; do
  for mode in 0 1; do
  FU_TC_EPI_GROUPS=1 timeout 120 python tools/conv_time.py $shape $mode 2>&1 | grep TFLOP
  FU_TC_EPI_GROUPS=2 timeout 120 python tools/conv_time.py $shape $mode 2>&1 | grep TFLOP
  FU_TC_EPI_GROUPS=4 timeout 120 python tools/conv_time.py $shape $mode 2>&1 | grep TFLOP
  done
done > gpurun_out/s4/conv_time_groups.log 2>&1
(timeout 300 python bench.py 2>&1 | tail -2) > gpurun_out/s4/bench.log
du -sh gpurun_out

set -x
mkdir -p gpurun_out/s3
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/s3/pytest.log
timeout 300 python tools/profile_layers.py 32 192 bf16 > gpurun_out/s3/layers.txt 2>&1
for shape in "32 32 32 192 192 3" "32 64 32 192 192 3" "32 64 64 96 96 3" "32 32 64 96 96 3"; do
  FU_TC_W3_STACK=0 timeout 120 python tools/conv_time.py $shape 2 2>&1 | grep TFLOP
  FU_TC_W3_STACK=1 timeout 120 python tools/conv_time.py $shape 2 2>&1 | grep TFLOP
done > gpurun_out/s3/conv_time_stack.log 2>&1
for shape in "32 256 256 24 24 3" "32 512 512 12 12 3" "32 1024 1024 6 6 3" "32 512 256 24 24 3"; do
  for mode in 0 1; do
    FU_TC_BN_MAX=256 timeout 120 python tools/conv_time.py $shape $mode 2>&1 | grep TFLOP
  done
done > gpurun_out/s3/conv_time_bn256.log 2>&1
(timeout 300 python bench.py 2>&1 | tail -2) > gpurun_out/s3/bench.log
du -sh gpurun_out

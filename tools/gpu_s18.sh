set -x
mkdir -p gpurun_out/s18
for m in 0 1 2; do FU_TC_M64=$m timeout 120 python tools/m64_probe.py 2>&1 | grep -E "rel err|rror" ; done > gpurun_out/s18/m64.txt 2>&1
for m in 0 1 2; do FU_TC_M64=$m timeout 200 python tools/conv_shapes.py --time "32 32 32 192 192 3 2" "32 64 32 192 192 3 2" "32 64 64 96 96 3 2" "32 128 64 96 96 3 2" "32 64 32 192 192 1 2" 2>&1 | grep TFLOP; done > gpurun_out/s18/m64_time.txt 2>&1

mkdir -p gpurun_out/s26
SH='"32 32 32 192 192 2 0 2" "32 32 32 192 192 2 1 2" "32 64 32 96 96 2 0 -2" "32 64 32 96 96 2 1 -2" "32 64 32 192 192 1 0" "32 64 32 192 192 1 1" "32 64 64 96 96 2 0 2" "32 128 64 48 48 2 0 -2"'
eval timeout 300 python tools/conv_shapes.py --time $SH > gpurun_out/s26/time.txt 2>&1
eval timeout 600 ncu --set full --clock-control none -k regex:"tc_conv_kernel" -c 8 -o gpurun_out/s26/s2 python tools/conv_shapes.py $SH > gpurun_out/s26/ncu.log 2>&1
bash tools/ncu_csv.sh gpurun_out/s26/s2.ncu-rep gpurun_out/s26/s2_raw.csv

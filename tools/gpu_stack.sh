# kw-stacked MMAs: kernel tests, per-layer times with and without, bench A/B
mkdir -p gpurun_out/stack
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "halo" 2>&1 | tail -5
SH='32 32 32 192 192 3 0|32 32 32 192 192 3 1|32 64 32 192 192 3 0|32 32 64 192 192 3 1|32 32 64 96 96 3 1|8 32 32 736 736 3 0'
for st in 0 1; do
  echo "== FU_TC_STACK=$st"
  IFS='|' read -ra A <<< "$SH"
  CONV_STATS=1 FU_TC_STACK=$st timeout 120 python tools/conv_shapes.py --time "${A[@]}" 2>&1 | grep -v "^done"
done
bash tools/gpu_ab.sh stack "FU_TC_STACK=0" "FU_TC_STACK=1" "FU_TC_STACK=0" "FU_TC_STACK=1"

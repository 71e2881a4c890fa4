mkdir -p gpurun_out/heads
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_large_goldens_gpu.py tests/test_prepost_gpu.py -x -q -m gpu 2>&1 | tail -4
for v in 0 1; do
  echo "== FU_HEADS_MMA=$v"; FU_HEADS_MMA=$v timeout 200 python tools/profile_layers.py 32 192 bf16 2>&1 | grep -i "heads\|engine kernel"
done
bash tools/gpu_ab.sh heads "FU_HEADS_MMA=0" "FU_HEADS_MMA=1" "FU_HEADS_MMA=0" "FU_HEADS_MMA=1"

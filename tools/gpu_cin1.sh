timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 200 python tools/profile_layers.py 32 192 bf16 2>&1 | grep -i "cin1\|engine kernel"
B=FLUORO_UNET_LIB=/root/repo/deepfluorolabeling-ipcai2020_b200/_variants/lib_base.so
bash tools/gpu_ab.sh cin1 $B FU_X=1 $B FU_X=1

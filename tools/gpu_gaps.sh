timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
B=FLUORO_UNET_LIB=/root/repo/deepfluorolabeling-ipcai2020_b200/_variants/lib_base.so
bash tools/gpu_ab.sh gaps $B FU_X=1 $B FU_X=1

mkdir -p gpurun_out/s30
timeout 300 python tools/profile_layers.py 32 192 bf16 > gpurun_out/s30/layers_merge.txt 2>&1
FU_TC_SCATTER_MERGE=0 timeout 300 python tools/profile_layers.py 32 192 bf16 > gpurun_out/s30/layers_nomerge.txt 2>&1
timeout 200 python tools/step_profile.py > gpurun_out/s30/sp_merge.txt 2>&1
FU_TC_SCATTER_MERGE=0 timeout 200 python tools/step_profile.py > gpurun_out/s30/sp_nomerge.txt 2>&1

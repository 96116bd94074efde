#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -m gpu -q -rf -p no:cacheprovider --timeout 300"
echo "=== unet tests"; timeout 1200 $P tests/test_unet_gpu.py tests/test_kernels_gpu.py -k "unet or initial" > gpurun_out/t_unet.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t_unet.log
echo "=== profile new"; timeout 300 python tools/profile_pass.py 512 2>&1 >/dev/null | awk '/=== pass/{p=1;next} p' | head -3
echo "=== profile old"; LDM_INITIAL_CONV_OLD=1 timeout 300 python tools/profile_pass.py 512 2>&1 >/dev/null | awk '/=== pass/{p=1;next} p' | head -3

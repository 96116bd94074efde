#!/bin/bash
for d in 0 8 16; do
echo "=== LDM_HALO_DEBUG=$d"; LDM_HALO_DEBUG=$d timeout 300 python tools/profile_pass.py 512 2>&1 >/dev/null | awk '/=== pass/{p=1;next} p' | awk '$4==0' | head -8 | awk '{printf "%s ", $5} END{print ""}'
done
echo "=== no xform"; LDM_NO_CONV_XFORM=1 timeout 300 python tools/profile_pass.py 512 2>&1 >/dev/null | awk '/=== pass/{p=1;next} p' | awk '$4==0' | head -8 | awk '{printf "%s ", $5} END{print ""}'

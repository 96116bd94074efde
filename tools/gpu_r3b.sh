#!/bin/bash
for pdl in 0 1; do for gn in 0 1; do
  echo "--- LDM_PDL=$pdl LDM_GN_NORM_SMALL_BATCH=$gn"
  if [ $gn = 1 ]; then export LDM_GN_NORM_SMALL_BATCH=1; else unset LDM_GN_NORM_SMALL_BATCH; fi
  LDM_PDL=$pdl timeout 300 python tools/batch1.py 2>&1 | tail -3
done; done

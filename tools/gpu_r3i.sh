#!/bin/bash
LDM_HALO_DEBUG=4 timeout 300 python tools/bench_conv.py 512 64 64 32 3 2>&1 | tail -16
echo; LDM_HALO_DEBUG=1 timeout 300 python tools/bench_conv.py 512 64 64 32 3 2>&1 | tail -1
LDM_HALO_DEBUG=2 timeout 300 python tools/bench_conv.py 512 64 64 32 3 2>&1 | tail -1

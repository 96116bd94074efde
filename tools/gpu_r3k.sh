#!/bin/bash
timeout 300 python tools/bench_conv.py 512 64 64 32 3 2>&1 | tail -1
LDM_EPI_DEBUG=4 timeout 300 python tools/bench_conv.py 512 64 64 32 3 2>&1 | tail -1
LDM_EPI_DEBUG=4 LDM_HALO_DEBUG=1 timeout 300 python tools/bench_conv.py 512 64 64 32 3 2>&1 | tail -1

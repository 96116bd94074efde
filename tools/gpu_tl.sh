#!/bin/bash
timeout 300 python tools/bench_conv_gn.py 512 64 64 32 3 8 1 0 0 2>&1 | tail -6
timeout 300 python tools/bench_conv_gn.py 512 64 64 32 3 8 1 1 0 2>&1 | tail -6

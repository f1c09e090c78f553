#!/bin/bash
set -u
OUT=gpurun_out/r2m
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_idct_rgb420" -s 1 -c 1 \
    -o $OUT/fused python bench.py --quick --steps 1 --warmup 1 > $OUT/ncu.log 2>&1
tail -2 $OUT/ncu.log
ls -la $OUT

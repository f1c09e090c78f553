#!/bin/bash
# one gpurun call: the ncu launch list of a short bench run and one --set full capture of every hot kernel of one step
# usage: tools/gpu_profile.sh [out dir under gpurun_out/]
set -u
OUT=gpurun_out/${1:-prof}
mkdir -p $OUT
# per step: 4 lexer kernels, k_build_luts, k_decode_par, k_zero_flagged, k_decode_fast, k_reduce_status, 3 x k_idct_tma, colour = 13
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_lex|k_build|k_decode|k_zero|k_reduce|k_idct|k_ycc" -s 13 -c 26 \
    --csv --log-file $OUT/launches.csv python bench.py --quick --steps 2 --warmup 1 > $OUT/launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_lex|k_decode_par|k_idct_tma|k_ycc420" -s 9 -c 9 \
    -o $OUT/full python bench.py --quick --steps 1 --warmup 1 > $OUT/full.log 2>&1
tail -2 $OUT/full.log
ls -la $OUT

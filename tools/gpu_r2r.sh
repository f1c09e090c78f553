#!/bin/bash
set -u
OUT=gpurun_out/r2r
mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_decode|k_zero|k_build|k_reduce|k_idct|k_ycc|k_acr" -c 80 \
    --csv --log-file $OUT/cfg4_launches.csv python tools/time_configs.py 4 > $OUT/cfg4_ncu.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$OUT/cfg4_launches.csv")) if len(r)>5 and r[0].isdigit()]
for r in rows[:60]: print(r[4][:70], r[-1])
PY
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/racecheck2.log \
    python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "fused_idct and (48 or 400 or 391) or progressive_ac_encoders and blocks and 0 or refinement_three_phase and 1-0" > $OUT/racecheck2_pytest.log 2>&1
echo "racecheck exit $?"
tail -3 $OUT/racecheck2_pytest.log
tail -8 $OUT/racecheck2.log

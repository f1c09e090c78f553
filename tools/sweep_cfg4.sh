#!/bin/bash
# K3p settings for the progressive configuration (#4): threads per interval x warm-up bits
for t in 4 5 6; do for w in 0 512; do
  JPEG_SM100_PAR_T=$t JPEG_SM100_PAR_WARM=$w python tools/time_configs.py 4 2>&1 | tail -1 > /tmp/c4.json
  python - <<PY
import json
d = json.load(open("/tmp/c4.json"))
print("T", $t, "warm", $w, d["ms"], d["decode_ms_per_scan"])
PY
done; done

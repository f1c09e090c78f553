#!/bin/bash
# e2e of the batch entry point against chunk size (RGB per device->host copy) and the number of concurrent calls
for mb in 50 100 200 400; do for st in 1 4; do
  JPEG_SM100_CHUNK_MB=$mb JPEG_BENCH_STREAMS=$st python bench.py 2>/dev/null > /tmp/e2e.json
  python - <<PY
import json
d = json.load(open("/tmp/e2e.json"))
print("chunk_MB", $mb, "streams", $st, d["e2e"]["value"], d["e2e"]["pcie_d2h_GBps_achieved"], d["e2e"]["pcie_d2h_GBps_measured"])
PY
done; done

#!/bin/bash
# e2e scaling A/B on N GPUs of one box: tools/scale_ab.sh N "<label:ENV=VAL,ENV=VAL ...>" ...   (bench lines into gpurun_out/scale/)
set -u
N=$1; shift
OUT=gpurun_out/scale
mkdir -p $OUT
nproc > $OUT/host.txt; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" >> $OUT/host.txt; nvidia-smi topo -m >> $OUT/host.txt 2>&1
for spec in "$@"; do
  label=${spec%%:*}; envs=${spec#*:}
  ( IFS=','; for kv in $envs; do [ -n "$kv" ] && export "$kv"; done
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
        bench.py --gpus $N --steps 6 --warmup 2 --no-configs > $OUT/n${N}_$label.json 2> $OUT/n${N}_$label.err )
  python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/n${N}_$label.json") if l.startswith("{")][-1])
    e = d["e2e"]
    print("$label", "N=$N value", d["value"], "e2e", e["value"], "GB/s per rank", e["pcie_d2h_GBps_achieved"], "alone", e["pcie_d2h_GBps_measured_alone_per_rank"],
          "concurrent", e["pcie_d2h_GBps_measured_concurrent"], "frac", e["frac_of_concurrent_ceiling"], d["details"]["cpu_affinity"][:2])
except Exception as ex:
    print("$label", "failed", ex)
    print(open("$OUT/n${N}_$label.err").read()[-1500:])
PY
done

#!/bin/bash
cd "$(dirname "$0")/.."
for cfg in "16 3" "32 2" "32 3" "16 4" "24 3"; do
  set -- $cfg
  python bench.py --batch $1 --inflight $2 --steps 16 --no-cpu-baseline > gpurun_out/bb_$1_$2.log 2>&1
  python - <<PY
import json
l=open("gpurun_out/bb_$1_$2.log").read().strip().split("\n")[-1]
try:
    d=json.loads(l); print("batch $1 inflight $2: resident %.0f e2e %.0f fps"%(d["value"], d["e2e"]["value"]))
except Exception as e: print("batch $1 inflight $2 failed", l[-200:])
PY
done

#!/bin/bash
# timing experiments on the tcgen05 conv (results are wrong under VTD_DBG): per-op device times of one bench step
cd "$(dirname "$0")/.."
run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --steps 6 --inflight 1 --no-cpu-baseline --profile-out gpurun_out/dbg_$name.json > gpurun_out/dbg_$name.log 2>&1
  echo "$name rc=$?"
}
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}
  run $name ${envs//,/ }
done

#!/bin/bash
# A/B of two builds on the SAME box: alternate, two rounds
cd "$(dirname "$0")/.."
for round in 1 2; do
  for v in "$@"; do
    cp profiles/ab/lib_$v.so video_text_detection_system_b200/libvtd_b200.so
    python bench.py --steps 10 --inflight 1 --no-cpu-baseline --profile-out gpurun_out/ab_${v}_$round.json > gpurun_out/ab_${v}_$round.log 2>&1
    echo "$v $round rc=$?"
  done
done

#!/bin/bash
# A/B of builds on the SAME box (box-to-box variance is ~3 %): profiles/ab_run.sh new prev
# expects video_text_detection_system_b200/libvtd_b200_<name>.so; alternates them over libvtd_b200.so, two rounds each.
cd "$(dirname "$0")/.."
L=video_text_detection_system_b200
for round in 1 2; do
  for v in "$@"; do
    cp $L/libvtd_b200_$v.so $L/libvtd_b200.so
    python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras --profile-out gpurun_out/ab_${v}_$round.json 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v $round', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],3))"
  done
done

#!/bin/bash
# Is the device code of the working tree identical to that of a git ref?  Used when a refactor must not change the
# GPU-verified library (no GPU needed): compiles csrc/*.cu of <ref> into a scratch directory and compares the SASS of
# every object with the in-tree build, ignoring symbol names (anonymous-namespace names embed a hash of the file).
#   profiles/sass_identity.sh <git-ref>
set -e
REF=${1:?usage: sass_identity.sh <git-ref>}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d)
git -C "$ROOT" archive "$REF" video_text_detection_system_b200/csrc include | tar -x -C "$TMP"
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr"
python "$ROOT/video_text_detection_system_b200/build.py" > /dev/null 2>&1 || python -m video_text_detection_system_b200.build > /dev/null 2>&1
strip_names() { cuobjdump -sass "$1" | grep -v "^Fatbin\|^=====\|host =\|compile_size\|identifier\|Function :" | md5sum | cut -c1-32; }
rc=0
for src in "$TMP"/video_text_detection_system_b200/csrc/*.cu; do
  b=$(basename "$src" .cu)
  (cd "$(dirname "$src")" && nvcc $F -c "$b.cu" -o "$TMP/$b.o" 2> /dev/null)
  if [ "$(strip_names "$TMP/$b.o")" = "$(strip_names "$ROOT/video_text_detection_system_b200/csrc/_obj/$b.o")" ]; then
    echo "same  $b"
  else
    echo "DIFF  $b"; rc=1
  fi
done
rm -rf "$TMP"
exit $rc

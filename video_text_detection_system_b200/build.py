"""Builds libvtd_b200.so (sm_100a only) in-tree with nvcc.

`python -m video_text_detection_system_b200.build` or build_library() from __graft_entry__.build().
The shared object lands next to this file so that it travels with the repository snapshot to the
GPU box; object files go to csrc/_obj/.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libvtd_b200.so")
SOURCES = ["api.cu", "preprocess.cu", "conv_generic.cu", "conv_tcgen05.cu", "db_head.cu", "boxes.cu", "crop.cu",
           "lstm.cu", "lstm_tcgen05.cu", "ctc.cu", "misc.cu", "trocr.cu", "overlay.cu"]
HEADERS = ["common.cuh", "box_geom.cuh", "tc_common.cuh", "resize_tab.h", "trocr_host.inc", "overlay_atlas.h", os.path.join("..", "..", "include", "vtd.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libvtd_b200.so cannot be built (there is no CPU fallback)")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def variant_paths(variant: str = ""):
    """(object directory, library path, extra nvcc flags) of a build variant.  "" (or "f16") = the shipped library, whose
    16-bit speed tier stores IEEE half; "bf16" = the same sources with bfloat16 as the 16-bit storage type
    (-DVTD_BF16_STORAGE, csrc/common.cuh): libvtd_b200_bf16.so, selected per Engine with dtype="bf16" or for a whole
    process with VTD_STORAGE=bf16."""
    if variant in ("", "f16", "fp16"):
        return OBJ, LIB, []
    if variant == "bf16":
        return OBJ + "_bf16", os.path.join(HERE, "libvtd_b200_bf16.so"), ["-DVTD_BF16_STORAGE"]
    if variant == "dev":      # developer build: tuning switches and per-role cycle counters compiled in (never shipped, never benched)
        return OBJ + "_dev", os.path.join(HERE, "libvtd_b200_dev.so"), ["-DVTD_DEV", "-DVTD_TIMERS"]
    if variant == "prev":     # A/B aid: a release build of an older checkout copied here as libvtd_b200_prev.so (VTD_STORAGE=prev)
        return OBJ + "_prev", os.path.join(HERE, "libvtd_b200_prev.so"), []
    raise ValueError("unknown build variant %r" % variant)


def build_library(force: bool = False, verbose: bool = False, variant: str = "") -> str:
    nvcc = _nvcc()
    OBJ, LIB, extra = variant_paths(variant)
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + hdrs):
            jobs.append((s, o))

    def run(job):
        s, o = job
        cmd = [nvcc] + NVCC_FLAGS + extra + os.environ.get("VTD_NVCC_EXTRA", "").split() + ["-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (s, r.stdout, r.stderr))
        if verbose and r.stderr.strip():
            print(r.stderr, file=sys.stderr)
        return o

    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True, variant="bf16" if "--bf16" in sys.argv else "dev" if "--dev" in sys.argv else ""))

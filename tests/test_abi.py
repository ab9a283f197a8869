"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/vtd.h declares; with no GPU
every compute entry point fails loudly (there is no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "vtd.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vtd_[a-z_0-9A-Z]+)\s*\(", src)))


def test_header_and_binding_agree(lib_built):
    from video_text_detection_system_b200 import _lib
    assert sorted(_lib.exported_symbols()) == header_functions()


def test_library_exports_every_declared_symbol(lib_built):
    lib = ctypes.CDLL(lib_built)
    for name in header_functions():
        assert hasattr(lib, name), name
    lib.vtd_abi_version.restype = ctypes.c_int
    assert lib.vtd_abi_version() == 1


def test_struct_layouts():
    from video_text_detection_system_b200 import _lib
    assert ctypes.sizeof(_lib.VtdRecord) == 128 == _lib.RECORD_DTYPE.itemsize
    assert ctypes.sizeof(_lib.VtdConfig) == 4 * 16
    for f in ("frame", "bbox", "polygon", "det_conf", "rec_conf", "len", "ids", "start_index"):
        assert getattr(_lib.VtdRecord, f).offset == _lib.RECORD_DTYPE.fields[f][1]


def test_built_for_sm100a_with_tcgen05_and_tma(lib_built):
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-sass", lib_built], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):      # tcgen05.mma, TMA tensor load, tcgen05.ld
        assert mnemonic in out, mnemonic


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_fails_loudly_without_gpu(lib_built):
    from video_text_detection_system_b200 import _lib
    with pytest.raises(_lib.VtdError) as ei:
        _lib.Engine()
    assert "no CPU fallback" in str(ei.value)
    from video_text_detection_system_b200 import DBNet, TextDetector
    with pytest.raises(_lib.VtdError):
        DBNet("resnet18", pretrained=False)(torch.zeros(1, 3, 64, 64))
    # the reference's never-raise convention at the detect() surface: [] plus a logged error
    assert TextDetector(backbone="resnet18", pretrained=False).detect(np.zeros((64, 64, 3), np.uint8)) == []


def test_bad_config_rejected(lib_built):
    from video_text_detection_system_b200 import _lib
    for kw in ({"backbone": 34}, {"det_h": 100}, {"crop_w": 130}, {"max_boxes": 5000}, {"max_batch": 0}):
        with pytest.raises(_lib.VtdError):
            _lib.Engine(**kw)


def test_storage_variants_hold_only_their_own_16bit_type(lib_built):
    """The shipped library stores IEEE half in its speed tier: not one BF16 conversion or BF16-typed MMA in its SASS.
    libvtd_b200_bf16.so (-DVTD_BF16_STORAGE, selected with dtype="bf16" / VTD_STORAGE=bf16): same exports, tcgen05/TMA
    present, BF16 conversions present."""
    import shutil
    import subprocess
    from video_text_detection_system_b200.build import build_library
    path = build_library(variant="bf16")
    lib = ctypes.CDLL(path)
    for name in header_functions():
        assert hasattr(lib, name), name
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-sass", path], capture_output=True, text=True).stdout
    assert "UTCHMMA" in out and "UTMALDG" in out and "BF16" in out
    shipped = subprocess.run([cuobjdump, "-sass", lib_built], capture_output=True, text=True).stdout
    assert "BF16" not in shipped and "F16" in shipped


def test_release_library_reads_no_tuning_environment(lib_built):
    """The VTD_DBG / VTD_NO_* / VTD_TILE ... experiment switches exist only in -DVTD_DEV builds (csrc/common.cuh dev_env):
    none of their names survives in the release binary."""
    blob = open(lib_built, "rb").read()
    for name in (b"VTD_DBG", b"VTD_NO_HALO", b"VTD_TILE", b"VTD_KPS", b"VTD_CTA2", b"VTD_NO_WIN", b"VTD_TC_STAGES"):
        assert name not in blob, name


def test_dtype_selects_the_storage_library(monkeypatch):
    """Engine(dtype="fp16") = the speed tier of the shipped library (IEEE half); "bf16" = the speed tier of
    libvtd_b200_bf16.so; "fp32" / "16bit" take the process default (VTD_STORAGE); both libraries stay loaded side by side."""
    from video_text_detection_system_b200 import _lib
    asked = []
    real = _lib.load_library

    def spy(variant=None):
        asked.append(variant)
        return real(variant)

    monkeypatch.setattr(_lib, "load_library", spy)
    monkeypatch.delenv("VTD_STORAGE", raising=False)
    for dt in ("fp32", "16bit", "fp16", "half", "bf16"):
        if torch.cuda.is_available():
            e = _lib.Engine(dtype=dt, det_h=64, det_w=64, max_src_h=64, max_src_w=64)
            assert e.dtype == {"fp32": "fp32", "16bit": "fp16", "fp16": "fp16", "half": "fp16", "bf16": "bf16"}[dt]
        else:
            with pytest.raises(_lib.VtdError):
                _lib.Engine(dtype=dt)
    assert asked == [None, None, "", "", "bf16"]
    with pytest.raises(ValueError):
        _lib.Engine(dtype="int8")
    a, b = real(""), real("bf16")
    assert a is not b and a is real(None) and a is real("f16")
    assert a._name.endswith("libvtd_b200.so") and b._name.endswith("libvtd_b200_bf16.so")
    monkeypatch.setenv("VTD_STORAGE", "bf16")
    assert real(None) is b


def test_header_is_plain_c99_and_layouts_match_the_binding(tmp_path):
    """include/vtd.h must be consumable from C (cgo / JNI / ctypes users): compiled as C99 with -pedantic, and the
    struct sizes/offsets a C compiler sees equal the ctypes and NumPy mirrors in _lib.py."""
    import shutil
    import subprocess
    from video_text_detection_system_b200 import _lib
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("gcc not available")
    src = tmp_path / "abi.c"
    src.write_text('#include "vtd.h"\n#include <stddef.h>\n#include <stdio.h>\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(vtd_record), sizeof(vtd_config),'
                   ' sizeof(vtd_tensor), offsetof(vtd_record, det_conf), offsetof(vtd_record, len),'
                   ' offsetof(vtd_record, ids), offsetof(vtd_record, start_index)); return 0;}\n')
    exe = tmp_path / "abi"
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)], text=True).split()]
    R = _lib.VtdRecord
    assert got == [ctypes.sizeof(R), ctypes.sizeof(_lib.VtdConfig), ctypes.sizeof(_lib.VtdTensor), R.det_conf.offset,
                   R.len.offset, R.ids.offset, R.start_index.offset]
    f = _lib.RECORD_DTYPE.fields
    assert got[3:] == [f["det_conf"][1], f["len"][1], f["ids"][1], f["start_index"][1]]

"""The C-ABI boundary without a GPU: the shared library builds / loads, exports every symbol ``include/dinopose.h``
declares (and nothing is bound in ``_lib.py`` that the header does not declare), the argument structs have the layout
the header gives them, and the entry points reject bad arguments with an error code + message instead of touching the
device.  No kernel is launched here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dinopose.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.findall(r"^\s*(?:const\s+char\s*\*|long\s+long|int)\s+(dp_\w+)\s*\(", src, flags=re.M)


@pytest.fixture(scope="module")
def lib():
    from dino_pose_b200 import build, _lib
    build.build()                      # incremental; cross-compiles for sm_100a without a GPU
    return _lib.lib()


def test_header_declares_what_python_binds(lib):
    from dino_pose_b200 import _lib
    declared = set(header_functions())
    assert len(declared) >= 36
    bound = set(_lib.SIGNATURES) | {"dp_last_error", "dp_abi_version", "dp_sizeof_gemm_args", "dp_sizeof_wgrad_args"}
    assert bound - declared == set(), f"bound in _lib.py but not declared in the header: {sorted(bound - declared)}"
    assert declared - bound == set(), f"declared in the header but not bound: {sorted(declared - bound)}"


def test_library_exports_every_declared_symbol(lib):
    for name in header_functions():
        assert hasattr(lib, name), f"{name} is declared in include/dinopose.h but not exported"


def test_abi_version_and_struct_layout(lib):
    from dino_pose_b200 import _lib
    assert lib.dp_abi_version() == _lib.ABI_VERSION
    assert lib.dp_sizeof_gemm_args() == C.sizeof(_lib.GemmArgs)
    assert lib.dp_sizeof_wgrad_args() == C.sizeof(_lib.WgradArgs)


def test_bad_arguments_are_rejected_on_the_host(lib):
    assert lib.dp_gemm_bf16(None, None) != 0
    assert b"null" in lib.dp_last_error()
    assert lib.dp_wgrad_bf16(None, None) != 0
    assert lib.dp_preprocess_u8(None, 1, 10, 10, 256, 224, None, None, None, None, 0, None) != 0
    assert b"dp_preprocess_u8" in lib.dp_last_error()


def test_preprocess_geometry_matches_the_oracle(lib):
    """dp_preprocess_workspace_bytes is pure host code: its plan (resized size, crop origin, needed input rows) must
    agree with the reference's Python geometry as restated by the oracle."""
    from oracle import preprocess_oracle as po
    for (h, w) in [(480, 640), (640, 427), (256, 256), (100, 150), (1080, 1920), (37, 53), (257, 511), (224, 224)]:
        need = lib.dp_preprocess_workspace_bytes(1, h, w, po.SHORT_EDGE, po.CROP)
        nh, nw = po.resized_size(h, w)
        top, _ = po.crop_origin(nh, nw)
        _, xmins, xsizes, _ = po.axis_weights(h, nh)
        rows = int(xmins[top + po.CROP - 1] + xsizes[top + po.CROP - 1] - xmins[top])
        tables = 2 * po.CROP * 160 * 2 + 2 * (2 * po.CROP * 4) + 256
        assert need >= tables + rows * po.CROP * 3 and need < tables + rows * po.CROP * 3 + 4 * 256, (h, w, need, rows)
    assert lib.dp_preprocess_workspace_bytes(1, 100, 100, 200, 224) == -1      # crop larger than the resized image
    assert lib.dp_preprocess_workspace_bytes(1, 12000, 12000, 256, 224) == -1   # 46.9x down-scaling: more than 160 taps

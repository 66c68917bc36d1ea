"""The C-ABI library builds, loads without a GPU and exports every symbol include/p2r_b200.h declares."""
import ctypes
import os.path as osp
import re

import pytest

from pose2room_b200 import _lib, build

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    return build.build()


def declared_symbols():
    text = open(osp.join(ROOT, "include", "p2r_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(p2r_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_nine_ext_ops():
    names = declared_symbols()
    for op in ["furthest_point_sampling", "gather_points", "gather_points_grad", "three_nn", "three_interpolate",
               "three_interpolate_grad", "ball_query", "group_points", "group_points_grad"]:
        assert "p2r_" + op in names


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_python_signatures_cover_header(lib_path):
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.p2r_abi_version() >= 1
    assert lib.p2r_compiled_arch() == 100


def test_bad_argument_is_reported_not_fatal(lib_path):
    lib = _lib.load()
    rc = lib.p2r_ball_query(None, None, 1, 0, 1, 0.3, 16, None, None)  # n == 0 -> argument error, no launch
    assert rc == -1
    assert b"p2r_ball_query" in lib.p2r_last_error()
    with pytest.raises(RuntimeError):
        _lib.call("p2r_nms3d", None, None, None, None, 1, 0, 0.1, 0, None, None, None)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_product_never_imports_the_oracle():
    import os
    pkg = osp.join(ROOT, "pose2room_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(osp.join(dirpath, f)).read()
                for line in src.splitlines():
                    if "oracle" in line:  # only a comment pointing at the oracle's header is tolerated
                        assert "import" not in line and "CDLL" not in line and line.lstrip().startswith(("//", "#", "*")), \
                            (osp.join(dirpath, f), line)

"""The C-ABI library loads (without a GPU) and exports every symbol include/mmemo.h declares; the
ctypes table in _lib.py covers exactly the same set.  No compute call is made."""
import os
import re
import subprocess

import pytest

from mmemo_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "mmemo.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(mmemo_[a-z0-9_]+)\s*\(", src))


@pytest.fixture(scope="module")
def built():
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib.load()


def test_header_symbols_are_exported(built):
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True,
                         text=True, check=True).stdout
    exported = set(re.findall(r" T (mmemo_[a-z0-9_]+)", out))
    missing = declared_symbols() - exported
    assert not missing, missing


def test_ctypes_table_matches_header(built):
    assert set(_lib.SIGNATURES) | {"mmemo_last_error"} == declared_symbols()
    for name in _lib.SIGNATURES:
        assert getattr(built, name).restype is not None


def test_version_and_error_string(built):
    assert built.mmemo_version() >= 100
    assert isinstance(built.mmemo_last_error(), bytes)


def test_only_sm100a_code_in_library(built):
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs

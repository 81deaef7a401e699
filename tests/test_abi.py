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
    # ... and nothing the header does not declare (no debug hooks, no process-wide setters)
    assert not (exported - declared_symbols()), exported - declared_symbols()
    assert not {"mmemo_set_workspace", "mmemo_set_sm_budget", "mmemo_set_pdl"} & exported


def test_launch_settings_are_per_stream_and_thread_safe(built):
    """re-entrant, no global state.  The settings table is keyed by the stream
    handle (opaque here: no CUDA call is made) and can be driven from many host threads."""
    import threading
    errs = []

    def worker(i):
        h = 0x1000 + 16 * i          # fake, distinct stream handles
        try:
            for _ in range(200):
                assert built.mmemo_stream_set_sm_budget(h, 2 + i) == 0
                assert built.mmemo_stream_set_pdl(h, i & 1) == 0
                assert built.mmemo_stream_set_workspace(h, None, 0) == 0
            assert built.mmemo_stream_reset(h) == 0
        except Exception as e:       # pragma: no cover
            errs.append(e)

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(8)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs
    assert built.mmemo_stream_set_sm_budget(0x1000, -1) == -1      # MMEMO_ERR_ARG
    assert built.mmemo_stream_set_workspace(0x1000, None, -5) == -1
    assert built.mmemo_stream_reset(0x1000) == 0


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

"""The C-ABI library loads and exports every symbol include/vlsat_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "vlsat_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vlsat_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from vlsat_b200 import _lib
    return _lib.load()


def test_header_symbols_exported(lib):
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vlsat_b200.h but not exported"


def test_binding_table_covers_header(lib):
    from vlsat_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()


def test_version_and_error_strings(lib):
    assert lib.vlsat_version() >= 100
    assert lib.vlsat_error_string(0) == b"ok"
    assert b"unsupported" in lib.vlsat_error_string(2).lower() or b"not supported" in lib.vlsat_error_string(2).lower()
    assert lib.vlsat_gemm_engine()


def test_argument_validation_without_a_gpu(lib):
    # null pointers / bad sizes are rejected before any launch, so these calls are safe on a CPU-only host
    assert lib.vlsat_linear_fwd(None, 4, None, 4, None, 4, 2, 2, 4, None, None, None) == 1
    assert lib.vlsat_linear_fwd(None, 4, None, 4, None, 4, 0, 2, 4, None, None, None) == 0      # empty batch is a no-op
    assert lib.vlsat_pointnet_fwd(None, 1, 3, 8, None, None, 64, None, None, 128, None, None, 768, None, None, None) == 1
    assert lib.vlsat_build_csr(None, 5, 3, None, None, None, 0, None) == 1
    assert lib.vlsat_flash_attn_fwd(None, 512, None, 512, None, 512, None, 512, None, 0, 1, 8, 64, None) == 0


def test_struct_layout_matches_header():
    from vlsat_b200._lib import Epilogue
    # 5 pointers, int64, pointer, int64, 2 floats, pointer, 2 ints, 2 pointers, int64, int (+ pad) = 120 bytes on LP64
    assert ctypes.sizeof(Epilogue) == 120 and Epilogue.split_hi.offset == 88 and Epilogue.split_fmt.offset == 112
    assert Epilogue.alpha.offset == 64 and Epilogue.scale_ptr.offset == 72 and Epilogue.act.offset == 80

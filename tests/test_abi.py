"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/qsb.h declares, with the prototypes the ctypes binding assumes (no device calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from qse_b200 import _build

    return _build.build()


def _header_functions():
    text = open(os.path.join(ROOT, "include", "qsb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qsb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_binding_one_to_one():
    from qse_b200 import lib

    assert _header_functions() == sorted(lib.SIGNATURES)


def test_library_exports_every_declared_symbol(built_lib):
    dll = ctypes.CDLL(built_lib)
    for name in _header_functions():
        assert hasattr(dll, name), f"{name} declared in include/qsb.h but not exported"
    dll.qsb_abi_version.restype = ctypes.c_int
    assert dll.qsb_abi_version() == 1
    dll.qsb_build_info.restype = ctypes.c_char_p
    assert b"sm_100a" in dll.qsb_build_info()


def test_struct_layouts_match_header():
    from qse_b200 import lib

    assert ctypes.sizeof(lib.QsbEop) == 40
    assert ctypes.sizeof(lib.QsbEops) == 8 + 8 + 8 + 5 * 8
    assert ctypes.sizeof(lib.QsbMetrics) == 4 * 8 + 8 + 4 + 4 + 3 * 8


def test_argument_errors_do_not_need_a_gpu(built_lib):
    from qse_b200 import lib

    dll = lib.load()
    ctx = ctypes.c_void_p()
    assert dll.qsb_create(0, 0, ctypes.byref(ctx)) == -1  # QSB_ERR_INVALID
    assert b"n_qubits" in dll.qsb_last_error(None)
    assert dll.qsb_create_sharded(10, 0, 0, 3, None, ctypes.byref(ctx)) == -1
    assert b"power of two" in dll.qsb_last_error(None)
    assert dll.qsb_destroy(None) == 0


def test_product_never_imports_the_oracle():
    """The product path must not route through oracle/ (no CPU fallback)."""
    pkg = os.path.join(ROOT, "qse_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "rydberg_oracle" not in src, f


def test_no_gpu_means_loud_failure():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import qse_b200 as qb

    with pytest.raises(Exception, match="no CPU fallback"):
        qb.B200(qb.lattices.chain(5.0, 3), qb.Signal([1.0], 10), qb.Signal([0.0], 10))

"""CPU: libb200q.so loads without a GPU and exports every entry point include/b200q.h declares; the ctypes
binding lists the same set.  No compute call is made here."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "b200q.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200q_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_all_declared_symbols():
    from quantizers_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        from quantizers_b200.build import build

        build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b200q.h but not exported"
    assert sorted(_lib.exported_symbols()) == names
    assert _lib.lib().b200q_version() >= 100


def test_invalid_arguments_fail_loudly_without_gpu():
    """Argument validation happens before any launch: NULL pointers / bad schemes give -EINVAL + a message."""
    from quantizers_b200 import _lib

    lib = _lib.lib()
    sc = _lib.make_scheme(__import__("torch").bfloat16, _lib.INT, 3, True, _lib.GROUP, 128)
    rc = lib.b200q_compress_int_packed(None, 1, 8, 128, ctypes.byref(sc), None, None, None, None)
    assert rc == -22
    assert b"num_bits" in lib.b200q_last_error()


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch

    from quantizers_b200 import ops

    class A:
        num_bits, type, symmetric, strategy, group_size, block_structure = 4, "int", True, "group", 128, None

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.compress_weight(torch.zeros(8, 128, dtype=torch.bfloat16), A())

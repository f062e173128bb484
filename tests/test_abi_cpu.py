"""CPU-only: the C-ABI library builds, loads, and exports every symbol the header declares;
host-side helpers answer without a GPU; creating a context without a GPU fails loudly."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as G
    G.build()
    from dfd_starter_b200 import _lib
    return _lib.load()


def test_header_symbols_are_exported(lib):
    from dfd_starter_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "dfd_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(dfd_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_pure_helpers(lib):
    from dfd_starter_b200._lib import DfdPolicyDesc
    assert lib.dfd_abi_version() == 2
    assert lib.dfd_table_replica_stride(25_000_000) % 32 == 0
    assert lib.dfd_table_replica_stride(25_000_000) >= 25_000_064
    for kind, n_in, h1, h2, a, P, B, W in [(0, 17, 64, 64, 6, 6092, 0, 12), (0, 376, 256, 256, 17, 171042, 0, 34),
                                           (1, 2, 64, 64, 9, 5197, 263, 9), (2, 0, 0, 0, 6, 678294, 611, 6),
                                           (3, 0, 0, 0, 15, 1158709, 5367, 15)]:
        d = DfdPolicyDesc(kind, n_in, h1, h2, a, 0)
        assert lib.dfd_policy_num_params(C.byref(d)) == P
        assert lib.dfd_policy_num_buffers(C.byref(d)) == B
        assert lib.dfd_policy_out_width(C.byref(d)) == W


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    from dfd_starter_b200 import _lib
    from dfd_starter_b200.device import Context
    h = C.c_void_p()
    rc = lib.dfd_ctx_create(0, C.byref(h))
    assert rc != 0 and b"no CPU fallback" in lib.dfd_last_error()
    with pytest.raises(_lib.DfdError):
        Context(0)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from dfd_starter_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.DfdError):
        _lib.load()

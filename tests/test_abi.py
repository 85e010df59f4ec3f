"""CPU: the C-ABI library loads and exports every symbol include/pvnet_b200.h declares
(no compute calls without a GPU), and argument validation returns status codes."""
import ctypes
import os
import re

from pvnet_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pvnet_b200.h")).read()
    return sorted(set(re.findall(r"PVNET_API\s+[\w\s\*]+?\b(pvnet_\w+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    L = _native.lib()
    names = _declared_symbols()
    assert len(names) >= 11
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/pvnet_b200.h but not exported"
        assert n in _native.SIGNATURES, f"{n} has no ctypes signature in pvnet_b200/_native.py"
    for n in _native.SIGNATURES:
        assert n in names, f"{n} bound but not declared in the header"


def test_status_codes_and_error_text():
    L = _native.lib()
    n = ctypes.c_size_t()
    assert L.pvnet_vote_workspace_bytes(16, 480, 640, 9, 256, ctypes.byref(n)) == 0 and n.value > 16 * 480 * 640 * 4
    assert L.pvnet_vote_workspace_bytes(0, 480, 640, 9, 256, ctypes.byref(n)) == -1
    assert b"dimension" in L.pvnet_last_error()
    assert L.pvnet_generate_hypothesis(None, None, None, None, 1, 1, 1, None) == -1
    assert b"null" in L.pvnet_last_error()
    assert L.pvnet_version() >= 1


def test_host_layer_refuses_cpu_tensors():
    import pytest
    import torch
    from pvnet_b200 import ransac_voting_gpu as rv
    with pytest.raises(RuntimeError, match="CUDA"):
        rv.ransac_voting_layer_v3(torch.zeros(1, 8, 8, dtype=torch.int64), torch.zeros(1, 8, 8, 1, 2), 8)

"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every
symbol include/ssdsb200.h declares; host-side logic raises like the reference."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ssdsb200.h")).read()
    return sorted(set(re.findall(r"SSDSB_API[^;(]*?\b(ssdsb_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from ssds_pytorch_b200 import build
    path = build.build()
    lib = ctypes.CDLL(path)
    names = declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ssdsb200.h but not exported"
    lib.ssdsb_version.restype = ctypes.c_int
    assert lib.ssdsb_version() >= 100


def test_python_binding_covers_header():
    from ssds_pytorch_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_argument_errors_without_gpu():
    """Validation happens before any CUDA call, so it is testable on CPU."""
    from ssds_pytorch_b200 import _lib
    lib = _lib.lib
    rc = lib.ssdsb_nms(None, None, None, 1, 10, 0.5, 0, 1, None, None, None, None, None, None, 0, None)
    assert rc == _lib.ERR_INVALID
    assert b"ndetections" in lib.ssdsb_last_error_string()
    with pytest.raises(ValueError):
        _lib.check(rc, "nms")
    lv = (_lib.Level * 1)(_lib.Level(None, None, None, 3, 5, 4, 4, 8))
    assert lib.ssdsb_decode_workspace_bytes(lv, 1, 2, 300) > 0
    rc = lib.ssdsb_decode(lv, 1, 2, 0.01, 100000, 1, None, None, None, None, None, 0, None)
    assert rc == _lib.ERR_UNSUPPORTED
    rc = lib.ssdsb_decode(lv, 0, 2, 0.01, 300, 1, None, None, None, None, None, 0, None)
    assert rc == _lib.ERR_INVALID


def test_configure_ratio_scale_matches_reference_behaviour():
    from ssds_pytorch_b200 import configure_ratio_scale
    r, s = configure_ratio_scale(2, [1, 2, 0.5], [[2.0, 2.828], 4.0])
    assert r == [[1, 2, 0.5], [1, 2, 0.5]] and s == [[2.0, 2.828], [4.0]]
    with pytest.raises(ValueError):
        configure_ratio_scale(3, [1, 2], [2.0, 4.0])
    with pytest.raises(ValueError):
        configure_ratio_scale(2, [[1], [2], [3]], [2.0, 4.0])


def test_extract_targets_bad_match_raises_like_reference():
    import torch
    from ssds_pytorch_b200 import extract_targets
    with pytest.raises(ValueError):
        extract_targets(torch.zeros(1, 1, 5), {8: torch.zeros(1, 4)}, 3, 8, (4, 4), match=["x", 1])


def test_workspace_is_grow_only_and_keyed():
    import torch
    from ssds_pytorch_b200 import _lib
    dev = torch.device("cpu")                     # exercises the keying / growth logic without a GPU
    a = _lib.workspace(100, dev)
    assert a.numel() >= 1 << 20 and a.dtype == torch.uint8
    assert _lib.workspace(4096, dev) is a         # reused while it is large enough
    b = _lib.workspace(a.numel() + 1, dev)
    assert b is not a and b.numel() >= a.numel() + 1 and _lib.workspace(10, dev) is b

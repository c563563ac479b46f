"""CPU suite, part 2: the C-ABI shared library loads and exports exactly what
include/nm_b200.h declares (no compute calls: there is no GPU here)."""
import ctypes as C
import os
import re

import pytest

import niftymatch_b200._lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "nm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nm_[a-z0-9_]+)\s*\(", src)))


def test_library_exists_and_loads():
    assert os.path.exists(L.LIB_PATH), "run `make` (or __graft_entry__.build()) first"
    L.load()


def test_every_declared_symbol_is_exported_and_bound():
    lib = L.load()
    names = _header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in nm_b200.h but not exported"
        assert n in L.SIGNATURES, f"{n} declared in nm_b200.h but not bound in _lib.SIGNATURES"
    for n in L.SIGNATURES:
        assert n in names, f"{n} bound in Python but not declared in nm_b200.h"


def test_multi_gpu_library_exports_its_header():
    """include/nm_b200_mgpu.h <-> niftymatch_b200/mgpu.py SIGNATURES <-> libnm_b200_mgpu.so (loads without a GPU)."""
    import niftymatch_b200.mgpu as M
    src = open(os.path.join(ROOT, "include", "nm_b200_mgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = sorted(set(re.findall(r"\b(nm_mgpu_[a-z0-9_]+)\s*\(", src)))
    assert len(names) >= 10
    lib = M.load()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in nm_b200_mgpu.h but not exported"
        assert n in M.SIGNATURES, f"{n} not bound in mgpu.SIGNATURES"
    for n in M.SIGNATURES:
        assert n in names, f"{n} bound in Python but not declared in nm_b200_mgpu.h"
    assert lib.nm_mgpu_world(None) == -1 and lib.nm_mgpu_destroy(None) == 0


def test_struct_layout_matches_header():
    # nm_sift_params: 6 ints, 5 floats, 8 floats, 1 int, 2 floats = 22 * 4 bytes
    assert C.sizeof(L.SiftParamsC) == 22 * 4


def test_host_only_entry_points():
    lib = L.load()
    assert lib.nm_version().startswith(b"nm-b200")
    assert lib.nm_strerror(0) == b"ok"
    assert lib.nm_strerror(-1) == b"invalid argument"
    p = L.SiftParamsC()
    assert lib.nm_sift_params_init(C.byref(p), 1920, 1080) == 0
    assert p.num_octaves == 6 and p.num_sigmas == 5 and p.num_dog_levels == 3
    assert lib.nm_sift_params_init(C.byref(p), 0, 10) == -1
    assert lib.nm_sift_params_init(None, 10, 10) == -1


def test_no_cpu_fallback_without_device():
    """Without a CUDA device every compute entry point must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = L.load()
    assert lib.nm_device_cc() == -5          # NM_ERR_NO_DEVICE
    p = L.SiftParamsC()
    lib.nm_sift_params_init(C.byref(p), 64, 64)
    ctx = C.c_void_p()
    assert lib.nm_sift_create(C.byref(ctx), C.byref(p), 1, 16) != 0
    import niftymatch_b200 as nm
    with pytest.raises(nm.NmError):
        nm.SiftBatch(nm.SiftParams(64, 64), 1, 16)


def test_product_does_not_reference_oracle():
    """Nothing under niftymatch_b200/ or include/ may import, link or execute oracle/."""
    for base in ("niftymatch_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    assert "libnm_oracle" not in txt and "nm_oracle" not in txt and "libnmref" not in txt, (dp, f)

"""ctypes binding of the multi-GPU C-ABI (include/nm_b200_mgpu.h, niftymatch_b200/libnm_b200_mgpu.so): database-sharded
matching with one NCCL all-gather, frame-sharded batched SIFT.  Test / bench plumbing like the rest of the package."""
from __future__ import annotations

import ctypes as C
import os

from . import _lib
from ._lib import SiftParamsC, check

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libnm_b200_mgpu.so")
ID_BYTES = 128
_vp, _i, _f = C.c_void_p, C.c_int, C.c_float

# name -> (restype, argtypes): every symbol include/nm_b200_mgpu.h declares (tests/test_abi.py checks the three agree)
SIGNATURES = {
    "nm_mgpu_unique_id": (_i, [_vp]),
    "nm_mgpu_create": (_i, [C.POINTER(_vp), _i, _vp, _vp]),
    "nm_mgpu_create_rank": (_i, [C.POINTER(_vp), _i, _i, _vp, _vp]),
    "nm_mgpu_destroy": (_i, [_vp]),
    "nm_mgpu_world": (_i, [_vp]),
    "nm_mgpu_local": (_i, [_vp]),
    "nm_mgpu_match_f32": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _f, _vp, _vp]),
    "nm_mgpu_set_query_groups": (_i, [_vp, _i]),
    "nm_mgpu_set_trace": (_i, [_vp, _i]),
    "nm_mgpu_match_phase_ms": (_i, [_vp, _vp]),
    "nm_mgpu_sift_create": (_i, [_vp, C.POINTER(SiftParamsC), _i, _i]),
    "nm_mgpu_sift_run_host": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp]),
}

_mlib = None


def load() -> C.CDLL:
    """Load libnm_b200_mgpu.so (after the core library, which it links).  Fails loudly."""
    global _mlib
    if _mlib is None:
        _lib.load()
        # One NCCL per process: libnm_b200_mgpu.so links libnccl.so.2 (the system one for C++ consumers).  In a Python
        # process PyTorch brings its own, newer libnccl.so.2 with the same SONAME; whichever is loaded first serves
        # both, and torch cannot start on the older one -- so the wheel's copy is loaded first when there is one.
        try:
            import glob
            import nvidia.nccl as _nccl_pkg
            for cand in sorted(glob.glob(os.path.join(list(_nccl_pkg.__path__)[0], "lib", "libnccl.so*"))):
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
                break
        except Exception:
            pass
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it first (`make`).  There is no fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _mlib = lib
    return _mlib


def unique_id() -> bytes:
    buf = C.create_string_buffer(ID_BYTES)
    check(load().nm_mgpu_unique_id(buf), "nm_mgpu_unique_id")
    return buf.raw


def _ptr_array(values):
    arr = (C.c_void_p * len(values))(*[C.c_void_p(v) for v in values])
    return arr


class MultiGpu:
    """nm_mgpu context.  MultiGpu(n_dev=2) = one process driving devices 0..n_dev-1; MultiGpu(rank=r, world=w, uid=...)
    = one process per GPU on the current device (uid from rank 0's unique_id(), distributed by the caller)."""

    def __init__(self, n_dev: int | None = None, devices=None, rank: int | None = None, world: int | None = None, uid: bytes | None = None):
        self.lib = load()
        self._ctx = C.c_void_p()
        if n_dev is not None:
            devs = (C.c_int * n_dev)(*(devices if devices is not None else range(n_dev)))
            check(self.lib.nm_mgpu_create(C.byref(self._ctx), n_dev, devs, None), "nm_mgpu_create")
            self.devices = list(devs)
        else:
            import torch
            buf = C.create_string_buffer(uid, ID_BYTES)
            check(self.lib.nm_mgpu_create_rank(C.byref(self._ctx), rank, world, buf, None), "nm_mgpu_create_rank")
            self.devices = [torch.cuda.current_device()]
        self.world = self.lib.nm_mgpu_world(self._ctx)
        self.n_local = self.lib.nm_mgpu_local(self._ctx)

    def close(self):
        if self._ctx:
            self.lib.nm_mgpu_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def match(self, A, B, offsets, ambiguity: float = 0.8, match_io=None, streams=None):
        """A, B, match_io: lists with one cuda tensor per local device (A[d] (nA,128) float32 on device d, B[d] that
        device's shard, offsets[d] its first global row).  Returns the list of merged match index tensors.  streams:
        optional list of torch streams (then nothing synchronises)."""
        import torch
        nA = A[0].shape[0]
        for d in range(self.n_local):
            assert A[d].is_cuda and A[d].dtype == torch.float32 and A[d].is_contiguous() and tuple(A[d].shape) == (nA, 128)
            assert B[d].shape[0] == 0 or (B[d].is_cuda and B[d].dtype == torch.float32 and B[d].is_contiguous() and B[d].shape[1] == 128)
        if match_io is None:
            match_io = [torch.full((nA,), -1, dtype=torch.int32, device=A[d].device) for d in range(self.n_local)]
        nB = (C.c_int * self.n_local)(*[int(b.shape[0]) for b in B])
        off = (C.c_int * self.n_local)(*[int(o) for o in offsets])
        st = _ptr_array([s.cuda_stream for s in streams]) if streams is not None else None
        check(self.lib.nm_mgpu_match_f32(self._ctx, _ptr_array([a.data_ptr() for a in A]), nA,
                                         _ptr_array([b.data_ptr() if b.shape[0] else 0 for b in B]), nB, off, ambiguity,
                                         _ptr_array([m.data_ptr() for m in match_io]), st), "nm_mgpu_match_f32")
        return match_io

    def set_query_groups(self, q_groups: int):
        """world = q_groups x D: rank r scans query block r // D against database shard r % D."""
        check(self.lib.nm_mgpu_set_query_groups(self._ctx, q_groups), "nm_mgpu_set_query_groups")
        self.q_groups = q_groups

    def set_trace(self, on: bool = True):
        check(self.lib.nm_mgpu_set_trace(self._ctx, int(on)), "nm_mgpu_set_trace")

    def match_phase_ms(self):
        buf = (C.c_float * 4)()
        check(self.lib.nm_mgpu_match_phase_ms(self._ctx, buf), "nm_mgpu_match_phase_ms")
        return dict(zip(["shard_scan", "all_gather", "merge", "total"], list(buf)))

    def sift_create(self, params, max_frames: int, capacity: int):
        self.capacity = capacity
        check(self.lib.nm_mgpu_sift_create(self._ctx, C.byref(params.c), max_frames, capacity), "nm_mgpu_sift_create")

    def sift_run_host(self, frames_pinned, out=None):
        """frames_pinned: pinned torch tensor (n, h, w) float32.  Returns dict of pinned outputs in frame order."""
        import torch
        n, cap = frames_pinned.shape[0], self.capacity
        if out is None:
            out = {"counts": torch.zeros(n, dtype=torch.int32).pin_memory(),
                   "desc": torch.zeros((n, cap, 128), dtype=torch.float32).pin_memory(),
                   "x": torch.zeros((n, cap), dtype=torch.float32).pin_memory(),
                   "y": torch.zeros((n, cap), dtype=torch.float32).pin_memory()}
        check(self.lib.nm_mgpu_sift_run_host(self._ctx, C.c_void_p(frames_pinned.data_ptr()), n, C.c_void_p(out["counts"].data_ptr()),
                                             C.c_void_p(out["desc"].data_ptr()), C.c_void_p(out["x"].data_ptr()),
                                             C.c_void_p(out["y"].data_ptr())), "nm_mgpu_sift_run_host")
        return out

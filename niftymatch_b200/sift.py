"""Batched SIFT detect+describe through the C-ABI (nm_sift_*), torch tensors as device memory."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import SiftParamsC, check


class SiftParams:
    """Mirror of the reference's SiftParams (src/gpu/sift/siftparams.h:14-99): same member
    names (with the leading underscore), same derivation, all members mutable."""

    _MAP = {
        "_width": "width", "_height": "height", "_num_octaves": "num_octaves",
        "_num_dog_levels": "num_dog_levels", "_level_max": "level_max", "_level_min": "level_min",
        "_sigma_d_0": "sigma_d_0", "_sigma_k": "sigma_k", "_sigma_0": "sigma_0", "_sigma_n": "sigma_n",
        "_base_smooth": "base_smooth", "_peak_threshold": "peak_threshold",
        "_edge_threshold": "edge_threshold",
    }

    def __init__(self, width: int, height: int):
        object.__setattr__(self, "c", SiftParamsC())
        check(_lib.load().nm_sift_params_init(C.byref(self.c), width, height), "nm_sift_params_init")

    def __getattr__(self, name):
        if name in SiftParams._MAP:
            return getattr(self.c, SiftParams._MAP[name])
        if name == "_sigmas":
            return [self.c.sigmas[i] for i in range(self.c.num_sigmas)]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in SiftParams._MAP:
            setattr(self.c, SiftParams._MAP[name], value)
        else:
            raise AttributeError(name)


def gaussian_taps(sigma: float):
    """Host taps + radius as PyramidData::create_kernel_for_sigma builds them."""
    buf = (C.c_float * 96)()
    r = C.c_int()
    check(_lib.load().nm_gaussian_taps(sigma, buf, C.byref(r)), "nm_gaussian_taps")
    return np.array(buf[: 2 * r.value + 1], dtype=np.float32), r.value


class _DevView:
    """Zero-copy view of a device buffer owned by the C side (cuda array interface v2)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {
            "shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2,
            "strides": None,
        }


def _as_tensor(ptr, shape, typestr="<f4"):
    import torch
    if ptr is None or int(np.prod(shape)) == 0:
        return torch.empty(tuple(shape), dtype=torch.float32 if typestr == "<f4" else torch.int32, device="cuda")
    return torch.as_tensor(_DevView(ptr, shape, typestr), device="cuda")


def _stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class SiftBatch:
    """Workspace + driver for a batch of frames (nm_sift_create / run / results)."""

    def __init__(self, params: SiftParams, max_batch: int, capacity: int = 2048):
        self.lib = _lib.load()
        self.params = params
        self.max_batch = max_batch
        self.capacity = capacity
        self.n_oct = params._num_octaves
        self._ctx = C.c_void_p()
        check(self.lib.nm_sift_create(C.byref(self._ctx), C.byref(params.c), max_batch, capacity), "nm_sift_create")

    def close(self):
        if self._ctx:
            self.lib.nm_sift_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_exact_descriptor(self, exact: bool):
        check(self.lib.nm_sift_set_exact_descriptor(self._ctx, int(exact)), "nm_sift_set_exact_descriptor")

    def set_dense_gradients(self, dense: bool):
        """True: gradient maps of every pixel (needed before reading grad()); False (default): only the
        blocks that keypoint windows read."""
        check(self.lib.nm_sift_set_dense_gradients(self._ctx, int(dense)), "nm_sift_set_dense_gradients")

    def run(self, frames) -> None:
        """frames: cuda float32 tensor (n, height, width), contiguous.  Asynchronous."""
        assert frames.is_cuda and frames.is_contiguous() and frames.dtype.is_floating_point and frames.element_size() == 4
        n = frames.shape[0]
        assert tuple(frames.shape[1:]) == (self.params._height, self.params._width)
        check(self.lib.nm_sift_run(self._ctx, C.c_void_p(frames.data_ptr()), n, _stream_ptr()), "nm_sift_run")
        self._n = n

    def run_bgra(self, frames) -> None:
        """frames: cuda uint8 tensor (n, height, width, 4), BGRA.  Grey conversion fused into the base blur."""
        assert frames.is_cuda and frames.is_contiguous() and frames.element_size() == 1 and frames.shape[-1] == 4
        n = frames.shape[0]
        assert tuple(frames.shape[1:3]) == (self.params._height, self.params._width)
        check(self.lib.nm_sift_run_bgra(self._ctx, C.c_void_p(frames.data_ptr()), n, _stream_ptr()), "nm_sift_run_bgra")
        self._n = n

    def run_host(self, frames_host, want_desc=True, out=None):
        """End to end from HOST memory (numpy array or pinned torch tensor): H2D copy, run,
        D2H of counts / descriptors / coordinates.  Returns dict of numpy views."""
        import torch
        if isinstance(frames_host, torch.Tensor):
            n, ptr = frames_host.shape[0], frames_host.data_ptr()
        else:
            frames_host = np.ascontiguousarray(frames_host, dtype=np.float32)
            n, ptr = frames_host.shape[0], frames_host.ctypes.data
        cap = self.capacity
        if out is None:
            out = {
                "counts": torch.zeros(self.max_batch, dtype=torch.int32).pin_memory(),
                "desc": torch.zeros((self.max_batch, cap, 128), dtype=torch.float32).pin_memory() if want_desc else None,
                "x": torch.zeros((self.max_batch, cap), dtype=torch.float32).pin_memory(),
                "y": torch.zeros((self.max_batch, cap), dtype=torch.float32).pin_memory(),
            }
        dp = C.c_void_p(out["desc"].data_ptr()) if out.get("desc") is not None else None
        check(self.lib.nm_sift_run_host(self._ctx, C.c_void_p(ptr), n, C.c_void_p(out["counts"].data_ptr()), dp,
                                        C.c_void_p(out["x"].data_ptr()), C.c_void_p(out["y"].data_ptr()),
                                        _stream_ptr()), "nm_sift_run_host")
        self._n = n
        return out

    def results(self):
        """Device-side results of the last run as torch views (valid until the next run)."""
        ptrs = [C.c_void_p() for _ in range(7)]
        check(self.lib.nm_sift_results(self._ctx, *[C.byref(p) for p in ptrs]), "nm_sift_results")
        B, cap, S = self.max_batch, self.capacity, self.n_oct * 3
        return {
            "desc": _as_tensor(ptrs[0].value, (B, cap, 128)),
            "x": _as_tensor(ptrs[1].value, (B, cap)),
            "y": _as_tensor(ptrs[2].value, (B, cap)),
            "counts": _as_tensor(ptrs[3].value, (B,), "<i4"),
            "kpts": _as_tensor(ptrs[4].value, (B, cap, 4)),
            "orient": _as_tensor(ptrs[5].value, (B, cap, 2)),
            "seg_counts": _as_tensor(ptrs[6].value, (B, S), "<i4"),
        }

    def level(self, frame: int, octave: int, level: int):
        """Gaussian level as a (h, w) torch view (pitch handled by slicing)."""
        p, pitch, w, h = C.c_void_p(), C.c_int(), C.c_int(), C.c_int()
        check(self.lib.nm_sift_level(self._ctx, frame, octave, level, C.byref(p), C.byref(pitch), C.byref(w), C.byref(h)), "nm_sift_level")
        return _as_tensor(p.value, (h.value, pitch.value))[:, : w.value]

    def grad(self, frame: int, octave: int, level: int):
        p, pitch, w, h = C.c_void_p(), C.c_int(), C.c_int(), C.c_int()
        check(self.lib.nm_sift_grad(self._ctx, frame, octave, level, C.byref(p), C.byref(pitch), C.byref(w), C.byref(h)), "nm_sift_grad")
        return _as_tensor(p.value, (h.value, pitch.value, 2))[:, : w.value]

    def last_launches(self) -> int:
        return self.lib.nm_sift_last_launches(self._ctx)

    def set_mask(self, mask=None):
        """Detector mask (compute_keypoints_with_mask, siftfunctions.cu:65-98): an h x w float image (numpy
        array or torch tensor, host or device); None removes it.  Pixels whose mask sample is < 1 give no keypoint."""
        if mask is None:
            check(self.lib.nm_sift_set_mask(self._ctx, 0), "nm_sift_set_mask")
            return
        import torch
        if isinstance(mask, torch.Tensor):
            mask = mask.contiguous().float()
            ptr, (h, w) = mask.data_ptr(), mask.shape
        else:
            mask = np.ascontiguousarray(mask, dtype=np.float32)
            ptr, (h, w) = mask.ctypes.data, mask.shape
        assert (h, w) == (self.params._height, self.params._width)
        check(self.lib.nm_sift_set_mask_image(self._ctx, C.c_void_p(ptr), int(w), int(h)), "nm_sift_set_mask_image")

    def enable_timing(self, on: bool = True):
        check(self.lib.nm_sift_enable_timing(self._ctx, int(on)), "nm_sift_enable_timing")

    def stage_ms(self):
        """Stage times of the last timed run: extrema (DoG + 26-neighbour test + refinement) and gradient (the
        gradient maps of the blocks keypoint windows read) are separate kernels; extrema_grad is their sum."""
        buf = (C.c_float * 7)()
        check(self.lib.nm_sift_stage_ms7(self._ctx, buf), "nm_sift_stage_ms7")
        d = dict(zip(["pyramid", "extrema", "compaction", "gradient", "orientation", "descriptor", "total"], list(buf)))
        d["extrema_grad"] = d["extrema"] + d["gradient"]
        return d


# ---- per-stage operators on torch tensors (used by the parity tests) --------------------
def blur(image, taps_dev, radius: int, buffer=None):
    import torch
    h, w = image.shape
    out = torch.empty_like(image)
    bp = C.c_void_p(buffer.data_ptr()) if buffer is not None else None
    check(_lib.load().nm_blur_f32(C.c_void_p(out.data_ptr()), C.c_void_p(image.data_ptr()), bp, w, h,
                                  C.c_void_p(taps_dev.data_ptr()), radius, _stream_ptr()), "nm_blur_f32")
    return out


def downsample2(image):
    import torch
    h, w = image.shape
    out = torch.empty((h // 2, w // 2), dtype=image.dtype, device=image.device)
    check(_lib.load().nm_downsample2_f32(C.c_void_p(out.data_ptr()), w // 2, h // 2, C.c_void_p(image.data_ptr()), w, h,
                                         _stream_ptr()), "nm_downsample2_f32")
    return out


def subtract(a, b):
    import torch
    h, w = a.shape
    out = torch.empty_like(a)
    check(_lib.load().nm_subtract_f32(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(out.data_ptr()), w, h,
                                      _stream_ptr()), "nm_subtract_f32")
    return out


def gradient(image):
    import torch
    h, w = image.shape
    out = torch.zeros((h, w, 2), dtype=image.dtype, device=image.device)
    check(_lib.load().nm_gradient_f32(C.c_void_p(image.data_ptr()), C.c_void_p(out.data_ptr()), w, h, _stream_ptr()), "nm_gradient_f32")
    return out


def keypoints_dense(cur, down, up, peak, edge, xper, sigma_0, num_dogs, level):
    import torch
    h, w = cur.shape
    out = torch.full((h, w, 4), -1.0, dtype=torch.float32, device=cur.device)
    check(_lib.load().nm_keypoints_dense_f32(C.c_void_p(cur.data_ptr()), C.c_void_p(down.data_ptr()), C.c_void_p(up.data_ptr()),
                                             w, h, peak, edge, xper, sigma_0, num_dogs, level, C.c_void_p(out.data_ptr()),
                                             _stream_ptr()), "nm_keypoints_dense_f32")
    return out


def collate(dense):
    import torch
    n = dense.shape[0] * dense.shape[1]
    out = torch.full((n, 4), -1.0, dtype=torch.float32, device=dense.device)
    cnt = torch.zeros(1, dtype=torch.int32, device=dense.device)
    check(_lib.load().nm_collate_f32(C.c_void_p(dense.data_ptr()), n, C.c_void_p(out.data_ptr()), C.c_void_p(cnt.data_ptr()),
                                     _stream_ptr()), "nm_collate_f32")
    return out[: int(cnt.item())]


def orientations(kpts, grad, ow, oh, xper):
    import torch
    n = kpts.shape[0]
    out = torch.full((n, 2), -1.0, dtype=torch.float32, device=kpts.device)
    if n:
        check(_lib.load().nm_orientations_f32(C.c_void_p(kpts.data_ptr()), C.c_void_p(grad.data_ptr()), n, ow, oh, 1.5, xper,
                                              C.c_void_p(out.data_ptr()), _stream_ptr()), "nm_orientations_f32")
    return out


def descriptors(kpts, orient, grad, ow, oh, num_dogs, xper):
    import torch
    n = kpts.shape[0]
    desc = torch.zeros((n, 128), dtype=torch.float32, device=kpts.device)
    x = torch.zeros(n, dtype=torch.float32, device=kpts.device)
    y = torch.zeros(n, dtype=torch.float32, device=kpts.device)
    if n:
        check(_lib.load().nm_descriptors_f32(C.c_void_p(kpts.data_ptr()), C.c_void_p(orient.data_ptr()), C.c_void_p(grad.data_ptr()),
                                             n, ow, oh, num_dogs, xper, C.c_void_p(desc.data_ptr()), C.c_void_p(x.data_ptr()),
                                             C.c_void_p(y.data_ptr()), _stream_ptr()), "nm_descriptors_f32")
    return desc, x, y


# ---- input preprocessing -----------------------------------------------------
def grayscale(bgra):
    """cuda_grayscale<float>: bgra (h, w, 4) uint8 cuda -> (h, w) float32."""
    import torch
    h, w = bgra.shape[:2]
    out = torch.empty((h, w), dtype=torch.float32, device=bgra.device)
    check(_lib.load().nm_grayscale_bgra_f32(C.c_void_p(bgra.data_ptr()), C.c_void_p(out.data_ptr()), w, h, _stream_ptr()),
          "nm_grayscale_bgra_f32")
    return out


def cast_u8(src, max_val: int = 0):
    """cuda_cast<float, unsigned char>."""
    import torch
    rows, cols = src.shape
    out = torch.empty((rows, cols), dtype=torch.uint8, device=src.device)
    check(_lib.load().nm_cast_f32_u8(C.c_void_p(src.data_ptr()), cols, rows, C.c_void_p(out.data_ptr()), max_val, _stream_ptr()),
          "nm_cast_f32_u8")
    return out


def undistort_map(x, y, camera_matrix, distortion_coeffs):
    """cuda_undistort: (u, v) for coordinate maps x, y (rows, cols); camera_matrix (fx, fy, cx, cy) and
    distortion_coeffs (k1, k2, k3) are cuda float32 tensors."""
    import torch
    rows, cols = x.shape
    u, v = torch.empty_like(x), torch.empty_like(y)
    check(_lib.load().nm_undistort_map_f32(C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), cols, rows,
                                           C.c_void_p(camera_matrix.data_ptr()), C.c_void_p(distortion_coeffs.data_ptr()),
                                           C.c_void_p(u.data_ptr()), C.c_void_p(v.data_ptr()), _stream_ptr()), "nm_undistort_map_f32")
    return u, v

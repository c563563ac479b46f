"""Registration after matching through the C-ABI (nm_align_points_f32, nm_ransac_*): the
reference's align_points / ransac_translation / ransac_similarity / ransac_homography
(src/gpu/kernels/ransac.h:8-22)."""
from __future__ import annotations

import ctypes as C

from . import _lib
from ._lib import check
from .sift import _stream_ptr

TRANSLATION, SIMILARITY, HOMOGRAPHY = 0, 1, 2


def _p(t):
    return C.c_void_p(t.data_ptr())


def _f32(*ts):
    """The C-ABI reads raw memory: every coordinate array must be a contiguous cuda float32 tensor of one shape."""
    import torch
    shape = ts[0].shape
    for t in ts:
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.shape == shape):
            raise TypeError("expected contiguous cuda float32 tensors of equal shape")


def _int32(t, name):
    import torch
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.int32 and t.is_contiguous()):
        raise TypeError(f"{name}: expected a contiguous cuda int32 tensor (torch.nonzero / argmax give int64)")


def align_points(src_x, src_y, dst_x, dst_y, matches):
    """Correspondences (c_src_x, c_src_y, c_dst_x, c_dst_y): src[i] with dst[matches[i]], -1 where unmatched."""
    import torch
    _f32(src_x, src_y); _f32(dst_x, dst_y); _int32(matches, "matches")
    n = matches.shape[0]
    assert src_x.shape[0] == n
    out = [torch.empty(n, dtype=torch.float32, device=src_x.device) for _ in range(4)]
    check(_lib.load().nm_align_points_f32(_p(src_x), _p(src_y), _p(dst_x), _p(dst_y), *[_p(o) for o in out],
                                          _p(matches), n, _stream_ptr()), "nm_align_points_f32")
    return tuple(out)


def ransac_hypotheses(kind, src_x, src_y, dst_x, dst_y, rand_list, inlier_threshold):
    """All hypotheses of a caller-supplied index list: (iterations, 9) homographies and inlier counts."""
    import torch
    _f32(src_x, src_y, dst_x, dst_y); _int32(rand_list, "rand_list")
    m = (1, 2, 4)[kind]
    iterations = rand_list.numel() // m
    H = torch.empty((iterations, 9), dtype=torch.float32, device=src_x.device)
    inl = torch.empty(iterations, dtype=torch.int32, device=src_x.device)
    check(_lib.load().nm_ransac_hypotheses_f32(kind, _p(src_x), _p(src_y), _p(dst_x), _p(dst_y), src_x.shape[0],
                                               _p(rand_list), iterations, inlier_threshold, _p(H), _p(inl),
                                               _stream_ptr()), "nm_ransac_hypotheses_f32")
    return H, inl


def ransac(kind, src_x, src_y, dst_x, dst_y, inlier_threshold, iterations, seed=0, homography=None):
    """Best model of `iterations` random hypotheses; no host synchronisation.  Returns (homography (9,) cuda,
    status (3,) cuda int32 = [ok, inliers, iteration])."""
    import torch
    _f32(src_x, src_y, dst_x, dst_y)
    if homography is None:
        homography = torch.zeros(9, dtype=torch.float32, device=src_x.device)
    status = torch.zeros(3, dtype=torch.int32, device=src_x.device)
    check(_lib.load().nm_ransac_f32(kind, _p(src_x), _p(src_y), _p(dst_x), _p(dst_y), src_x.shape[0], inlier_threshold,
                                    iterations, seed & 0xFFFFFFFFFFFFFFFF, _p(homography), _p(status), _stream_ptr()),
          "nm_ransac_f32")
    return homography, status


def ransac_batch(kind, src_x, src_y, dst_x, dst_y, counts, inlier_threshold, iterations, seed=0):
    """nm_ransac_batch_f32: src_x ... dst_y are (n_pairs, max_pts) cuda float32 (contiguous), counts (n_pairs,) cuda
    int32 or None.  Returns homographies (n_pairs, 9) and status (n_pairs, 3); pair p uses seed + p."""
    import torch
    _f32(src_x, src_y, dst_x, dst_y)
    if counts is not None:
        _int32(counts, "counts")
    n_pairs, max_pts = src_x.shape
    H = torch.zeros((n_pairs, 9), dtype=torch.float32, device=src_x.device)
    status = torch.zeros((n_pairs, 3), dtype=torch.int32, device=src_x.device)
    check(_lib.load().nm_ransac_batch_f32(kind, _p(src_x), _p(src_y), _p(dst_x), _p(dst_y), max_pts,
                                          _p(counts) if counts is not None else None, max_pts, n_pairs, inlier_threshold,
                                          iterations, seed & 0xFFFFFFFFFFFFFFFF, _p(H), _p(status), _stream_ptr()),
          "nm_ransac_batch_f32")
    return H, status

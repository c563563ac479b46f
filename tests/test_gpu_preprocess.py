"""GPU suite, input preprocessing: nm_grayscale_bgra_f32, nm_cast_f32_u8,
nm_undistort_map_f32, nm_resample_tex_f32 and the BGRA entry of the batched SIFT path (grey conversion fused
into the base blur), against the golden vectors captured from the reference (tests/golden/preprocess_160x96.npz),
the CPU oracle and -- when oracle/_ref is built -- the reference library live.  Integer / byte results and the
grey values are held to bitwise equality; the undistortion map to 2e-6 relative (powf)."""
import ctypes as C
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from niftymatch_b200 import synth  # noqa: E402
from tests._util import GOLDEN, _p, all_bgr_words, gray_double_formula, load_reflib  # noqa: E402
from tests.test_gpu_parity import run_product, assert_frame_matches  # noqa: E402


@pytest.fixture(scope="module")
def nm():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import niftymatch_b200 as nm
    nm.load()
    return nm


def _cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _gold():
    return np.load(os.path.join(GOLDEN, "preprocess_160x96.npz"))


def test_grayscale_all_colours_bitwise(nm):
    """All 2^24 (b, g, r): the dp4a + one fp32 division form equals the reference's double expression."""
    px = all_bgr_words()
    got = nm.grayscale(_cu(px)).cpu().numpy()
    assert np.array_equal(got, gray_double_formula(px))
    ref = load_reflib()
    if ref is not None:
        want = np.zeros((4096, 4096), np.float32)
        assert ref.lib.nmref_grayscale(_p(px), 4096, 4096, _p(want)) == 0
        assert np.array_equal(got, want)


def test_preprocess_vs_reference_golden(nm):
    g = _gold()
    assert np.array_equal(nm.grayscale(_cu(g["bgra"])).cpu().numpy(), g["gray"])
    # unaligned / odd sizes take the scalar tail
    sub = np.ascontiguousarray(g["bgra"][:7, :9])
    assert np.array_equal(nm.grayscale(_cu(sub)).cpu().numpy(), gray_double_formula(sub))
    for mv in (0, 200):
        assert np.array_equal(nm.cast_u8(_cu(g["fimg"]), mv).cpu().numpy(), g[f"cast_{mv}"]), mv
    u, v = nm.undistort_map(_cu(g["x"]), _cu(g["y"]), _cu(g["cam"]), _cu(g["dist"]))
    assert np.array_equal(u.cpu().numpy(), g["u"]) and np.array_equal(v.cpu().numpy(), g["v"])
    # pointers that are not 16-byte aligned and a length that is not a multiple of 4: the one-element kernels
    n = g["fimg"].size - 3
    f1 = _cu(g["fimg"]).flatten()[1: 1 + n].view(1, n)
    assert f1.data_ptr() % 16 == 4
    assert np.array_equal(nm.cast_u8(f1, 200).cpu().numpy().ravel(), g["cast_200"].ravel()[1: 1 + n])
    x1, y1 = _cu(g["x"]).flatten()[1: 1 + n].view(1, n), _cu(g["y"]).flatten()[1: 1 + n].view(1, n)
    u1, v1 = nm.undistort_map(x1, y1, _cu(g["cam"]), _cu(g["dist"]))
    assert np.array_equal(u1.cpu().numpy().ravel(), g["u"].ravel()[1: 1 + n])
    assert np.array_equal(v1.cpu().numpy().ravel(), g["v"].ravel()[1: 1 + n])
    # aligned start, ragged tail (quads + 1..3 single elements)
    f2 = _cu(g["fimg"]).flatten()[: n].view(1, n)
    assert np.array_equal(nm.cast_u8(f2, 0).cpu().numpy().ravel(), g["cast_0"].ravel()[: n])


def test_preprocess_vs_oracle(nm, oracle):
    g = _gold()
    h, w = g["fimg"].shape
    gray = np.zeros((h, w), np.float32)
    oracle.lib.orc_grayscale_bgra(_p(g["bgra"]), _p(gray), C.c_longlong(h * w))
    assert np.array_equal(nm.grayscale(_cu(g["bgra"])).cpu().numpy(), gray)
    for mv in (0, 200):
        c = np.zeros((h, w), np.uint8)
        oracle.lib.orc_cast_f32_u8(_p(g["fimg"]), C.c_longlong(h * w), _p(c), C.c_ubyte(mv))
        assert np.array_equal(nm.cast_u8(_cu(g["fimg"]), mv).cpu().numpy(), c), mv
    u, v = np.zeros((h, w), np.float32), np.zeros((h, w), np.float32)
    oracle.lib.orc_undistort_map(_p(g["x"]), _p(g["y"]), C.c_longlong(h * w), _p(g["cam"]), _p(g["dist"]), _p(u), _p(v))
    pu, pv = nm.undistort_map(_cu(g["x"]), _cu(g["y"]), _cu(g["cam"]), _cu(g["dist"]))
    assert np.abs(pu.cpu().numpy() - u).max() <= 2e-6 * np.abs(u).max()
    assert np.abs(pv.cpu().numpy() - v).max() <= 2e-6 * np.abs(v).max()


def _bgra_frames(n, w, h, seed0):
    """BGRA frames whose grey image has SIFT structure: the synthetic scene in the green and red channels, a
    shifted copy in blue, so that the three weights all matter."""
    out = np.zeros((n, h, w, 4), np.uint8)
    for f in range(n):
        s = synth.scene(w, h, seed0 + f).astype(np.uint8)
        out[f, ..., 0] = np.roll(s, 3, axis=1)
        out[f, ..., 1] = s
        out[f, ..., 2] = 255 - s // 2
        out[f, ..., 3] = 255
    return out


@pytest.mark.parametrize("shape", [(6, 256, 192), (12, 1920, 1080)])
def test_sift_on_bgra_frames_equals_grey_then_sift(nm, oracle, shape):
    """nm_sift_run_bgra: the small batch converts through the staging buffer, the 1080p batch takes the strip
    kernel with the conversion fused into the base blur.  Both equal grayscale -> nm_sift_run bit for bit, and
    the CPU oracle on the oracle's own grey frame."""
    n, w, h = shape
    bgra = _bgra_frames(n, w, h, synth.SEED_BASE + 40)
    P = nm.SiftParams(w, h)
    sb = nm.SiftBatch(P, n, 16384)
    sb.run_bgra(_cu(bgra))
    torch.cuda.synchronize()
    r = {k: v.cpu().numpy().copy() for k, v in sb.results().items()}
    lv = [sb.level(n - 1, 0, l).cpu().numpy().copy() for l in range(6)]
    gray = np.stack([nm.grayscale(_cu(bgra[f])).cpu().numpy() for f in range(n)])
    assert np.array_equal(gray, gray_double_formula(bgra))
    sb.run(_cu(gray))
    torch.cuda.synchronize()
    r2 = {k: v.cpu().numpy() for k, v in sb.results().items()}
    for l in range(6):
        assert np.array_equal(lv[l], sb.level(n - 1, 0, l).cpu().numpy()), l
    assert np.array_equal(r["counts"], r2["counts"]) and (r["counts"] > 50).all()
    for f in range(n):
        c = r["counts"][f]
        assert np.array_equal(r["kpts"][f, :c], r2["kpts"][f, :c])
        assert np.array_equal(r["desc"][f, :c], r2["desc"][f, :c])
    sb.close()
    o = oracle.sift_frame(gray[n - 1], capacity=16384)
    p = run_product(nm, gray[n - 1:n], capacity=16384)[0]
    assert_frame_matches(p, o)
    assert np.array_equal(lv[0], o["levels"][0][0])


def test_dropin_preprocess_headers(nm):
    """bgra_2_gray.h / cast.h / undistort.h / resample.h of the drop-in layer through the client code that
    drives the reference (oracle/ref_preprocess_driver.cu built with -DNM_COMPAT_BUILD)."""
    client = os.path.abspath(os.path.join(os.path.dirname(GOLDEN), os.pardir, "build", "compat", "libnmcompat.so"))
    if not os.path.exists(client):
        pytest.skip("build/compat/libnmcompat.so not built")
    cl = C.CDLL(client)
    g = _gold()
    h, w = g["fimg"].shape
    gray = np.zeros((h, w), np.float32)
    assert cl.nmcompat_grayscale(_p(g["bgra"]), w, h, _p(gray)) == 0 and np.array_equal(gray, g["gray"])
    c = np.zeros((h, w), np.uint8)
    assert cl.nmcompat_cast(_p(g["fimg"]), w, h, _p(c), 200) == 0 and np.array_equal(c, g["cast_200"])
    u, v = np.zeros((h, w), np.float32), np.zeros((h, w), np.float32)
    assert cl.nmcompat_undistort(_p(g["x"]), _p(g["y"]), w, h, _p(g["cam"]), _p(g["dist"]), _p(u), _p(v)) == 0
    assert np.array_equal(u, g["u"]) and np.array_equal(v, g["v"])
    res = np.zeros((h, w), np.float32)
    assert cl.nmcompat_resample_undistort(_p(g["gray8"]), w, h, _p(u), _p(v), w, h, _p(res)) == 0
    assert np.array_equal(res, g["resampled"])          # same texture unit, same filtering arithmetic


def test_channel_helpers_and_bgra_downsample_vs_reference_golden(nm):
    """cuda_extract_channel / cuda_put_channel / cuda_set_alpha_to_const / downsample_by_2<uchar4> of the drop-in
    headers, through the client code that drives the reference."""
    client = os.path.abspath(os.path.join(os.path.dirname(GOLDEN), os.pardir, "build", "compat", "libnmcompat.so"))
    if not os.path.exists(client):
        pytest.skip("build/compat/libnmcompat.so not built")
    cl = C.CDLL(client)
    g = _gold()
    h, w = g["fimg"].shape
    chans, rot = np.zeros((4, h, w), np.float32), np.zeros((h, w, 4), np.uint8)
    assert cl.nmcompat_channels(_p(g["bgra"]), w, h, 77, _p(chans), _p(rot)) == 0
    assert np.array_equal(chans, g["channels"]) and np.array_equal(rot, g["rotated_alpha77"])
    assert np.array_equal(chans, np.moveaxis(g["bgra"], 2, 0).astype(np.float32))
    half = np.zeros((h // 2, w // 2, 4), np.uint8)
    assert cl.nmcompat_downsample_bgra(_p(g["bgra"]), w, h, _p(half)) == 0
    assert np.array_equal(half, g["bgra_half"]) and np.array_equal(half, g["bgra"][::2, ::2][: h // 2, : w // 2])


def test_preprocess_bad_arguments(nm):
    lib = nm.load()
    t = torch.zeros(64, device="cuda")
    p = C.c_void_p(t.data_ptr())
    assert lib.nm_grayscale_bgra_f32(None, p, 4, 4, None) == -1
    assert lib.nm_grayscale_bgra_f32(p, p, 0, 4, None) == 0
    assert lib.nm_grayscale_bgra_f32(C.c_void_p(t.data_ptr() + 1), p, 2, 2, None) == -1
    assert lib.nm_cast_f32_u8(p, -1, 4, p, 0, None) == -1
    assert lib.nm_undistort_map_f32(p, p, 4, 4, None, p, p, p, None) == -1
    assert lib.nm_resample_tex_f32(0, p, p, 4, 4, p, None) == -1
    assert lib.nm_sift_run_bgra(None, p, 1, None) == -1

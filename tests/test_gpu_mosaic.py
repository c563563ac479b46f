"""GPU suite, mosaic rendering: resample_perspective_transform, resample_mask and
transform_blend of the drop-in resample.h -> nm_resample_perspective_bgra / nm_resample_mask_tex_u8 /
nm_transform_blend_bgra, driven by the client code that drives the reference (oracle/ref_mosaic_driver.cu,
-DNM_COMPAT_BUILD), against the golden vectors the reference produced on a B200 (tests/golden/mosaic_128x90.npz).
The texture unit is the same hardware in both, so everything is held to bitwise equality."""
import ctypes as C
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from tests._util import GOLDEN, _p  # noqa: E402

CLIENT = os.path.abspath(os.path.join(os.path.dirname(GOLDEN), os.pardir, "build", "compat", "libnmcompat.so"))


@pytest.fixture(scope="module")
def client():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not os.path.exists(CLIENT):
        pytest.skip("build/compat/libnmcompat.so not built")
    return C.CDLL(CLIENT)


def test_perspective_resample_and_mask_vs_reference_golden(client):
    g = np.load(os.path.join(GOLDEN, "mosaic_128x90.npz"))
    fh, fw = g["mask"].shape
    rows, cols = g["xpos_0"].shape
    for inv in (0, 1):
        res = np.zeros((rows, cols, 4), np.uint8)
        xp, yp = np.zeros((rows, cols), np.float32), np.zeros((rows, cols), np.float32)
        assert client.nmcompat_resample_perspective(_p(g["frame"]), fw, fh, _p(g["mats"][1]), inv, cols, rows, _p(res), _p(xp), _p(yp)) == 0
        assert np.array_equal(xp, g[f"xpos_{inv}"]) and np.array_equal(yp, g[f"ypos_{inv}"])
        assert np.array_equal(res, g[f"persp_{inv}"])
        m = np.zeros((rows, cols), np.uint8)
        assert client.nmcompat_resample_mask(_p(g["mask"]), fw, fh, _p(xp), _p(yp), cols, rows, C.c_float(0.5), _p(m)) == 0
        assert np.array_equal(m, g[f"maskres_{inv}"]) and 0 < (m > 0).sum() < m.size


def test_transform_blend_vs_reference_golden(client):
    g = np.load(os.path.join(GOLDEN, "mosaic_128x90.npz"))
    fh, fw = g["mask"].shape
    ch, cw = g["canvas_wts"].shape
    canvas, cwts = np.zeros((ch, cw, 4), np.uint8), np.zeros((ch, cw), np.float32)
    assert client.nmcompat_transform_blend(_p(g["frame"]), _p(g["mask"]), _p(g["wts"]), fw, fh, 3, _p(g["mats"]), _p(g["tx"]),
                                           _p(g["ty"]), fw + 10, fh + 10, cw, ch, _p(canvas), _p(cwts)) == 0
    assert np.array_equal(cwts, g["canvas_wts"])
    assert np.array_equal(canvas, g["canvas"])
    assert (cwts > 0).sum() > 5000 and (canvas[..., 3][cwts > 0] == 255).all()


def test_mosaic_bad_arguments():
    import niftymatch_b200 as nm
    lib = nm.load()
    t = torch.zeros(64, device="cuda")
    p = C.c_void_p(t.data_ptr())
    assert lib.nm_resample_perspective_bgra(p, 0, 4, 4, p, p, p, 1, None) == -1
    assert lib.nm_resample_perspective_bgra(p, 0, 0, 4, p, p, p, 1, None) == 0
    assert lib.nm_resample_mask_tex_u8(p, 0, 4, 4, p, p, 0.5, None) == -1
    assert lib.nm_transform_blend_bgra(p, 4, 4, 0, 4, 4, 4, 4, p, 0, 0, 0, p, 0, None) == -1

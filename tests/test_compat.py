"""Drop-in layer (compat/): the reference's header names, library names and CMake package
variables (reference src/cmake/NiftyMatchConfig.cmake:13-46, src/CMakeLists.txt:63-88), and the
reference-driving client (oracle/ref_driver.cu) built UNCHANGED against it."""
import ctypes as C
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PREFIX = os.path.join(ROOT, "build", "compat", "prefix")
CLIENT = os.path.join(ROOT, "build", "compat", "libnmcompat.so")

HOT_PATH_HEADERS = ["siftparams.h", "pyramidata.h", "siftdata.h", "siftfunctions.h", "convolution.h", "downsample.h",
                    "cudamath.h", "keypoint.h", "orientation.h", "descriptor.h", "match.h", "transpose.h",
                    "cudatex2D.h", "cudatimer.h", "exception.h", "macros.h", "ransac.h", "bgra_2_gray.h", "cast.h", "undistort.h",
                    "resample.h", "cudautils.h"]


def _built():
    if not os.path.exists(os.path.join(PREFIX, "lib", "nm", "libsift.a")):
        pytest.skip("compat tree not built (make compat-client)")


def test_install_layout():
    _built()
    for h in HOT_PATH_HEADERS + ["NiftyMatchConfig.cmake", "nm_b200.h"]:
        assert os.path.exists(os.path.join(PREFIX, "include", "nm", h)), h
    for lib in ("libgpuutils.a", "libkernels.a", "libsift.a", "libnm_b200.so"):
        assert os.path.exists(os.path.join(PREFIX, "lib", "nm", lib)), lib


def test_exported_cxx_symbols():
    """The C++ symbols a reference client links against are defined by the three archives."""
    _built()
    out = ""
    for lib in ("libgpuutils.a", "libkernels.a", "libsift.a"):
        out += subprocess.run(["nm", "-C", "--defined-only", os.path.join(PREFIX, "lib", "nm", lib)],
                              capture_output=True, text=True, check=True).stdout
    for sym in ["compute_sift_matches(SiftData*, SiftData*, float*, float, CUstream_st*)",
                "compute_dog(PyramidData&, int, int, CUstream_st*)",
                "compute_gradients(PyramidData&, SiftParams const&, int, int, CUstream_st*)",
                "compute_keypoints(PyramidData&, SiftParams const&, int, int, int, CUstream_st*)",
                "compute_keypoints_with_mask(PyramidData&, SiftParams&, unsigned long long, int, int, int, CUstream_st*)",
                "compute_orientations(PyramidData&, SiftParams const&, int, int, int, CUstream_st*)",
                "compute_descriptors(PyramidData&, SiftParams const&, int, int, int, SiftData&, CUstream_st*)",
                "PyramidData::PyramidData(SiftParams const&)", "PyramidData::initialize(SiftParams const&)",
                "PyramidData::gpu_collate_keypoints_for_level(int, int)", "SiftData::SiftData(int)",
                "SiftData::copy_from(SiftData const&)", "SiftData::initialize_data(int)", "SiftData::clear_data()",
                "void convolve<float>(", "void downsample_by_2<float>(", "void subtract<float>(", "void gradient<float>(",
                "void transpose<float>(", "void compute_brute_force_distance<float>(", "void get_sift_matches<float>(",
                "detect_orientations(", "compute_sift_descriptors(", "find_keypoints(", "CudaTex2D::set(", "CudaTimer::stop()",
                "DivUp", "AlignDown", "align_points(", "ransac_homography(", "ransac_translation(", "ransac_similarity(",
                "void cuda_grayscale<float>(", "void cuda_cast<float, unsigned char>(", "cuda_undistort(", "resample_undistort(",
                "resample_perspective_transform(", "resample_mask(", "transform_blend(", "CudaUtils::get_max_flops_device_id()",
                "CudaUtils::setup_CUDA(int)", "void cuda_extract_channel<float>(", "void cuda_put_channel<float>(", "cuda_set_alpha_to_const(",
                "void downsample_by_2<uchar4>("]:
        assert sym in out, sym


def test_find_package_variables():
    """FIND_PACKAGE(NiftyMatch CONFIG) with NiftyMatch_DIR=<prefix>/include/nm (reference README.md:20-22)."""
    _built()
    cmake = shutil.which("cmake")
    if cmake is None:
        pytest.skip("cmake not available")
    with tempfile.TemporaryDirectory() as tmp:
        r = subprocess.run([cmake, "-S", os.path.join(ROOT, "tests", "cmake_consumer"), "-B", tmp,
                            f"-DNiftyMatch_DIR={os.path.join(PREFIX, 'include', 'nm')}"],
                           capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    vars_ = dict(l.split("NM_VAR ", 1)[1].split("=", 1) for l in r.stdout.splitlines() if "NM_VAR " in l)
    assert vars_["FOUND"] in ("1", "TRUE")
    assert os.path.samefile(vars_["INCLUDE_DIR"], os.path.join(PREFIX, "include", "nm"))
    assert vars_["PATH_SUFFIX"] == "nm"
    for k in ("gpuutils", "kernels", "sift"):
        assert vars_[k].endswith(f"lib/nm/lib{k}.a"), vars_[k]
        assert vars_[k] in vars_["LIBS"]


@pytest.mark.gpu
def test_reference_client_on_dropin_matches_reference_golden(oracle):
    """The client loop that produced tests/golden from the reference library, now linked against the
    drop-in: same keypoints (bitwise), orientations / descriptors within the north_star tolerances."""
    if not os.path.exists(CLIENT):
        pytest.skip("build/compat/libnmcompat.so not built")
    from tests._util import FrameChecker, GOLDEN, ang_diff
    cl = FrameChecker(C.CDLL(CLIENT), "nmcompat")
    g = np.load(os.path.join(GOLDEN, "sift_256x192.npz"))
    img = g["image"]
    r = cl.sift_frame(img, peak=0.0, orient_mode=0, want_levels=True)
    o = oracle.sift_frame(img, peak=0.0, orient_mode=0, want_levels=True)
    assert r["n"] == o["n"] > 100
    assert np.array_equal(r["seg_counts"], o["seg_counts"])
    for oc in range(r["n_oct"]):
        for l in range(6):
            assert np.array_equal(r["levels"][oc][l], o["levels"][oc][l]), (oc, l)
    assert np.array_equal(r["kpts"], o["kpts"])
    ok = o["orient"][:, 0] >= 0
    assert ang_diff(r["orient"][ok, 0], o["orient"][ok, 0]).max() < 1e-3
    rel = np.linalg.norm(r["desc"] - o["desc"], axis=1) / np.linalg.norm(o["desc"], axis=1)
    assert rel.max() < 1e-3, rel.max()
    assert np.array_equal(r["x"], o["x"]) and np.array_equal(r["y"], o["y"])
    # ... and against what the reference library itself produced on a B200 (tests/golden)
    assert np.array_equal(r["seg_counts"], g["p0_seg_counts"][: len(r["seg_counts"])])
    assert np.array_equal(r["kpts"], g["p0_kpts"])
    relg = np.linalg.norm(r["desc"] - g["p0_desc"], axis=1) / np.linalg.norm(g["p0_desc"], axis=1)
    assert relg.max() < 1e-3, relg.max()
    # capacity truncation rule through the client (SiftData capacity 100)
    rt = cl.sift_frame(img, peak=0.0, orient_mode=0, capacity=100, want_levels=False)
    assert rt["n"] == 100 and np.array_equal(rt["x"], o["x"][:100])


@pytest.mark.gpu
def test_reference_client_masked_detector_on_dropin(oracle):
    """compute_keypoints_with_mask through the drop-in headers / libraries, driven by the same client code
    that drove the reference (a float cudaArray behind CudaTex2D): keypoints bitwise the reference's."""
    if not os.path.exists(CLIENT):
        pytest.skip("build/compat/libnmcompat.so not built")
    from tests._util import FrameChecker, GOLDEN
    cl = FrameChecker(C.CDLL(CLIENT), "nmcompat")
    g = np.load(os.path.join(GOLDEN, "sift_256x192_masked.npz"))
    img = g["image"]
    for name in ("fov", "soft"):
        r = cl.sift_frame(img, peak=0.0, orient_mode=0, want_levels=False, mask=g[f"{name}_mask"])
        assert np.array_equal(r["seg_counts"], g[f"{name}_seg_counts"][: len(r["seg_counts"])]), name
        assert np.array_equal(r["kpts"], g[f"{name}_kpts"]), name
        rel = np.linalg.norm(r["desc"] - g[f"{name}_desc"], axis=1) / np.linalg.norm(g[f"{name}_desc"], axis=1)
        assert rel.max() < 1e-3, rel.max()
        assert np.array_equal(r["x"], g[f"{name}_x"]) and np.array_equal(r["y"], g[f"{name}_y"])


@pytest.mark.gpu
def test_reference_client_matcher_on_dropin():
    if not os.path.exists(CLIENT):
        pytest.skip("build/compat/libnmcompat.so not built")
    from tests._util import FrameChecker, GOLDEN
    cl = FrameChecker(C.CDLL(CLIENT), "nmcompat")
    g = np.load(os.path.join(GOLDEN, "match_200x250.npz"))
    m, D = cl.match(g["A"], g["B"], 0.8, match_io=g["m0"], want_distance=True)
    assert np.array_equal(m, g["m"])
    assert np.array_equal(D, g["D"])


@pytest.mark.gpu
def test_dropin_matcher_reaches_the_tensor_core_engine_without_the_distance_matrix():
    """compute_sift_matches fills the caller's distance matrix bitwise (exact engine).  NM_COMPAT_SKIP_DISTANCE=1 (or a
    null `distance`) skips the matrix and lets the drop-in take the tcgen05 engine: same match indices.  3000 x 3000
    descriptors = 9M pairs, above the auto-engine threshold; run in a child process because the switch is read once."""
    if not os.path.exists(CLIENT):
        pytest.skip("build/compat/libnmcompat.so not built")
    import sys
    from niftymatch_b200 import synth
    from tests._util import FrameChecker
    B = synth.descriptors(3000, 2)
    A = synth.descriptors(3000, 1, planted_from=B)
    m_exact = FrameChecker(C.CDLL(CLIENT), "nmcompat").match(A, B, 0.8)
    code = ("import ctypes as C, numpy as np, sys; sys.path.insert(0, %r);"
            "from niftymatch_b200 import synth; from tests._util import FrameChecker;"
            "B = synth.descriptors(3000, 2); A = synth.descriptors(3000, 1, planted_from=B);"
            "m = FrameChecker(C.CDLL(%r), 'nmcompat').match(A, B, 0.8); np.save(sys.argv[1], m)" % (ROOT, CLIENT))
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "m.npy")
        r = subprocess.run([sys.executable, "-c", code, out], env=dict(os.environ, NM_COMPAT_SKIP_DISTANCE="1"),
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        m_tc = np.load(out)
    assert (m_exact >= 0).sum() > 300
    assert np.array_equal(m_tc, m_exact)

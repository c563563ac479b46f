"""GPU suite: brute-force k=2 ratio-test matching through the C-ABI."""
import ctypes as C
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from niftymatch_b200 import synth  # noqa: E402
from tests._util import GOLDEN  # noqa: E402


@pytest.fixture(scope="module")
def nm():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import niftymatch_b200 as nm
    nm.load()
    return nm


def _cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _engines(nm):
    e = [0]
    try:
        nm.set_engine(1)
        e.append(1)
    except nm.NmError:
        pass
    nm.set_engine(-1)
    return e


def test_match_vs_reference_golden(nm):
    g = np.load(os.path.join(GOLDEN, "match_200x250.npz"))
    for eng in _engines(nm):
        nm.set_engine(eng)
        m = nm.match(_cu(g["A"]), _cu(g["B"]), 0.8, match_io=_cu(g["m0"]))
        assert np.array_equal(m.cpu().numpy(), g["m"]), f"engine {eng}"
    nm.set_engine(-1)
    m, D = nm.match(_cu(g["A"]), _cu(g["B"]), 0.8, match_io=_cu(g["m0"]), want_distance=True)
    assert np.array_equal(D.cpu().numpy(), g["D"]), "distance matrix not bitwise equal to the reference's"
    assert np.array_equal(m.cpu().numpy(), g["m"])


@pytest.mark.parametrize("nA,nB", [(1, 1), (1, 2), (3, 1), (64, 64), (65, 129), (500, 700), (1000, 37), (2048, 2048)])
def test_match_vs_oracle(nm, oracle, nA, nB):
    B = synth.descriptors(nB, 100 + nB)
    A = synth.descriptors(nA, 200 + nA, planted_from=B)
    mo = oracle.match(A, B, 0.8)
    for eng in _engines(nm):
        nm.set_engine(eng)
        m = nm.match(_cu(A), _cu(B), 0.8)
        assert np.array_equal(m.cpu().numpy(), mo), f"engine {eng}"
    nm.set_engine(-1)


def test_compat_matcher_operators(nm, oracle):
    """transpose + compute_brute_force_distance (A dim-major in, D^T out) + get_sift_matches,
    as compute_sift_matches chains them (reference siftfunctions.cu:15-40)."""
    import niftymatch_b200._lib as L
    lib = L.load()
    B = synth.descriptors(150, 7)
    A = synth.descriptors(90, 8, planted_from=B)
    At, Bt = _cu(A), _cu(B)
    A_T = torch.empty((128, 90), dtype=torch.float32, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert lib.nm_transpose_f32(C.c_void_p(A_T.data_ptr()), C.c_void_p(At.data_ptr()), 128, 90, st) == 0
    assert np.array_equal(A_T.cpu().numpy(), A.T)
    D_T = torch.empty((150, 90), dtype=torch.float32, device="cuda")
    assert lib.nm_dist2_f32(C.c_void_p(A_T.data_ptr()), 90, C.c_void_p(Bt.data_ptr()), 150, 128, C.c_void_p(D_T.data_ptr()), st) == 0
    mo, Do = oracle.match(A, B, 0.8, want_distance=True)
    assert np.array_equal(D_T.cpu().numpy().T, Do)
    D = D_T.t().contiguous()
    m = torch.full((90,), -1, dtype=torch.int32, device="cuda")
    assert lib.nm_set_matches_f32(C.c_void_p(D.data_ptr()), 90, 150, 150, C.c_void_p(m.data_ptr()), 0.8, st) == 0
    assert np.array_equal(m.cpu().numpy(), mo)


@pytest.mark.parametrize("shards", [2, 3, 8])
def test_sharded_records_merge_bit_identical(nm, oracle, shards):
    """Database sharding emulated on one GPU: per-shard records + merge == unsharded result
    (the N>1 collective path all-gathers exactly these records)."""
    from niftymatch_b200.dist import shard_bounds
    B = synth.descriptors(1000, 31)
    A = synth.descriptors(600, 32, planted_from=B)
    At = _cu(A)
    single = nm.match(At, _cu(B), 0.8).cpu().numpy()
    recs = []
    for s in range(shards):
        lo, hi = shard_bounds(len(B), shards, s)
        recs.append(nm.match_top2(At, _cu(B[lo:hi]), lo))
    m = nm.merge_top2(torch.stack(recs).contiguous(), 0.8).cpu().numpy()
    assert np.array_equal(m, single)
    assert np.array_equal(m, oracle.match(A, B, 0.8))


def test_invalid_arguments_return_codes(nm):
    import niftymatch_b200._lib as L
    lib = L.load()
    a = torch.zeros((4, 128), device="cuda")
    m = torch.zeros(4, dtype=torch.int32, device="cuda")
    assert lib.nm_match_f32(None, 4, C.c_void_p(a.data_ptr()), 4, 0.8, C.c_void_p(m.data_ptr()), None, None) == -1
    assert lib.nm_match_f32(C.c_void_p(a.data_ptr()), 0, C.c_void_p(a.data_ptr()), 4, 0.8, C.c_void_p(m.data_ptr()), None, None) == -1
    assert lib.nm_blur_f32(C.c_void_p(a.data_ptr()), C.c_void_p(a.data_ptr()), None, 4, 4, C.c_void_p(a.data_ptr()), 46, None) == -1


def test_large_match_properties(nm, oracle):
    """20k x 20k: planted rows must match their source; a row subset is checked against the
    oracle exactly."""
    n = 20000
    B = synth.descriptors(n, 2)
    A = synth.descriptors(n, 1, planted_from=B)
    m = nm.match(_cu(A), _cu(B), 0.8).cpu().numpy()
    sub = np.arange(0, n, 313)
    mo = oracle.match(np.ascontiguousarray(A[sub]), B, 0.8)
    assert np.array_equal(m[sub], mo)
    assert (m >= 0).sum() > 0.15 * n

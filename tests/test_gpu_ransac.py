"""GPU suite, registration after matching: nm_align_points_f32, nm_ransac_hypotheses_f32
and nm_ransac_f32 through the C-ABI, against the golden vectors captured from the reference's own kernels
(tests/golden/ransac_400.npz: translation_kernel / similarity_transformation_kernel / homography_kernel of
ransac.cu:437-520 launched on supplied index lists), the CPU oracle, and -- when oracle/_ref is built -- the
reference library live.

The hypothesis arithmetic is fp32 with a few double intermediates (svd.cu:291-292); the product writes the
same expressions and is compiled by the same nvcc, so its homographies and inlier counts are held to BITWISE
equality with the reference's kernels.  The CPU oracle (explicit fmaf where the GPU build contracts) is allowed 5e-4 on the
Frobenius-normalised homography and +-2 inliers (see tests/test_oracle_golden.py)."""
import ctypes as C
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from tests._util import (GOLDEN, ransac_scene, ransac_rand_lists, checker_ransac_hypotheses, normalise_h, _p)  # noqa: E402


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
    return np.load(os.path.join(GOLDEN, "ransac_400.npz"))


def test_align_points_vs_reference_golden(nm):
    g = _gold()
    out = nm.align_points(_cu(g["al_src_x"]), _cu(g["al_src_y"]), _cu(g["al_dst_x"]), _cu(g["al_dst_y"]), _cu(g["al_matches"]))
    for got, key in zip(out, ("al_c_src_x", "al_c_src_y", "al_c_dst_x", "al_c_dst_y")):
        assert np.array_equal(got.cpu().numpy(), g[key]), key
    assert (g["al_matches"] == -1).sum() > 20
    # empty input is a no-op
    lib = nm.load()
    assert lib.nm_align_points_f32(None, None, None, None, None, None, None, None, None, 0, None) == 0
    assert lib.nm_align_points_f32(None, None, None, None, None, None, None, None, None, 5, None) == -1


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_hypotheses_bitwise_vs_reference_golden(nm, kind):
    g = _gold()
    H, inl = nm.ransac_hypotheses(kind, _cu(g["src_x"]), _cu(g["src_y"]), _cu(g["dst_x"]), _cu(g["dst_y"]),
                                  _cu(g[f"rand_{kind}"]), float(g["thr"]))
    assert np.array_equal(inl.cpu().numpy(), g[f"inliers_{kind}"])
    assert np.array_equal(H.cpu().numpy(), g[f"H_{kind}"])
    if kind:
        zero = (g[f"H_{kind}"] == 0).all(axis=1)
        assert zero[5] and zero[17] and (g[f"inliers_{kind}"][zero] == 0).all()      # planted repeated indices


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_hypotheses_vs_oracle_and_live_reference(nm, oracle, kind):
    """Other scenes / sizes (2 500 correspondences: three correspondence blocks in the scoring kernel)."""
    from tests._util import load_reflib
    ref = load_reflib()
    for n, seed in ((2500, 21), (37, 22)):
        sx, sy, dx, dy, _ = ransac_scene(n=n, seed=seed)
        rl = ransac_rand_lists(sx, iterations=160, seed=seed)[kind]
        H, inl = nm.ransac_hypotheses(kind, _cu(sx), _cu(sy), _cu(dx), _cu(dy), _cu(rl), 3.0)
        H, inl = H.cpu().numpy(), inl.cpu().numpy()
        Ho, io = checker_ransac_hypotheses(oracle.lib, "orc", kind, sx, sy, dx, dy, rl, 3.0)
        assert np.array_equal((H == 0).all(axis=1), (Ho == 0).all(axis=1))
        # hypotheses from (nearly) degenerate samples have an ill-conditioned null vector: compare the
        # well-conditioned ones tightly, all of them through their scores
        d = np.abs(normalise_h(H) - normalise_h(Ho)).max(axis=1)
        assert np.median(d) < 1e-5 and (d < 5e-4).mean() > 0.95, (np.median(d), d.max())
        assert (np.abs(inl - io) <= 2).mean() > 0.97 and abs(int(inl.max()) - int(io.max())) <= 2
        if ref is not None:
            Hr, ir = checker_ransac_hypotheses(ref.lib, "nmref", kind, sx, sy, dx, dy, rl, 3.0)
            assert np.array_equal(inl, ir) and np.array_equal(H, Hr)


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_full_ransac_is_reproducible_and_finds_the_model(nm, kind):
    sx, sy, dx, dy, Ht = ransac_scene(n=1500, seed=5, noise=0.2)
    a = [_cu(v) for v in (sx, sy, dx, dy)]
    H1, st1 = nm.ransac(kind, *a, 4.0, 1500, seed=1234)
    H2, st2 = nm.ransac(kind, *a, 4.0, 1500, seed=1234)
    H3, st3 = nm.ransac(kind, *a, 4.0, 1500, seed=99)
    assert torch.equal(H1, H2) and torch.equal(st1, st2)            # same seed, same answer
    ok, n_in, it = st1.cpu().tolist()
    assert ok == 1 and 0 <= it < 1500
    assert st3[2].item() != it or not torch.equal(H1, H3)           # another seed draws other samples
    # the chosen model's score is what the scoring rule gives for it
    H = H1.cpu().numpy().astype(np.float64).reshape(3, 3)
    v = sx >= 0
    q = H @ np.stack([sx[v], sy[v], np.ones(v.sum())])
    err2 = (dx[v] - q[0] / q[2]) ** 2 + (dy[v] - q[1] / q[2]) ** 2
    assert abs(int((err2 < 4.0).sum()) - n_in) <= 2
    if kind == 2:
        # 70 % of the valid correspondences follow the homography: nearly all of them are inliers of the best model
        true_in = int(v.sum() * 0.7)
        assert n_in > 0.9 * true_in, (n_in, true_in)
        assert np.abs(H / H[2, 2] - Ht).max() < 2.0 and np.abs((H / H[2, 2] - Ht)[:2, :2]).max() < 5e-3


def _draws(seed, n_draws, valid):
    """The index draws of nm_ransac_f32 (nm_ransac.cu draw_kernel: splitmix64 of seed ^ K*(d+1), high word
    scaled to [0, n_valid)) restated with numpy."""
    with np.errstate(over="ignore"):
        d = np.arange(1, n_draws + 1, dtype=np.uint64)
        x = np.uint64(seed) ^ (np.uint64(0xD1B54A32D192ED03) * d)
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
        r = x >> np.uint64(32)
        idx = (r * np.uint64(len(valid))) >> np.uint64(32)
    return valid[idx.astype(np.int64)].astype(np.int32)


def test_full_ransac_equals_hypotheses_on_its_own_draws(nm):
    """nm_ransac_f32 = valid list -> draws -> hypotheses -> first maximum.  With the draws restated on the host,
    nm_ransac_hypotheses_f32 (pinned bitwise to the reference's kernels above) must give the same winner."""
    sx, sy, dx, dy, _ = ransac_scene(n=1300, seed=8)
    a = [_cu(v) for v in (sx, sy, dx, dy)]
    valid = np.nonzero(sx >= 0)[0]
    for kind, m in ((0, 1), (1, 2), (2, 4)):
        for seed in (7, 2 ** 40 + 3):
            H, st = nm.ransac(kind, *a, 2.0, 700, seed=seed)
            rl = _draws(seed, 700 * m, valid)
            Hall, inl = nm.ransac_hypotheses(kind, *a, _cu(rl), 2.0)
            inl = inl.cpu().numpy()
            best = int(inl.argmax())                       # numpy: first maximum, like thrust::max_element
            assert st.cpu().tolist() == [1, int(inl[best]), best]
            assert torch.equal(H, Hall[best])


def test_too_few_correspondences_and_bad_arguments(nm):
    lib = nm.load()
    sx = np.full(50, -1.0, np.float32)
    sx[7] = 10.0                                       # one valid correspondence only
    sy, dx, dy = sx.copy(), sx.copy(), sx.copy()
    a = [_cu(v) for v in (sx, sy, dx, dy)]
    for kind in (0, 1, 2):
        H0 = torch.full((9,), 7.0, device="cuda")
        H, st = nm.ransac(kind, *a, 4.0, 64, seed=3, homography=H0)
        assert st.cpu().tolist() == [0, 0, -1]         # the reference's `return false`
        assert (H.cpu().numpy() == 7.0).all()          # homography untouched
    sx[[3, 9, 20]] = [1.0, 5.0, 9.0]                   # four valid: enough for all estimators
    a = [_cu(v) for v in (sx, sx.copy(), sx.copy(), sx.copy())]
    assert nm.ransac(2, *a, 4.0, 64, seed=3)[1][0].item() == 1
    p = C.c_void_p(a[0].data_ptr())
    assert lib.nm_ransac_f32(3, p, p, p, p, 50, 4.0, 10, 0, p, p, None) == -1
    assert lib.nm_ransac_f32(0, p, p, p, p, 0, 4.0, 10, 0, p, p, None) == -1
    assert lib.nm_ransac_f32(0, p, p, p, p, 50, 4.0, 0, 0, p, p, None) == -1
    assert lib.nm_ransac_f32(0, p, p, p, p, 50, 4.0, 10, 0, None, p, None) == -1
    assert lib.nm_ransac_hypotheses_f32(0, p, p, p, p, 50, None, 10, 4.0, p, p, None) == -1
    # an index outside the arrays in a caller's list: that iteration is skipped (H = 0, 0 inliers), nothing is read
    rl = torch.tensor([3, 9, 20, 7, 3, 9, 20, 50, -1, 9, 20, 7], dtype=torch.int32, device="cuda")
    H, inl = nm.ransac_hypotheses(2, *a, rl, 4.0)
    assert (H[1:] == 0).all().item() and inl[1:].tolist() == [0, 0]


def test_sift_match_ransac_chain(nm):
    """The consumer chain of SURVEY.md 8f rank 1 on the product: SIFT on a frame and its shifted copy, match,
    align_points, ransac_translation -> the known shift."""
    from niftymatch_b200 import synth
    base = synth.scene(640 + 16, 480 + 16, synth.SEED_BASE + 2)
    f0, f1 = base[8:488, 8:648], base[5:485, 2:642]     # f1(x, y) = f0(x - 6, y - 3)
    P = nm.SiftParams(640, 480)
    P._peak_threshold = 2.0
    sb = nm.SiftBatch(P, 2, 8192)
    sb.run(_cu(np.stack([f0, f1])))
    r = sb.results()
    n0, n1 = int(r["counts"][0]), int(r["counts"][1])
    assert n0 > 200 and n1 > 200
    m = nm.match(r["desc"][0, :n0].contiguous(), r["desc"][1, :n1].contiguous(), 0.8)
    c = nm.align_points(r["x"][0, :n0].contiguous(), r["y"][0, :n0].contiguous(), r["x"][1, :n1].contiguous(),
                        r["y"][1, :n1].contiguous(), m)
    H, st = nm.ransac(nm.TRANSLATION, *c, 1.0, 256, seed=1)
    ok, n_in, _ = st.cpu().tolist()
    H = H.cpu().numpy()
    assert ok == 1 and n_in > 0.8 * int((m >= 0).sum().item()) > 50
    # a single-correspondence model: exact for an octave-0 keypoint, a few tenths of a pixel off for one of a
    # coarser octave (the 6-px shift is not a multiple of that octave's sampling step)
    assert abs(H[2] - 6.0) < 0.5 and abs(H[5] - 3.0) < 0.5, H
    Hh, sth = nm.ransac(nm.HOMOGRAPHY, *c, 1.0, 512, seed=1)
    Hh = Hh.cpu().numpy()
    Hh = Hh / Hh[8]
    assert abs(Hh[2] - 6.0) < 1.0 and abs(Hh[5] - 3.0) < 1.0 and abs(Hh[0] - 1) < 5e-3 and abs(Hh[4] - 1) < 5e-3, Hh
    sb.close()


def test_batched_ransac_equals_single_calls(nm):
    """nm_ransac_batch_f32: pair p == nm_ransac_f32 on that pair with seed + p, for ragged counts, a pair with too
    few valid correspondences and an empty pair."""
    max_pts, n_pairs = 700, 6
    arrs = [np.full((n_pairs, max_pts), -1.0, np.float32) for _ in range(4)]
    counts = np.array([700, 523, 64, 3, 0, 300], np.int32)
    for p in range(n_pairs):
        sc = ransac_scene(n=max_pts, seed=30 + p)
        for a, v in zip(arrs, sc[:4]):
            a[p, : counts[p]] = v[: counts[p]]
    arrs[0][3, :3] = [5.0, -1.0, 7.0]                    # pair 3: two valid correspondences only
    d = [_cu(a) for a in arrs]
    for kind in (0, 1, 2):
        H, st = nm.ransac_batch(kind, *d, _cu(counts), 3.0, 300, seed=77)
        H, st = H.cpu().numpy(), st.cpu().numpy()
        for p in range(n_pairs):
            if counts[p] == 0:
                assert st[p].tolist() == [0, 0, -1] and (H[p] == 0).all()
                continue
            one = [x[p, : counts[p]].contiguous() for x in d]
            H1, st1 = nm.ransac(kind, *one, 3.0, 300, seed=77 + p)
            assert st[p].tolist() == st1.cpu().tolist(), (kind, p)
            assert np.array_equal(H[p], H1.cpu().numpy()), (kind, p)
        assert st[3][0] == (0 if kind == 2 else 1)        # two valid points: enough for translation / similarity only
        assert st[0][0] == 1 and (kind != 2 or st[0][1] > 100)
    # counts = NULL: every pair uses max_pts (the -1 padding is ignored by the valid-index rule)
    H2, st2 = nm.ransac_batch(2, *d, None, 3.0, 300, seed=77)
    assert np.array_equal(st2.cpu().numpy()[[0, 1, 2, 5]], st[[0, 1, 2, 5]]) and np.array_equal(H2.cpu().numpy()[0], H[0])


def test_register_stream_recovers_the_motion_and_shards(nm):
    """BASELINE.json configs[4] with its consumer: frames cut from one scene at known offsets; SIFT, consecutive
    matching, align_points and one batched RANSAC recover the shifts; two ranks (one overlap frame) reproduce the
    single-rank homographies bit for bit."""
    from niftymatch_b200 import synth
    w, h = 512, 384
    base = synth.scene(w + 64, h + 64, synth.SEED_BASE + 9)
    offs = [(8, 8), (12, 10), (18, 9), (20, 16), (27, 20), (30, 26), (38, 30)]      # window origin of frame t
    frames = np.stack([base[oy: oy + h, ox: ox + w] for ox, oy in offs])
    P = nm.SiftParams(w, h)
    P._peak_threshold = 2.0
    sb = nm.SiftBatch(P, 4, 8192)
    fr = _cu(frames)
    pairs, H, st = nm.register_stream(sb, fr, kind=nm.TRANSLATION, inlier_threshold=1.0, iterations=256, seed=5, chunk=4)
    assert pairs == [(t, t + 1) for t in range(6)]
    H, st = H.cpu().numpy(), st.cpu().numpy()
    for t in range(6):
        sx, sy = offs[t][0] - offs[t + 1][0], offs[t][1] - offs[t + 1][1]          # a point moves by -(window motion)
        assert st[t][0] == 1 and st[t][1] > 30, st[t]
        assert abs(H[t][2] - sx) < 0.5 and abs(H[t][5] - sy) < 0.5, (t, H[t], sx, sy)
    Hh, sth = nm.register_stream(sb, fr, kind=nm.HOMOGRAPHY, inlier_threshold=1.0, iterations=512, seed=5, chunk=4)[1:]
    Hh = Hh.cpu().numpy()
    for t in range(6):
        sx, sy = offs[t][0] - offs[t + 1][0], offs[t][1] - offs[t + 1][1]
        g = Hh[t] / Hh[t][8]
        assert abs(g[2] - sx) < 1.0 and abs(g[5] - sy) < 1.0 and abs(g[0] - 1) < 1e-2 and abs(g[4] - 1) < 1e-2, (t, g)
    got = {}
    for rank in range(2):
        pr, Hr, sr = nm.register_stream(sb, fr, kind=nm.HOMOGRAPHY, inlier_threshold=1.0, iterations=512, seed=5, world=2,
                                        rank=rank, chunk=4)
        for k, (t, _) in enumerate(pr):
            got[t] = (Hr[k].cpu().numpy(), sr[k].cpu().numpy())
    assert sorted(got) == list(range(6))
    for t in range(6):
        assert np.array_equal(got[t][0], Hh[t]) and np.array_equal(got[t][1], sth.cpu().numpy()[t]), t
    sb.close()


def test_stream_registrar_device_pipeline_equals_the_per_pair_path(nm):
    """StreamRegistrar = configs[4] without host round trips: batched SIFT, nm_match_pairs_f32 (tcgen05 engine, device
    counts), nm_align_pairs_f32 and one batched RANSAC per chunk.  Its homographies and inlier counts must be BITWISE
    those of register_stream (one nm_match_f32 / align_points call per pair with host-side counts), whatever the chunk
    size, and a two-rank sharding with one overlap frame must reproduce the single-rank run."""
    from niftymatch_b200 import synth
    from niftymatch_b200.dist import StreamRegistrar, frame_range
    w, h = 512, 384
    base = synth.scene(w + 64, h + 64, synth.SEED_BASE + 9)
    offs = [(8, 8), (12, 10), (18, 9), (20, 16), (27, 20), (30, 26), (38, 30), (40, 35), (44, 41)]
    frames = np.stack([base[oy: oy + h, ox: ox + w] for ox, oy in offs])
    n = len(offs)
    P = nm.SiftParams(w, h)
    P._peak_threshold = 2.0
    sb = nm.SiftBatch(P, 5, 8192)
    fr = _cu(frames)
    _, Href, sref = nm.register_stream(sb, fr, kind=nm.HOMOGRAPHY, inlier_threshold=1.0, iterations=512, seed=5, chunk=4)
    Href, sref = Href.cpu().numpy(), sref.cpu().numpy()
    assert (sref[:, 0] == 1).all() and sref[:, 1].min() > 30
    frames_of = lambda t0, t1: fr[t0:t1].contiguous()
    for chunk in (5, 3, 2):
        reg = StreamRegistrar(sb, chunk=chunk, kind=nm.HOMOGRAPHY, inlier_threshold=1.0, iterations=512, seed=5)
        H, st = reg.run(frames_of, 0, n)
        assert np.array_equal(H.cpu().numpy(), Href), chunk
        assert np.array_equal(st.cpu().numpy(), sref), chunk
    reg = StreamRegistrar(sb, chunk=4, kind=nm.HOMOGRAPHY, inlier_threshold=1.0, iterations=512, seed=5)
    for world in (2, 3):
        got = []
        for rank in range(world):
            lo, hi = frame_range(n, world, rank, overlap=1)
            H, st = reg.run(frames_of, lo, hi)
            got.append((H.cpu().numpy(), st.cpu().numpy()))
        assert np.array_equal(np.concatenate([g[0] for g in got]), Href), world
        assert np.array_equal(np.concatenate([g[1] for g in got]), sref), world
    sb.close()


def test_dropin_ransac_header(nm):
    """ransac.h of the drop-in layer (compat/include/nm/ransac.h) through the same client code that drives the
    reference (oracle/ref_ransac_driver.cu built with -DNM_COMPAT_BUILD)."""
    client = os.path.join(os.path.dirname(GOLDEN), os.pardir, "build", "compat", "libnmcompat.so")
    if not os.path.exists(client):
        pytest.skip("build/compat/libnmcompat.so not built")
    cl = C.CDLL(os.path.abspath(client))
    g = _gold()
    c = [np.zeros(300, np.float32) for _ in range(4)]
    assert cl.nmcompat_align_points(_p(g["al_src_x"]), _p(g["al_src_y"]), 300, _p(g["al_dst_x"]), _p(g["al_dst_y"]), 260,
                                    _p(g["al_matches"]), *[_p(a) for a in c]) == 0
    for got, key in zip(c, ("al_c_src_x", "al_c_src_y", "al_c_dst_x", "al_c_dst_y")):
        assert np.array_equal(got, g[key]), key
    sx, sy, dx, dy, Ht = ransac_scene(n=800, seed=4, noise=0.2)
    os.environ["NM_RANSAC_SEED"] = "42"
    try:
        Hs = []
        for _ in range(2):
            H9 = np.zeros(9, np.float32)
            assert cl.nmcompat_ransac(2, _p(sx), _p(sy), _p(dx), _p(dy), len(sx), C.c_float(4.0), 1000, _p(H9)) == 1
            Hs.append(H9)
        assert np.array_equal(Hs[0], Hs[1])
        assert np.abs((Hs[0] / Hs[0][8]).reshape(3, 3)[:2, :2] - Ht[:2, :2]).max() < 5e-3
    finally:
        del os.environ["NM_RANSAC_SEED"]
    few = np.full(20, -1.0, np.float32)
    H9 = np.full(9, 3.0, np.float32)
    assert cl.nmcompat_ransac(2, _p(few), _p(few), _p(few), _p(few), 20, C.c_float(4.0), 100, _p(H9)) == 0

"""Capture golden vectors from the REFERENCE'S OWN CUDA code (oracle/_ref/libnmref.so, built
by oracle/build_ref.sh from /root/reference) on a B200.  Run on the GPU box:

    python tests/golden/make_golden.py gpurun_out/golden

and copy the .npz files into tests/golden/ (`... gpurun_out/golden masked` regenerates only the
masked-detector fixture, `... ransac` only the registration fixture).  The fixtures pin the CPU oracle
(tests/test_oracle_golden.py, no GPU needed) and the CUDA product (tests/test_gpu_*.py).

What the reference can and cannot produce on sm_100: everything up to the collated
keypoints, and the descriptors, run unmodified through its public API.  Its orientation
kernel (kernel_orientations_optim) deadlocks on sm_70+ (divergent __syncthreads), so
orientations are captured from the reference's other kernel (kernel_orientations_naive,
same arithmetic, no 10-px window clamp), and descriptors are captured with injected
orientations (the CPU oracle's public-API orientations), see oracle/ref_driver.cu.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from niftymatch_b200 import synth  # noqa: E402
from tests._util import load_oracle, load_reflib  # noqa: E402


def crafted_match_sets():
    """Descriptor sets with the edge cases of the reference's scan (match.cu:88-116)."""
    B = synth.descriptors(250, 2)
    A = synth.descriptors(200, 1, planted_from=B)
    # exact duplicates -> min1 = min2 = 0 -> entry left untouched (match.cu:107)
    B[9] = B[5]
    A[0] = B[5]
    # min2 start value 2139095040.0f only survives when column 0 is the minimum
    A[1] = 0; A[1, 1] = 1.0e6
    B[0] = A[1]; B[0, 0] = 42426.0            # d1 = 1.8e9 at index 0
    B[11] = A[1]; B[11, 0] = 54772.0          # d2 = 3.0e9  -> ratio vs init = 0.84 -> -1
    A[2] = 0; A[2, 2] = 1.0e6
    B[7] = A[2]; B[7, 0] = 42426.0            # d1 = 1.8e9 at index 7
    B[3] = A[2]; B[3, 0] = 54772.0            # d2 = 3.0e9  -> ratio 0.6 -> match 7
    # tie on the minimum: lowest index wins, min2 == min1 -> ratio 1 -> -1
    B[21] = B[20]
    A[3] = B[20] + 1.0
    m0 = np.full(200, 77, np.int32)           # sentinel initial values
    return A.astype(np.float32), B.astype(np.float32), m0


def detector_masks(w, h):
    """Masks for compute_keypoints_with_mask: a fetoscope-style circular field of view (binary), and a
    mask with a half-valued ring and a rectangular hole (exercises the `< 1` test on filtered samples:
    in octaves >= 1 the sample is the mean of four texels)."""
    yy, xx = np.mgrid[0:h, 0:w]
    r2 = (xx - w / 2 + 0.5) ** 2 + (yy - h / 2 + 0.5) ** 2
    fov = (r2 < (0.42 * h) ** 2).astype(np.float32)
    soft = np.where(r2 < (0.46 * h) ** 2, 1.0, 0.0).astype(np.float32)
    soft[(r2 >= (0.40 * h) ** 2) & (r2 < (0.46 * h) ** 2)] = 0.5
    soft[h // 3: h // 3 + 21, w // 4: w // 4 + 37] = 0.0
    return {"fov": fov, "soft": soft}


def masked(outdir):
    """compute_keypoints_with_mask of the reference (siftfunctions.cu:65-98) on a float mask texture."""
    os.makedirs(outdir, exist_ok=True)
    ref, orc = load_reflib(), load_oracle()
    assert ref is not None, "oracle/_ref/libnmref.so missing"
    w, h = 256, 192
    img = synth.scene(w, h, synth.SEED_BASE)
    out = {"image": img}
    for name, mask in detector_masks(w, h).items():
        r = ref.sift_frame(img, peak=0.0, want_levels=False, orient_mode=1, mask=mask)
        c = orc.sift_frame(img, peak=0.0, want_levels=False, orient_mode=0, mask=mask)
        assert np.array_equal(r["seg_counts"], c["seg_counts"]), (name, r["seg_counts"], c["seg_counts"])
        ri = ref.sift_frame(img, peak=0.0, want_levels=False, orient_mode=2, orient_in=c["orient"], mask=mask)
        out[f"{name}_mask"] = mask
        out[f"{name}_seg_counts"] = r["seg_counts"]
        out[f"{name}_kpts"] = r["kpts"]
        out[f"{name}_orient_in"] = c["orient"]
        out[f"{name}_desc"] = ri["desc"]
        out[f"{name}_x"] = ri["x"]
        out[f"{name}_y"] = ri["y"]
        print("masked", name, "n =", r["n"], r["seg_counts"])
    np.savez_compressed(os.path.join(outdir, f"sift_{w}x{h}_masked.npz"), **out)


def ransac(outdir):
    """The reference's hypothesis kernels (ransac.cu:437-520, Jacobi SVD of svd.cu) on supplied index lists,
    and align_points (ransac.cu:29-59)."""
    from tests._util import ransac_scene, ransac_rand_lists, checker_ransac_hypotheses, _p
    os.makedirs(outdir, exist_ok=True)
    ref = load_reflib()
    assert ref is not None, "oracle/_ref/libnmref.so missing"
    sx, sy, dx, dy, Ht = ransac_scene()
    out = {"src_x": sx, "src_y": sy, "dst_x": dx, "dst_y": dy, "H_true": Ht, "thr": np.float32(4.0)}
    for kind, rl in ransac_rand_lists(sx).items():
        H, inl = checker_ransac_hypotheses(ref.lib, "nmref", kind, sx, sy, dx, dy, rl, 4.0)
        out[f"rand_{kind}"] = rl
        out[f"H_{kind}"] = H
        out[f"inliers_{kind}"] = inl
        print("ransac kind", kind, "best", inl.max(), "at", int(inl.argmax()), "zero rows", int((H == 0).all(axis=1).sum()))
    # align_points: 300 source points, 260 destination points, a fifth unmatched
    rng = np.random.default_rng(3)
    ax, ay = (rng.random(300) * 640).astype(np.float32), (rng.random(300) * 480).astype(np.float32)
    bx, by = (rng.random(260) * 640).astype(np.float32), (rng.random(260) * 480).astype(np.float32)
    m = rng.integers(0, 260, 300).astype(np.int32)
    m[rng.random(300) < 0.2] = -1
    c = [np.zeros(300, np.float32) for _ in range(4)]
    rc = ref.lib.nmref_align_points(_p(ax), _p(ay), 300, _p(bx), _p(by), 260, _p(m), *[_p(a) for a in c])
    assert rc == 0, rc
    out.update(al_src_x=ax, al_src_y=ay, al_dst_x=bx, al_dst_y=by, al_matches=m,
               al_c_src_x=c[0], al_c_src_y=c[1], al_c_dst_x=c[2], al_c_dst_y=c[3])
    np.savez_compressed(os.path.join(outdir, "ransac_400.npz"), **out)


def preprocess(outdir):
    """The reference's cuda_grayscale<float>, cuda_cast<float, uchar>, cuda_undistort and resample_undistort."""
    from tests._util import preprocess_inputs, _p
    os.makedirs(outdir, exist_ok=True)
    ref = load_reflib()
    assert ref is not None, "oracle/_ref/libnmref.so missing"
    d = preprocess_inputs()
    h, w = d["fimg"].shape
    out = dict(d)
    gray = np.zeros((h, w), np.float32)
    assert ref.lib.nmref_grayscale(_p(d["bgra"]), w, h, _p(gray)) == 0
    out["gray"] = gray
    for mv in (0, 200):
        c = np.zeros((h, w), np.uint8)
        assert ref.lib.nmref_cast(_p(d["fimg"]), w, h, _p(c), mv) == 0
        out[f"cast_{mv}"] = c
    u, v = np.zeros((h, w), np.float32), np.zeros((h, w), np.float32)
    assert ref.lib.nmref_undistort(_p(d["x"]), _p(d["y"]), w, h, _p(d["cam"]), _p(d["dist"]), _p(u), _p(v)) == 0
    out["u"], out["v"] = u, v
    res = np.zeros((h, w), np.float32)
    assert ref.lib.nmref_resample_undistort(_p(d["gray8"]), w, h, _p(u), _p(v), w, h, _p(res)) == 0
    out["resampled"] = res
    chans = np.zeros((4, h, w), np.float32)
    rot = np.zeros((h, w, 4), np.uint8)
    assert ref.lib.nmref_channels(_p(d["bgra"]), w, h, 77, _p(chans), _p(rot)) == 0
    out["channels"], out["rotated_alpha77"] = chans, rot
    half = np.zeros((h // 2, w // 2, 4), np.uint8)
    assert ref.lib.nmref_downsample_bgra(_p(d["bgra"]), w, h, _p(half)) == 0
    out["bgra_half"] = half
    print("preprocess: gray", gray.min(), gray.max(), "cast", out["cast_0"][0, :8], out["cast_0"][1, :6], "u", u.min(), u.max(),
          "resampled", res.min(), res.max())
    np.savez_compressed(os.path.join(outdir, "preprocess_160x96.npz"), **out)


def mosaic(outdir):
    """The reference's resample_perspective_transform, resample_mask and transform_blend (resample.h:7-23)."""
    from tests._util import mosaic_inputs, _p
    os.makedirs(outdir, exist_ok=True)
    ref = load_reflib()
    assert ref is not None, "oracle/_ref/libnmref.so missing"
    d = mosaic_inputs()
    fh, fw = d["mask"].shape
    out = dict(d)
    cols, rows = 150, 100
    for inv in (0, 1):
        res = np.zeros((rows, cols, 4), np.uint8)
        xp, yp = np.zeros((rows, cols), np.float32), np.zeros((rows, cols), np.float32)
        assert ref.lib.nmref_resample_perspective(_p(d["frame"]), fw, fh, _p(d["mats"][1]), inv, cols, rows, _p(res), _p(xp), _p(yp)) == 0
        out[f"persp_{inv}"], out[f"xpos_{inv}"], out[f"ypos_{inv}"] = res, xp, yp
        m = np.zeros((rows, cols), np.uint8)
        assert ref.lib.nmref_resample_mask(_p(d["mask"]), fw, fh, _p(xp), _p(yp), cols, rows, C.c_float(0.5), _p(m)) == 0
        out[f"maskres_{inv}"] = m
    cw, ch = 170, 140
    canvas, cwts = np.zeros((ch, cw, 4), np.uint8), np.zeros((ch, cw), np.float32)
    assert ref.lib.nmref_transform_blend(_p(d["frame"]), _p(d["mask"]), _p(d["wts"]), fw, fh, 3, _p(d["mats"]), _p(d["tx"]), _p(d["ty"]),
                                         fw + 10, fh + 10, cw, ch, _p(canvas), _p(cwts)) == 0
    out["canvas"], out["canvas_wts"] = canvas, cwts
    print("mosaic: painted", int((cwts > 0).sum()), "of", cw * ch, "max weight", cwts.max(), "mask kept", int((out["maskres_1"] > 0).sum()))
    np.savez_compressed(os.path.join(outdir, "mosaic_128x90.npz"), **out)


def orient_pub(outdir):
    """Orientations and descriptors through the reference's PUBLIC orientation kernel (kernel_orientations_optim,
    orientation.cu:11-129) in the terminating form oracle/build_ref.sh builds (its two divergent __syncthreads()
    hoisted out of the branch, arithmetic untouched): the 10-pixel window clamp (:29-30), the smoothing with the
    in-place hist[35] update (:78-80), first-two-peaks (:118-127), then compute_descriptors on ITS orientations.
    The kernel has a write/read race on hist[35] (thread 0 writes it in place while thread 34 reads it),
    so the capture is repeated and the run-to-run spread is stored beside the values."""
    os.makedirs(outdir, exist_ok=True)
    ref, orc = load_reflib(), load_oracle()
    assert ref is not None, "oracle/_ref/libnmref.so missing"
    out = {}
    for (w, h, seed) in [(256, 192, synth.SEED_BASE), (384, 256, synth.SEED_BASE + 3)]:
        img = synth.scene(w, h, seed)
        for peak in (0.0, 2.0):
            tag = f"{w}x{h}_p{int(peak)}"
            runs = [ref.sift_frame(img, peak=peak, want_levels=False, orient_mode=3) for _ in range(5)]
            r = runs[0]
            c = orc.sift_frame(img, peak=peak, want_levels=False, orient_mode=0)
            assert np.array_equal(r["seg_counts"], c["seg_counts"])
            spread = max(float(np.abs(q["orient"] - r["orient"]).max()) for q in runs[1:])
            dspread = max(float(np.abs(q["desc"] - r["desc"]).max() / max(np.abs(r["desc"]).max(), 1e-20)) for q in runs[1:])
            out[f"{tag}_seg_counts"] = r["seg_counts"]
            out[f"{tag}_kpts"] = r["kpts"]
            out[f"{tag}_orient"] = r["orient"]
            out[f"{tag}_desc"] = r["desc"]
            out[f"{tag}_x"], out[f"{tag}_y"] = r["x"], r["y"]
            out[f"{tag}_run_spread"] = np.array([spread, dspread])
            ok = (r["orient"] >= 0) & (c["orient"] >= 0)
            d = np.abs(r["orient"] - c["orient"])[ok]
            d = np.minimum(d, 2 * np.pi - d)
            print("orient_pub", tag, "n =", r["n"], "run-to-run spread", spread, dspread,
                  "| vs oracle mode 0: peaks differ", int(((r["orient"] >= 0) != (c["orient"] >= 0)).sum()),
                  "max", float(d.max()) if d.size else 0.0, "count > 1e-5:", int((d > 1e-5).sum()), "> 1e-3:", int((d > 1e-3).sum()))
    np.savez_compressed(os.path.join(outdir, "sift_orient_public.npz"), **out)


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    ref, orc = load_reflib(), load_oracle()
    assert ref is not None, "oracle/_ref/libnmref.so missing"
    # ---- parameters and taps ----
    par = {}
    for (w, h) in [(256, 192), (640, 480), (1920, 1080), (3840, 2160)]:
        no, sk, s0, sd, bs = C.c_int(), C.c_float(), C.c_float(), C.c_float(), C.c_float()
        sg = (C.c_float * 5)()
        ref.lib.nmref_params(w, h, C.byref(no), C.byref(sk), C.byref(s0), C.byref(sd), C.byref(bs), sg)
        par[f"params_{w}x{h}"] = np.array([no.value, sk.value, s0.value, sd.value, bs.value] + list(sg), np.float64)
    t = (C.c_float * 96)()
    for which in range(-1, 5):
        r = ref.lib.nmref_taps(640, 480, which, t)
        par[f"taps_{which}"] = np.array(t[: 2 * r + 1], np.float32)
    np.savez_compressed(os.path.join(outdir, "params_taps.npz"), **par)
    # ---- SIFT frames ----
    for (w, h, seed) in [(256, 192, synth.SEED_BASE), (384, 256, synth.SEED_BASE + 3)]:
        img = synth.scene(w, h, seed)
        out = {"image": img}
        for peak in (0.0, 2.0):
            tag = f"p{int(peak)}"
            r = ref.sift_frame(img, peak=peak, want_grad=True, orient_mode=1)
            c = orc.sift_frame(img, peak=peak, orient_mode=0)
            assert np.array_equal(r["seg_counts"], c["seg_counts"]), "oracle/reference keypoint counts differ"
            ri = ref.sift_frame(img, peak=peak, want_levels=False, orient_mode=2, orient_in=c["orient"])
            out[f"{tag}_seg_counts"] = r["seg_counts"]
            out[f"{tag}_kpts"] = r["kpts"]
            out[f"{tag}_orient_naive"] = r["orient"]
            out[f"{tag}_orient_in"] = c["orient"]
            out[f"{tag}_desc"] = ri["desc"]
            out[f"{tag}_x"] = ri["x"]
            out[f"{tag}_y"] = ri["y"]
            if peak == 0.0:
                for o in range(r["n_oct"]):
                    out[f"level5_oct{o}"] = r["levels"][o][5]
                    out[f"level3_oct{o}"] = r["levels"][o][3]
                last = r["n_oct"] - 1
                for l in range(6):
                    out[f"level{l}_oct{last}"] = r["levels"][last][l]
                out[f"grad_oct{last}"] = r["grad"][last]
            print(w, h, peak, "n =", r["n"], r["seg_counts"])
        np.savez_compressed(os.path.join(outdir, f"sift_{w}x{h}.npz"), **out)
    # ---- matcher ----
    A, B, m0 = crafted_match_sets()
    m, D = ref.match(A, B, 0.8, match_io=m0, want_distance=True)
    print("match: matched", int((m >= 0).sum()), "untouched", int((m == 77).sum()), "rows 0..3:", m[:4])
    np.savez_compressed(os.path.join(outdir, "match_200x250.npz"), A=A, B=B, m0=m0, m=m, D=D)
    # convolution with a width that is not a multiple of 16 and an odd radius / generic taps
    img = synth.scene(208, 144, synth.SEED_BASE + 9)
    taps = np.array(par["taps_2"], np.float32)
    res = np.zeros_like(img)
    ref.lib.nmref_convolve(res.ctypes.data_as(C.c_void_p), img.ctypes.data_as(C.c_void_p), 208, 144,
                           taps.ctypes.data_as(C.c_void_p), (len(taps) - 1) // 2)
    np.savez_compressed(os.path.join(outdir, "convolve_208x144.npz"), image=img, taps=taps, result=res)


if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
    if len(sys.argv) > 2 and sys.argv[2] == "masked":      # only the masked-detector fixture
        masked(out)
    elif len(sys.argv) > 2 and sys.argv[2] == "ransac":    # only the registration fixture
        ransac(out)
    elif len(sys.argv) > 2 and sys.argv[2] == "preprocess":
        preprocess(out)
    elif len(sys.argv) > 2 and sys.argv[2] == "mosaic":
        mosaic(out)
    elif len(sys.argv) > 2 and sys.argv[2] == "orient":
        orient_pub(out)
    else:
        main(out)
        masked(out)
        ransac(out)
        preprocess(out)
        mosaic(out)
        orient_pub(out)

"""Stage-by-stage parity report: product (CUDA, through the C-ABI) vs the reference's own
CUDA build (oracle/_ref) vs the CPU oracle, on identical synthetic inputs.  GPU box only.
Usage: python tools/parity_report.py [--sizes 256x192,640x480] [--no-oracle]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import niftymatch_b200 as nm  # noqa: E402
from niftymatch_b200 import synth  # noqa: E402
from tests._util import load_oracle, load_reflib, match_keypoints, ang_diff  # noqa: E402


def product_frame(image, peak, capacity=65536, exact=False, num_octaves=-1):
    h, w = image.shape
    P = nm.SiftParams(w, h)
    P._peak_threshold = peak
    if num_octaves > 0:
        P._num_octaves = num_octaves
    sb = nm.SiftBatch(P, 1, capacity)
    sb.set_dense_gradients(True)          # the whole maps are compared below
    sb.set_exact_descriptor(exact)
    fr = torch.from_numpy(image[None]).cuda()
    sb.run(fr)
    torch.cuda.synchronize()
    r = sb.results()
    n = int(r["counts"][0].item())
    out = {"n": n, "n_oct": P._num_octaves, "desc": r["desc"][0, :n].cpu().numpy(), "x": r["x"][0, :n].cpu().numpy(),
           "y": r["y"][0, :n].cpu().numpy(), "kpts": r["kpts"][0, :n].cpu().numpy(),
           "orient": r["orient"][0, :n].cpu().numpy(), "seg_counts": r["seg_counts"][0].cpu().numpy()}
    out["levels"] = [[sb.level(0, o, l).cpu().numpy() for l in range(6)] for o in range(P._num_octaves)]
    out["grad"] = [np.stack([sb.grad(0, o, l).cpu().numpy() for l in range(3)]) for o in range(P._num_octaves)]
    sb.close()
    return out


def cmp_frames(name_a, a, name_b, b, check_grad=False):
    print(f"  -- {name_a} vs {name_b}: n={a['n']}/{b['n']} seg_counts equal={np.array_equal(a['seg_counts'], b['seg_counts'])}")
    if not np.array_equal(a['seg_counts'], b['seg_counts']):
        print("     seg a:", a['seg_counts'], "\n     seg b:", b['seg_counts'])
    if "levels" in a and "levels" in b:
        for o in range(min(len(a["levels"]), len(b["levels"]))):
            row = []
            for l in range(6):
                d = a["levels"][o][l] != b["levels"][o][l]
                row.append(f"{int(d.sum())}")
            print(f"     octave {o}: mismatching pixels per level (bitwise): {' '.join(row)}", end="")
            mx = max(float(np.abs(a['levels'][o][l] - b['levels'][o][l]).max()) for l in range(6))
            print(f"   max|diff|={mx:.3g}")
    if check_grad and "grad" in a and "grad" in b:
        for o in range(min(len(a["grad"]), len(b["grad"]))):
            ga, gb = a["grad"][o], b["grad"][o]
            inner = (slice(None), slice(1, -1), slice(1, -1))
            dm = np.abs(ga[inner][..., 0] - gb[inner][..., 0]).max()
            da = ang_diff(ga[inner][..., 1], gb[inner][..., 1]).max()
            nb = int((ga[inner] != gb[inner]).sum())
            print(f"     grad octave {o}: max|dmag|={dm:.3g} max|dang|={da:.3g} bitwise-different values={nb}")
    ka, kb = a["kpts"], b["kpts"]
    if len(ka) == len(kb) and len(ka):
        same = (ka == kb).all(axis=1).sum()
        print(f"     keypoints bitwise identical: {same}/{len(ka)}  max|dpos|={np.abs(ka[:, :2] - kb[:, :2]).max():.3g} max|dsigma|={np.abs(ka[:, 2] - kb[:, 2]).max():.3g}")
    ia, ib = match_keypoints(ka, kb, 0.01)
    print(f"     keypoints associated within 0.01 px: {len(ia)} of {len(ka)} / {len(kb)}  (recall {len(ia) / max(1, len(kb)):.4f})")
    if len(ia) == 0:
        return
    oa, ob = a["orient"][ia], b["orient"][ib]
    d0 = ang_diff(oa[:, 0], ob[:, 0])
    both = (oa[:, 0] >= 0) & (ob[:, 0] >= 0)
    print(f"     orientation[0]: max diff={d0[both].max() if both.any() else 0:.3g}  >1e-3: {(d0[both] > 1e-3).sum()} of {both.sum()};  "
          f"-1 mismatch: {((oa[:, 0] < 0) != (ob[:, 0] < 0)).sum()}")
    bad = np.where(d0 > 1e-3)[0]
    for q in bad[:8]:
        print(f"        kp {ia[q]}: {name_a} th={oa[q]} {name_b} th={ob[q]} bin={ob[q, 0] / (2 * np.pi) * 36:.2f} kp={ka[ia[q]]}")
    n = min(a["n"], b["n"])
    sel = [(i, j) for i, j in zip(ia, ib) if i < a["n"] and j < b["n"]]
    if sel:
        ii = np.array([s[0] for s in sel]); jj = np.array([s[1] for s in sel])
        da, db = a["desc"][ii], b["desc"][jj]
        rel = np.linalg.norm(da - db, axis=1) / np.maximum(np.linalg.norm(db, axis=1), 1e-20)
        okori = d0[: len(sel)] <= 1e-3
        print(f"     descriptor rel-L2: median={np.median(rel):.3g} max={rel.max():.3g}  >1e-3: {(rel > 1e-3).sum()} of {len(rel)}"
              f"  (with orientation agreeing: {(rel[okori] > 1e-3).sum()} of {okori.sum()})")
        print(f"     x/y out max diff: {np.abs(a['x'][ii] - b['x'][jj]).max():.3g} {np.abs(a['y'][ii] - b['y'][jj]).max():.3g}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="256x192,640x480")
    ap.add_argument("--no-oracle", action="store_true")
    ap.add_argument("--peaks", default="0,2")
    args = ap.parse_args()
    ref = load_reflib()
    orc = None if args.no_oracle else load_oracle()
    print("device:", torch.cuda.get_device_name(0), "ref lib:", ref is not None)
    for size in args.sizes.split(","):
        w, h = map(int, size.split("x"))
        img = synth.scene(w, h, synth.SEED_BASE)
        for peak in map(float, args.peaks.split(",")):
            print(f"== {w}x{h} peak={peak}")
            t = time.time(); p = product_frame(img, peak); tp = time.time() - t
            pe = product_frame(img, peak, exact=True)
            print(f"   product n={p['n']} seg={p['seg_counts']} ({tp:.2f}s incl. setup)")
            cmp_frames("product(fp32 desc)", p, "product(exact desc)", pe)
            if ref is not None:
                # (1) reference with its own naive orientation kernel (the used one deadlocks on sm_70+)
                t = time.time(); r = ref.sift_frame(img, peak=peak, want_grad=True, orient_mode=1); tr = time.time() - t
                print(f"   reference-cuda[naive orientation kernel] n={r['n']} ({tr:.2f}s)")
                cmp_frames("product", p, "ref-cuda[naive-orient]", r, check_grad=True)
                # (2) reference descriptors on the product's orientations (injected)
                if np.array_equal(p["seg_counts"], r["seg_counts"]) and p["n"] == len(p["kpts"]):
                    ri = ref.sift_frame(img, peak=peak, want_levels=False, orient_mode=2, orient_in=p["orient"])
                    print("   reference-cuda descriptors with the product's orientations injected:")
                    cmp_frames("product", p, "ref-cuda[injected]", ri)
                    cmp_frames("product(exact desc)", pe, "ref-cuda[injected]", ri)
                    ri2 = ref.sift_frame(img, peak=peak, want_levels=False, orient_mode=2, orient_in=p["orient"])
                    rr = np.linalg.norm(ri['desc'] - ri2['desc'], axis=1) / np.maximum(np.linalg.norm(ri['desc'], axis=1), 1e-20)
                    print(f"   ref-cuda run-to-run descriptor rel diff (float atomics): max={rr.max():.3g}")
            if orc is not None and w * h <= 640 * 480:
                t = time.time(); c = orc.sift_frame(img, peak=peak, want_grad=True, orient_mode=0); tc = time.time() - t
                print(f"   cpu-oracle n={c['n']} ({tc:.2f}s)")
                cmp_frames("product", p, "cpu-oracle", c, check_grad=True)
                if ref is not None:
                    c1 = orc.sift_frame(img, peak=peak, want_grad=True, orient_mode=1)
                    cmp_frames("cpu-oracle[naive-orient]", c1, "ref-cuda[naive-orient]", r, check_grad=True)
    # ---- matcher -------------------------------------------------------------------
    print("== matcher")
    B = synth.descriptors(700, 2)
    A = synth.descriptors(500, 1, planted_from=B)
    At, Bt = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    m, D = nm.match(At, Bt, 0.8, want_distance=True)
    m2 = nm.match(At, Bt, 0.8)
    torch.cuda.synchronize()
    print("   engine:", nm.get_engine(), " matched:", int((m >= 0).sum()), " with/without distance equal:", bool((m == m2).all()))
    if ref is not None:
        mr, Dr = ref.match(A, B, 0.8, want_distance=True)
        print("   vs ref-cuda: indices equal:", np.array_equal(m.cpu().numpy(), mr), " D bitwise equal:", np.array_equal(D.cpu().numpy(), Dr))
    if orc is not None:
        mo, Do = orc.match(A, B, 0.8, want_distance=True)
        print("   vs cpu-oracle: indices equal:", np.array_equal(m.cpu().numpy(), mo), " D bitwise equal:", np.array_equal(D.cpu().numpy(), Do))


if __name__ == "__main__":
    main()

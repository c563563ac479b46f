"""Registration parity report on a GPU box: product (C-ABI) vs the reference's kernels (oracle/_ref) vs the CPU
oracle on the fixture scene of tests/_util.py (development aid)."""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import niftymatch_b200 as nm  # noqa: E402
from tests._util import (load_oracle, load_reflib, ransac_scene, ransac_rand_lists, checker_ransac_hypotheses,  # noqa: E402
                         normalise_h)

orc, ref = load_oracle(), load_reflib()
sx, sy, dx, dy, Ht = ransac_scene()
cu = lambda a: torch.from_numpy(a).cuda()
for kind, rl in ransac_rand_lists(sx).items():
    Hp, ip = nm.ransac_hypotheses(kind, cu(sx), cu(sy), cu(dx), cu(dy), cu(rl), 4.0)
    Hp, ip = Hp.cpu().numpy(), ip.cpu().numpy()
    Ho, io = checker_ransac_hypotheses(orc.lib, "orc", kind, sx, sy, dx, dy, rl, 4.0)
    line = f"kind {kind}: product best {ip.max()}@{ip.argmax()}  oracle best {io.max()}@{io.argmax()}  |inl p-o| max {np.abs(ip - io).max()}"
    line += f"  H bitwise p==o {np.array_equal(Hp, Ho)}  |Hn p-o| max {np.abs(normalise_h(Hp) - normalise_h(Ho)).max():.3g}"
    if ref is not None:
        Hr, ir = checker_ransac_hypotheses(ref.lib, "nmref", kind, sx, sy, dx, dy, rl, 4.0)
        line += f"\n        reference best {ir.max()}@{ir.argmax()}  |inl p-r| max {np.abs(ip - ir).max()} (n diff {int((ip != ir).sum())})"
        line += f"  H bitwise p==r {np.array_equal(Hp, Hr)}  |Hn p-r| max {np.abs(normalise_h(Hp) - normalise_h(Hr)).max():.3g}"
        line += f"  |Hn o-r| max {np.abs(normalise_h(Ho) - normalise_h(Hr)).max():.3g} |inl o-r| max {np.abs(io - ir).max()}"
    print(line)
    for seed in (1, 2):
        H, st = nm.ransac(kind, cu(sx), cu(sy), cu(dx), cu(dy), 4.0, 2000, seed)
        H2, st2 = nm.ransac(kind, cu(sx), cu(sy), cu(dx), cu(dy), 4.0, 2000, seed)
        print("   full", seed, st.cpu().tolist(), "repeatable", torch.equal(H, H2) and torch.equal(st, st2),
              (H.cpu().numpy() / max(abs(float(H[8])), 1e-30)).round(5).tolist())
    if ref is not None:
        H9 = np.zeros(9, np.float32)
        import ctypes as C
        P = lambda v: v.ctypes.data_as(C.c_void_p)
        ok = ref.lib.nmref_ransac(kind, P(sx), P(sy), P(dx), P(dy), len(sx), C.c_float(4.0), 2000, P(H9))
        print("   reference full", ok, (H9 / max(abs(H9[8]), 1e-30)).round(5).tolist())
def timing():
    # timing
    n = 8192
    sx, sy, dx, dy, _ = ransac_scene(n=n, seed=9)
    a = [cu(v) for v in (sx, sy, dx, dy)]
    for kind in (0, 1, 2):
        for it in (1024, 4096):
            for _ in range(5):
                nm.ransac(kind, *a, 4.0, it, 1)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                nm.ransac(kind, *a, 4.0, it, 1)
            e1.record()
            torch.cuda.synchronize()
            print(f"nm_ransac_f32 kind {kind} n {n} iterations {it}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us")

    # batched: 64 frame pairs in one launch sequence
    npairs = 64
    arrs = [torch.stack([cu(ransac_scene(n=n, seed=100 + p)[k]) for p in range(npairs)]).contiguous() for k in range(4)]
    for kind in (0, 2):
        for _ in range(3):
            nm.ransac_batch(kind, *arrs, None, 4.0, 1024, 1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            nm.ransac_batch(kind, *arrs, None, 4.0, 1024, 1)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"nm_ransac_batch_f32 kind {kind}: {npairs} pairs x {n} correspondences x 1024 hypotheses: {ms * 1e3:.0f} us = {ms * 1e3 / npairs:.1f} us per pair")


for rep in range(2):
    print("timing pass", rep)
    timing()
